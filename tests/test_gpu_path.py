"""End-to-end parity of the CUDA path against the oracle / the golden vectors frozen from the reference.

Tolerances (BASELINE.json north_star): fp16 storage — logits within 2e-2 max-abs of the fp32 reference,
identical real/fake verdicts at threshold 0.5.  bf16 storage is compared against the bf16 emulation of the
same rounding points (tests/bf16_emulation.py) and, loosely, against the fp32 oracle: it does NOT meet
2e-2 on this checkpoint, which is why fp16 is the default."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_LOGITS_FP16 = 2e-2
TOL_FEAT_REL_FP16 = 1.5e-2


@pytest.fixture(scope="module")
def scorer(synth_sd):
    from deepfake_video_detection_b200 import FrameScorer
    return FrameScorer(synth_sd, "fp16", "cuda")


def test_golden_videos_fp16(scorer, golden, golden_crops):
    from deepfake_video_detection_b200 import decide, make_offsets
    from oracle import effnet_b0_oracle as O
    crops, offsets = golden_crops
    lens = np.diff(offsets)
    x = torch.from_numpy(crops).cuda()
    logits, scores, feat = scorer.score(x, make_offsets(lens, "cuda"), return_features=True)
    ref_feat = torch.from_numpy(golden["features"])
    rel = ((feat.cpu() - ref_feat).norm() / ref_feat.norm()).item()
    assert rel < TOL_FEAT_REL_FP16, f"feature relative error {rel}"
    err = (logits.cpu() - torch.from_numpy(golden["logits"])).abs().max().item()
    assert err <= TOL_LOGITS_FP16, f"logit max-abs error {err}"
    assert (scores.cpu() - torch.from_numpy(golden["frame_scores"])).abs().max().item() < 5e-3
    ours = [d["is_fake"] for d in decide(logits)]
    ref = [d["is_fake"] for d in O.decide(torch.from_numpy(golden["logits"]))]
    assert ours == ref


def test_mean_pool_mode(scorer, golden, golden_crops):
    from deepfake_video_detection_b200 import make_offsets
    crops, offsets = golden_crops
    x = torch.from_numpy(crops[: offsets[4]]).cuda()
    logits, scores = scorer.score(x, make_offsets([8, 8, 8, 8], "cuda"), use_temporal_attention=False)
    assert (logits.cpu() - torch.from_numpy(golden["meanpool_logits"])).abs().max().item() <= TOL_LOGITS_FP16
    assert torch.allclose(scores.cpu(), torch.full((32,), 0.125))


def test_input_layouts_agree(scorer, golden_crops):
    """u8 HWC (fused prep), fp32 NCHW (forward()'s input) and K1's 16-bit NCHW give the same features."""
    from oracle import effnet_b0_oracle as O
    crops, _ = golden_crops
    u8 = torch.from_numpy(crops[:6]).cuda()
    f_u8 = scorer.features(u8)
    f_f32 = scorer.features(O.prep_u8_hwc(crops[:6]).cuda().contiguous())
    f_h16 = scorer.features(scorer.preprocess(u8))
    # uint8 goes through the tcgen05 stem (hi/lo split operands), fp32 through the CUDA-core stem: same maths to ~1e-6
    assert ((f_u8 - f_f32).norm() / f_f32.norm()).item() < 6e-3     # last-bit flips of the stem output, amplified by the trunk
    assert ((f_h16 - f_u8).norm() / f_u8.norm()).item() < 1.5e-2   # one extra fp16 rounding of the input, amplified by the trunk


def test_module_forward_is_dropin(synth_sd, golden, golden_crops):
    from deepfake_video_detection_b200 import PretrainedBackboneDetector, imagenet_normalize
    crops, offsets = golden_crops
    model = PretrainedBackboneDetector("efficientnet_b0", pretrained=False, num_classes=2, dropout_rate=0.5,
                                       use_temporal_attention=True)
    model.load_state_dict(synth_sd, strict=True)
    model.to("cuda").eval()
    faces = torch.from_numpy(crops[: offsets[4]]).view(4, 8, 224, 224, 3)
    x = imagenet_normalize(faces.permute(0, 1, 4, 2, 3).float() / 255.0).cuda()       # app.py:2084-2086, batched
    with torch.no_grad():
        logits, frame_scores = model(x)
    assert logits.shape == (4, 2) and frame_scores.shape == (4, 8)
    assert (logits.cpu() - torch.from_numpy(golden["batched_logits"])).abs().max().item() <= TOL_LOGITS_FP16
    assert (frame_scores.cpu() - torch.from_numpy(golden["batched_frame_scores"])).abs().max().item() < 5e-3
    # weights changed in place -> engine repacks
    with torch.no_grad():
        model.fc2.bias.add_(1.0)
        logits2, _ = model(x)
    assert torch.allclose(logits2, logits + 1.0, atol=1e-5)


def test_bf16_storage_against_the_oracle_with_its_measured_tolerance(synth_sd, golden, golden_crops):
    """bf16 storage (precision="bf16") does NOT meet north_star's 2e-2: 8-bit significands cost 8x the fp16 error on this
    checkpoint (profiles/r02_parity_diag.json: max 5.8e-2 on the 64 x 32 batch, 0.36 on single-frame videos).  It is kept
    selectable and held to what it measures against the ORACLE (no self-comparison): features within 5 %, golden logits
    within 0.15, verdicts equal outside the band that tolerance implies."""
    from deepfake_video_detection_b200 import FrameScorer, decide, make_offsets
    from oracle import effnet_b0_oracle as O
    crops, offsets = golden_crops
    lens = np.diff(offsets)
    sc = FrameScorer(synth_sd, "bf16", "cuda")
    logits, scores, feat = sc.score(torch.from_numpy(crops).cuda(), make_offsets(lens, "cuda"), return_features=True)
    ref_feat, ref_logits = torch.from_numpy(golden["features"]), torch.from_numpy(golden["logits"])
    assert ((feat.cpu() - ref_feat).norm() / ref_feat.norm()).item() < 5e-2
    err = (logits.cpu() - ref_logits).abs().max().item()
    assert 2e-2 < err <= 0.15 or err <= 2e-2, err
    for o, r in zip(decide(logits), O.decide(ref_logits)):
        assert o["is_fake"] == r["is_fake"] or abs(r["prob_fake"] - 0.5) < 2 * 0.15


def test_determinism_and_chunking(scorer, golden_crops, monkeypatch):
    crops, _ = golden_crops
    x = torch.from_numpy(crops[:10]).cuda()
    a = scorer.features(x)
    b = scorer.features(x)
    assert torch.equal(a, b)
    monkeypatch.setenv("DFD_CHUNK_FRAMES", "3")           # 10 frames -> chunks of 3,3,3,1
    c = scorer.features(x)
    assert torch.equal(a, c)


def test_errors_are_loud(scorer):
    with pytest.raises(RuntimeError):
        scorer.features(torch.zeros((1, 224, 224, 3), dtype=torch.uint8))          # CPU tensor
    with pytest.raises(RuntimeError):
        scorer.features(torch.zeros((1, 100, 100, 3), dtype=torch.uint8, device="cuda"))   # unsupported size


def test_score_host_matches_device_path(scorer, golden_crops):
    """Public end-to-end entry (pinned host crops, overlapped H2D): same bits as scoring device-resident crops."""
    from deepfake_video_detection_b200 import make_offsets
    crops, offsets = golden_crops
    lens = np.diff(offsets).tolist()
    host = torch.from_numpy(crops).pin_memory()
    ref_l, ref_s = scorer.score(torch.from_numpy(crops).cuda(), make_offsets(lens, "cuda"))
    for chunk in (None, 3, 1):
        for _ in range(2):                                  # second call reuses the staging buffers
            lg, sc = scorer.score_host(host, lens, chunk_videos=chunk)
            assert torch.equal(lg, ref_l) and torch.equal(sc, ref_s)


def test_score_host_right_after_score_and_staging_regrow(scorer, golden_crops):
    """No synchronisation between calls: score() leaves work in flight on the main stream, then score_host() allocates its
    staging buffers (which may alias memory that work still uses) and copies on a side stream; a later, larger call regrows
    the staging buffers while the previous call is still running.  Results must be the bits of the synchronous path."""
    from deepfake_video_detection_b200 import FrameScorer, make_offsets
    crops, offsets = golden_crops
    lens = np.diff(offsets).tolist()
    dev = torch.from_numpy(crops).cuda()
    off = make_offsets(lens, "cuda")
    ref_l, ref_s = scorer.score(dev, off)
    ref_small, _ = scorer.score(dev[: offsets[2]], make_offsets(lens[:2], "cuda"))
    torch.cuda.synchronize()
    host_small = torch.from_numpy(crops[: offsets[2]]).pin_memory()
    host_all = torch.from_numpy(crops).pin_memory()
    for _ in range(3):
        fresh = FrameScorer.__new__(FrameScorer)                  # same packed weights, no staging / workspace yet
        fresh.__dict__.update({k: v for k, v in scorer.__dict__.items() if k not in ("_staging", "_ws", "_offsets_cache")})
        fresh._ws = {}
        big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        big.fill_(7)                                              # in flight on main; freed right away -> reusable block
        del big
        a, _ = fresh.score(dev, off)                              # in flight on main
        b, _ = fresh.score_host(host_small, lens[:2])             # first staging allocation, no sync before it
        c, cs = fresh.score_host(host_all, lens)                  # staging regrow while the previous call is in flight
        d, _ = fresh.score_host(host_small, lens[:2])
        torch.cuda.synchronize()
        assert torch.equal(a, ref_l) and torch.equal(b, ref_small) and torch.equal(c, ref_l) and torch.equal(cs, ref_s) and torch.equal(d, ref_small)


def test_bad_offsets_poison_instead_of_reading_out_of_bounds(scorer, golden_crops):
    """Offsets that run past the feature matrix give NaN logits / frame scores for that video, never an out-of-bounds read."""
    crops, offsets = golden_crops
    feat = scorer.features(torch.from_numpy(crops[:6]).cuda())
    bad = torch.tensor([0, 4, 9], dtype=torch.int32, device="cuda")          # second video claims frames 4..8 of a 6-frame batch
    logits, scores = scorer.pool_head(feat, bad)
    good, _ = scorer.pool_head(feat[:4], torch.tensor([0, 4], dtype=torch.int32, device="cuda"))
    assert torch.equal(logits[0], good[0]) and torch.isnan(logits[1]).all() and torch.isnan(scores[4:6]).all()


def test_full_c2_batch_properties(scorer, synth_sd):
    """BASELINE configs[1] size (64 videos x 32 crops): size-independent properties (finite, softmax sums, batch invariance,
    permutation equivariance)."""
    from deepfake_video_detection_b200 import decide, make_offsets
    from oracle import effnet_b0_oracle as O
    V, T = 64, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    # smooth face-crop-like frames built on the device (same recipe as synthetic.synth_crops: low-frequency field
    # per video + mid-frequency detail + noise); uniform noise would push the BN-calibrated trunk far out of range
    lo = torch.randint(40, 216, (V, 1, 7, 7, 3), device="cuda", generator=g) + torch.randint(-12, 13, (V, T, 7, 7, 3), device="cuda", generator=g)
    mid = torch.randint(-40, 41, (V, 1, 28, 28, 3), device="cuda", generator=g) + torch.randint(-10, 11, (V, T, 28, 28, 3), device="cuda", generator=g)
    img = lo.repeat_interleave(32, 2).repeat_interleave(32, 3) + mid.repeat_interleave(8, 2).repeat_interleave(8, 3)
    img = img + torch.randint(-16, 17, (V, T, 224, 224, 3), device="cuda", generator=g)
    crops = img.clamp_(0, 255).to(torch.uint8).view(V * T, 224, 224, 3).contiguous()
    del lo, mid, img
    off = make_offsets([T] * V, "cuda")
    logits, scores = scorer.score(crops, off)
    assert torch.isfinite(logits).all() and torch.isfinite(scores).all()
    assert torch.allclose(scores.view(V, T).sum(1), torch.ones(V, device="cuda"), atol=1e-5)       # softmax over T per video
    # batch invariance: a video scored alone, or in a permuted batch, gives the same bits as inside the big batch
    for v in (0, 17, 63):
        lg1, sc1 = scorer.score(crops[v * T:(v + 1) * T], make_offsets([T], "cuda"))
        assert torch.equal(lg1[0], logits[v]) and torch.equal(sc1, scores[v * T:(v + 1) * T])
    perm = torch.randperm(V, generator=torch.Generator().manual_seed(1))
    crops_p = crops.view(V, T, 224, 224, 3)[perm.cuda()].reshape(V * T, 224, 224, 3).contiguous()
    lg_p, _ = scorer.score(crops_p, off)
    assert torch.equal(lg_p, logits[perm.cuda()])
    # parity at this size: tests/test_gpu_parity_large.py compares all 64 videos of a SEEDED in-distribution batch with the oracle
    # (these device-generated crops are unblurred and drive the calibrated trunk far out of range: |logit| ~ 100)


def test_ragged_extremes(scorer):
    """T = 1 and T = 64 (the reference clamps MAX_FRAMES to 1..64, app.py:2053) in one call; empty batch."""
    from deepfake_video_detection_b200 import make_offsets
    g = torch.Generator(device="cuda").manual_seed(9)
    crops = torch.randint(0, 256, (65, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
    logits, scores = scorer.score(crops, make_offsets([1, 64], "cuda"))
    assert scores[0].item() == 1.0 and abs(scores[1:].sum().item() - 1.0) < 1e-5
    l0, _ = scorer.score(crops[:1], make_offsets([1], "cuda"))
    assert torch.equal(l0[0], logits[0])
    e_l, e_s = scorer.score(crops[:0], make_offsets([], "cuda"))
    assert e_l.shape == (0, 2) and e_s.shape == (0,)
    with pytest.raises(ValueError):
        make_offsets([0, 3], "cuda")


def test_cuda_graph_capture_matches_eager(scorer, golden_crops):
    """The whole step captured into one CUDA graph (small-batch latency path) gives the same bits as eager launches."""
    from deepfake_video_detection_b200 import make_offsets
    crops, offsets = golden_crops
    lens = np.diff(offsets).tolist()
    x = torch.from_numpy(crops).cuda()
    ref_l, ref_s = scorer.score(x, make_offsets(lens, "cuda"))
    g = scorer.capture(lens)
    for _ in range(2):
        lg, sc = g.run(x)
        assert torch.equal(lg, ref_l) and torch.equal(sc, ref_s)


def test_npz_batch_scorer_cli(tmp_path, synth_sd, golden, golden_crops):
    """The reference's on-disk crop format (`data_prepare.py:279-281`) -> prediction CSV (`evaluate.py:469-475`)."""
    import csv
    from deepfake_video_detection_b200 import score_npz
    crops, offsets = golden_crops
    for v in range(4):
        name = f"{'fake' if v % 2 else 'real'}_{v}.npz"
        kw = {"label": np.int64(v % 2)} if v < 2 else {}
        np.savez_compressed(tmp_path / name, faces=crops[offsets[v]:offsets[v + 1]], **kw)
    torch.save(synth_sd, tmp_path / "ckpt.pt")
    out = tmp_path / "preds.csv"
    assert score_npz.main(["--data_dir", str(tmp_path), "--checkpoint", str(tmp_path / "ckpt.pt"), "--out_csv", str(out), "--batch_videos", "3"]) == 0
    rows = list(csv.DictReader(open(out)))
    assert [r["file"] for r in rows] == sorted(f"{'fake' if v % 2 else 'real'}_{v}.npz" for v in range(4))
    probs = torch.softmax(torch.from_numpy(golden["logits"][:4]), 1)[:, 1]
    for r in rows:
        v = int(r["file"].split("_")[1].split(".")[0])
        assert int(r["label"]) == v % 2 and abs(float(r["prob"]) - probs[v].item()) < 1e-2
        assert int(r["pred"]) == int(probs[v].item() >= 0.5)


def test_fused_early_blocks_are_on_the_default_path(scorer, golden_crops):
    """Blocks 2.1.0, 2.1.1 and 2.2.0 run expand 1x1 + depthwise as one kernel (mbconv_fused.cu): 71 - 3 launches per pass."""
    from deepfake_video_detection_b200 import make_offsets
    crops, offsets = golden_crops
    scorer.score(torch.from_numpy(crops[:4]).cuda(), make_offsets([4], "cuda"))
    assert scorer.last_launch_count == 68
