"""CPU: the oracle restatement against (1) the golden vectors frozen from the unmodified reference class
(oracle/make_golden.py) and (2) torchvision's independent EfficientNet-B0 (key map SURVEY.md App. C)."""
import hashlib

import numpy as np
import torch

from oracle import effnet_b0_oracle as O
from oracle import synth_checkpoint as S


def test_checkpoint_schema_matches_reference(synth_sd):
    keys = [k for k, _ in S.schema()]
    assert list(synth_sd.keys()) == keys and len(keys) == 366
    n_params = sum(v.numel() for k, v in synth_sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == 4_418_047
    assert sum(v.numel() * v.element_size() for v in synth_sd.values()) == 17_840_644


def test_synthetic_crops_are_reproducible(golden, golden_crops):
    crops, offsets = golden_crops
    assert np.array_equal(offsets, golden["offsets"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(crops.tobytes()).digest(), np.uint8), golden["crops_sha256"])


def test_prep_matches_reference(golden, golden_crops):
    crops, _ = golden_crops
    assert np.array_equal(O.prep_u8_hwc(crops[:1]).numpy(), golden["prepped_frame0"])     # bit-exact fp32


def test_oracle_matches_reference_goldens(synth_sd, golden, golden_crops):
    crops, offsets = golden_crops
    torch.set_num_threads(8)
    logits, scores = O.score_ragged(synth_sd, crops, offsets)
    assert np.abs(logits.numpy() - golden["logits"]).max() < 1e-4        # north_star fp32 bar
    assert np.abs(scores.numpy() - golden["frame_scores"]).max() < 1e-5
    with torch.no_grad():
        feats = O.trunk_features(synth_sd, O.prep_u8_hwc(crops[:8]))
    assert np.abs(feats.numpy() - golden["features"][:8]).max() < 1e-4
    # activations are not vacuous
    assert np.abs(golden["features"]).mean() > 0.05 and np.abs(golden["logits"]).max() > 0.5


def test_oracle_batched_and_meanpool(synth_sd, golden, golden_crops):
    crops, offsets = golden_crops
    x = O.prep_u8_hwc(crops[: offsets[4]]).view(4, 8, 3, 224, 224)
    lg, fs = O.detector_forward(synth_sd, x, True)
    assert np.abs(lg.numpy() - golden["batched_logits"]).max() < 1e-4
    assert np.abs(fs.numpy() - golden["batched_frame_scores"]).max() < 1e-5
    lg, fs = O.detector_forward(synth_sd, x, False)
    assert np.abs(lg.numpy() - golden["meanpool_logits"]).max() < 1e-4
    assert np.allclose(fs.numpy(), golden["meanpool_frame_scores"])


def test_trunk_matches_torchvision(synth_sd):
    """Independent implementation of the same published architecture, weights mapped per SURVEY.md App. C."""
    from torchvision.models import efficientnet_b0
    tv = efficientnet_b0(weights=None).eval()
    m = {"backbone.0.": "features.0.0.", "backbone.1.": "features.0.1.", "backbone.3.": "features.8.0.", "backbone.4.": "features.8.1."}
    sub_ds = {"conv_dw": "block.0.0", "bn1": "block.0.1", "se.conv_reduce": "block.1.fc1", "se.conv_expand": "block.1.fc2",
              "conv_pw": "block.2.0", "bn2": "block.2.1"}
    sub_ir = {"conv_pw": "block.0.0", "bn1": "block.0.1", "conv_dw": "block.1.0", "bn2": "block.1.1", "se.conv_reduce": "block.2.fc1",
              "se.conv_expand": "block.2.fc2", "conv_pwl": "block.3.0", "bn3": "block.3.1"}
    new = {}
    for k, v in synth_sd.items():
        if k.startswith(("temporal_attention", "fc1", "fc2")):
            continue
        if k.startswith("backbone.2."):
            _, _, s, b, rest = k.split(".", 4)
            sub = sub_ds if s == "0" else sub_ir
            name = next(n for n in sorted(sub, key=len, reverse=True) if rest.startswith(n + "."))
            new[f"features.{int(s) + 1}.{b}.{sub[name]}.{rest[len(name) + 1:]}"] = v
        else:
            p = next(p for p in m if k.startswith(p))
            new[m[p] + k[len(p):]] = v
    missing = tv.load_state_dict(new, strict=False)
    assert all(k.startswith("classifier") for k in missing.missing_keys) and not missing.unexpected_keys
    crops, _ = S.synth_crops(3, 1, 4)
    x = O.prep_u8_hwc(crops)
    with torch.no_grad():
        ref = torch.flatten(tv.avgpool(tv.features(x)), 1)
        ours = O.trunk_features(synth_sd, x)
    assert (ref - ours).abs().max().item() < 1e-4


def test_decision_rule():
    lg = torch.tensor([[0.0, 0.0], [2.0, -1.0], [-1.0, 2.0], [0.0, 0.3]])
    d = O.decide(lg)
    assert [x["is_fake"] for x in d] == [True, False, True, True]       # prob_fake >= 0.5 (app.py:2110)
    assert [x["abstained"] for x in d] == [True, False, False, True]     # confidence < 0.60 (app.py:2193)
    assert O.decide(lg, threshold=0.99)[1]["threshold"] == 0.5           # extreme-threshold guard (app.py:2106-2109)
    assert O.decide(lg, threshold=0.99, allow_extreme_threshold=True)[2]["is_fake"] is False
    from deepfake_video_detection_b200 import decide
    for a, b in zip(decide(lg), d):
        assert a["is_fake"] == b["is_fake"] and a["abstained"] == b["abstained"] and abs(a["prob_fake"] - b["prob_fake"]) < 1e-7


def test_fp16_storage_meets_logit_bar_and_bf16_does_not(synth_sd):
    """Why fp16 is the default storage type: emulate both rounding schemes on CPU."""
    import bf16_emulation as E
    crops, _ = S.synth_crops(11, 4, 4)
    x = O.prep_u8_hwc(crops)
    with torch.no_grad():
        f32 = O.trunk_features(synth_sd, x)
        l32, _ = O.attention_pool_head(synth_sd, f32.view(4, 4, -1))
        orig = E.bf
        try:
            E.bf = lambda t: t.to(torch.float16).float()
            l16, _ = O.attention_pool_head(synth_sd, E.trunk_features_bf16(synth_sd, x).view(4, 4, -1))
        finally:
            E.bf = orig
        lbf, _ = O.attention_pool_head(synth_sd, E.trunk_features_bf16(synth_sd, x).view(4, 4, -1))
    assert (l16 - l32).abs().max().item() < 2e-2
    assert (lbf - l32).abs().max().item() > (l16 - l32).abs().max().item()


def test_batched_oracle_scorer_matches_per_video_calls(synth_sd, golden, golden_crops):
    """`score_ragged_batched` (large parity samples) against `score_ragged` (one reference B=1 call per video) and the goldens."""
    crops, offsets = golden_crops
    lg, fs, feats = O.score_ragged_batched(synth_sd, crops, offsets, chunk=7)
    lg1, fs1 = O.score_ragged(synth_sd, crops, offsets)
    assert (lg - lg1).abs().max().item() < 1e-5 and (fs - fs1).abs().max().item() < 1e-6
    assert np.abs(lg.numpy() - golden["logits"]).max() < 1e-4
    assert np.abs(feats.numpy() - golden["features"]).max() < 1e-4
