"""Freeze golden vectors of the UNMODIFIED reference `PretrainedBackboneDetector('resnet50')` and
`EnsembleDetector(['efficientnet_b0', 'resnet50'], 'weighted')` (build container only).

    python -m oracle.make_golden_resnet
The reference classes are imported as-is from /root/reference (torchvision provides the resnet trunk, oracle/timm_standin the
efficientnet one).  Writes tests/golden/resnet50_ref_seed0.npz: the data-dependent parts of the synthetic resnet50
checkpoint (BN running statistics, head scale) and the reference outputs."""
import os, sys
import numpy as np, torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "timm_standin"))
sys.path.insert(0, "/root/reference")
from oracle import effnet_b0_oracle as O, resnet50_oracle as R, synth_checkpoint as S   # noqa: E402
from deepfake_video_detection_b200.synthetic import synth_crops                            # noqa: E402
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden", "resnet50_ref_seed0.npz")
ENSEMBLE_WEIGHTS = (0.3, -0.2)


def golden_inputs():
    crops, _ = synth_crops(4242, 4, 4)                             # 4 videos x 4 frames
    return O.prep_u8_hwc(crops).view(4, 4, 3, 224, 224)


def main():
    torch.set_num_threads(8)
    from src.pretrained_detector import EnsembleDetector, PretrainedBackboneDetector    # the unmodified reference classes
    calib, _ = synth_crops(1000, 6, 4)
    sd = R.synth_state_dict(0, calib_frames=O.prep_u8_hwc(calib))
    # head: logit margin of std ~1 over calibration videos, centred
    cal2, _ = synth_crops(2000, 12, 4)
    with torch.no_grad():
        feats = R.trunk_features(sd, O.prep_u8_hwc(cal2)).view(12, 4, -1)
        lg, _ = R.pool_head(sd, feats)
        m = lg[:, 1] - lg[:, 0]
        scale = float(1.0 / (m.std() + 1e-6))
        sd["fc2.weight"] = sd["fc2.weight"] * scale
        lg, _ = R.pool_head(sd, feats)
        m = lg[:, 1] - lg[:, 0]
        sd["fc2.bias"] = sd["fc2.bias"] + torch.tensor([0.5, -0.5]) * m.median()
    blob = {k: v.numpy() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    blob["__fc2_scale__"] = np.float64(scale)
    blob["fc2.bias"] = sd["fc2.bias"].numpy()
    # the frozen pieces must reproduce the state_dict exactly
    sd2 = R.synth_state_dict(0, frozen=blob)
    bad = [k for k in sd if not k.endswith('num_batches_tracked') and not torch.equal(sd[k], sd2[k])]
    assert not bad, bad[:5]
    sd = sd2                                                       # what tests rebuild from the fixture

    x = golden_inputs()
    ref = PretrainedBackboneDetector("resnet50", pretrained=False, num_classes=2, dropout_rate=0.5, use_temporal_attention=True)
    print("strict load:", ref.load_state_dict(sd, strict=True))
    ref.eval()
    ens = EnsembleDetector(["efficientnet_b0", "resnet50"], pretrained=False, ensemble_method="weighted")
    esd = {"weights": torch.tensor(ENSEMBLE_WEIGHTS)}
    esd.update({"models.0." + k: v for k, v in S.make_checkpoint(0).items()})
    esd.update({"models.1." + k: v for k, v in sd.items()})
    print("ensemble strict load:", ens.load_state_dict(esd, strict=False))
    ens.eval()
    with torch.no_grad():
        lg, fs = ref(x)
        elg, efs = ens(x)
        f16 = R.trunk_features(sd, x.view(16, 3, 224, 224), torch.float16).view(4, 4, -1)
        lg16, _ = R.pool_head(sd, f16)
    blob.update(logits=lg.numpy(), frame_scores=fs.numpy(), ensemble_logits=elg.numpy(), ensemble_frame_scores=efs.numpy())
    np.savez_compressed(GOLDEN, **blob)
    print("resnet50 logits\\n", lg, "\\nframe scores[0]", fs[0], "\\nensemble logits\\n", elg)
    print("fp16-storage emulation: max |dlogit|", (lg16 - lg).abs().max().item(), " file", os.path.getsize(GOLDEN) // 1024, "KB")


if __name__ == "__main__":
    main()
