"""`resnet50` ensemble member (src/pretrained_detector.py:38-41, :146-218): oracle pins (CPU), CUDA path
vs goldens frozen from the unmodified reference classes (GPU)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "resnet50_ref_seed0.npz")


def _sd():
    from oracle import resnet50_oracle as R
    return R.synth_state_dict(0, frozen=np.load(GOLDEN))


def _inputs():
    from deepfake_video_detection_b200.synthetic import synth_crops
    from oracle import effnet_b0_oracle as O
    crops, _ = synth_crops(4242, 4, 4)                              # as oracle/make_golden_resnet.py::golden_inputs
    return O.prep_u8_hwc(crops).view(4, 4, 3, 224, 224)


def test_resnet50_oracle_matches_reference_goldens_and_schema():
    from deepfake_video_detection_b200 import EnsembleDetector, PretrainedBackboneDetector
    from oracle import resnet50_oracle as R
    g, sd, x = np.load(GOLDEN), _sd(), _inputs()
    with torch.no_grad():
        feats = R.trunk_features(sd, x.view(16, 3, 224, 224)).view(4, 4, -1)
        lg, fs = R.pool_head(sd, feats)
    assert np.abs(lg.numpy() - g["logits"]).max() < 1e-4 and np.abs(fs.numpy() - g["frame_scores"]).max() < 1e-5
    assert np.ptp(g["logits"][:, 1] - g["logits"][:, 0]) > 0.5      # video-dependent: not vacuous
    m = PretrainedBackboneDetector("resnet50", pretrained=False)
    assert m.feature_dim == 2048 and set(m.state_dict()) == set(sd)  # torchvision trunk keys + pool/head keys (326)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)
    e = EnsembleDetector(["efficientnet_b0", "resnet50"], pretrained=False, ensemble_method="weighted")
    assert len(e.models) == 2 and e.models[1].backbone_name == "resnet50" and e.weights.shape == (2,)
    with pytest.raises(ValueError):
        PretrainedBackboneDetector("resnet18", pretrained=False)


@pytest.mark.gpu
def test_resnet50_member_and_ensemble_cuda_match_reference(synth_sd):
    from deepfake_video_detection_b200 import EnsembleDetector, PretrainedBackboneDetector, decide
    g, sd, x = np.load(GOLDEN), _sd(), _inputs()
    m = PretrainedBackboneDetector("resnet50", pretrained=False).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    with torch.no_grad():
        lg, fs = m(x.cuda())
        one, _ = m(x[2:3, :3].cuda())                                # a ragged B = 1 call (T = 3) goes through the same kernels
    err = np.abs(lg.cpu().numpy() - g["logits"]).max()
    print(f"resnet50 member: max |dlogit| vs reference goldens {err:.3e}")
    assert err <= 2e-2 and np.abs(fs.cpu().numpy() - g["frame_scores"]).max() <= 5e-3
    assert [d["is_fake"] for d in decide(lg)] == [d["is_fake"] for d in decide(torch.from_numpy(g["logits"]))]
    assert torch.isfinite(one).all()
    e = EnsembleDetector(["efficientnet_b0", "resnet50"], pretrained=False, ensemble_method="weighted")
    esd = {"weights": torch.tensor([0.3, -0.2])}
    esd.update({"models.0." + k: v for k, v in synth_sd.items()})
    esd.update({"models.1." + k: v for k, v in sd.items()})
    e.load_state_dict(esd, strict=False)
    e = e.eval().cuda()
    with torch.no_grad():
        elg, efs = e(x.cuda())
    eerr = np.abs(elg.cpu().numpy() - g["ensemble_logits"]).max()
    print(f"ensemble (weighted): max |dlogit| vs reference goldens {eerr:.3e}")
    assert eerr <= 2e-2 and np.abs(efs.cpu().numpy() - g["ensemble_frame_scores"]).max() <= 5e-3


@pytest.mark.gpu
def test_resnet50_second_call_reuses_the_zero_halo():
    """The stride-1 3x3 convs run as implicit GEMMs over a zero-haloed map whose halo is zeroed once per geometry: a second call
    on the same workspace must give the same bits, and the goldens from the unmodified reference must hold."""
    from deepfake_video_detection_b200 import PretrainedBackboneDetector
    g, sd, x = np.load(GOLDEN), _sd(), _inputs()
    m = PretrainedBackboneDetector("resnet50", pretrained=False).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    with torch.no_grad():
        lg, fs = m(x.cuda())
        lg2, _ = m(x.cuda())
    err = np.abs(lg.cpu().numpy() - g["logits"]).max()
    assert err <= 2e-2 and torch.equal(lg, lg2)
