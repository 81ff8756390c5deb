"""LogicRNNLSTM temporal head (BASELINE config 3): oracle vs reference goldens (CPU), CUDA path vs oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights(seed=0, IN=1280, H=512, L=2):
    """Same draws as oracle/make_golden_rnn.py::make_state_dict, without importing the reference: nn.Linear default
    init in module construction order, x2, last Linear calibrated."""
    from deepfake_video_detection_b200.rnn_model import LogicRNNLSTM
    from oracle import rnn_oracle as R
    from oracle.make_golden_rnn import make_inputs
    torch.manual_seed(seed)
    m = LogicRNNLSTM(IN, H, L, dropout=0.5).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(2.0)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        x, _ = make_inputs(seed)
        # replicate the calibration of the last Linear with the oracle's own forward
        B, T, _ = x.shape
        h, c, outs = torch.zeros(B, H), torch.zeros(B, H), []
        for t in range(T):
            ht, ct = h, c
            for i in range(L):
                ht, ct = R.logic_cell(sd, f"logic_cells.{i}", x[:, t, :] if i == 0 else ht, ht, ct)
            outs.append(ht); h, c = ht, ct
        outs = torch.stack(outs, 1)
        a = torch.nn.functional.linear(torch.tanh(torch.nn.functional.linear(outs, sd["attention.0.weight"], sd["attention.0.bias"])),
                                       sd["attention.2.weight"], sd["attention.2.bias"])
        ctx = torch.sum(torch.softmax(a, dim=1) * outs, dim=1)
        f = torch.relu(torch.nn.functional.linear(ctx, sd["classifier.0.weight"], sd["classifier.0.bias"]))
        z = torch.nn.functional.linear(f, sd["classifier.3.weight"])
        sd["classifier.3.weight"] = sd["classifier.3.weight"] * (1.5 / (z.std() + 1e-9))
        sd["classifier.3.bias"] = -(z * (1.5 / (z.std() + 1e-9))).mean().reshape(1)
    m.load_state_dict(sd)
    return m, sd


def test_rnn_schema_and_oracle_match_reference_goldens():
    from oracle import rnn_oracle as R
    from oracle.make_golden_rnn import make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "rnn_ref_seed0.npz"))
    m, sd = _weights()
    assert len(sd) == 2 * 14 + 8                                   # 7 Linears per LogicCell x 2 layers + attention + classifier
    x, lengths = make_inputs()
    with torch.no_grad():
        p = R.rnn_forward(sd, x)
        pl = R.rnn_forward(sd, x, lengths)
    assert np.abs(p.numpy() - g["prob"]).max() < 1e-5
    assert np.abs(pl.numpy() - g["prob_lengths"]).max() < 1e-5
    assert g["prob"].std() > 0.1                                    # not vacuous
    m.train()
    with torch.no_grad():
        m.dropout.p = 0.0; m.classifier[2].p = 0.0
        assert (m(x, lengths) - pl).abs().max().item() < 1e-5       # eager training-mode graph = same function
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16", 5e-3), ("bf16", 4e-2)])
def test_rnn_cuda_matches_oracle(prec, tol):
    from oracle import rnn_oracle as R
    from oracle.make_golden_rnn import make_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "rnn_ref_seed0.npz"))
    m, sd = _weights()
    m.precision = prec
    m = m.cuda().eval()
    x, lengths = make_inputs()
    with torch.no_grad():
        p = m(x.cuda()).cpu()
        pl = m(x.cuda(), lengths.cuda()).cpu()
    assert p.shape == (6, 1)
    assert np.abs(p.numpy() - g["prob"]).max() < tol, np.abs(p.numpy() - g["prob"]).max()
    assert np.abs(pl.numpy() - g["prob_lengths"]).max() < tol
    assert torch.equal(m.predict(x.cuda()).cpu(), (torch.from_numpy(g["prob"]) >= 0.5).float())


@pytest.mark.gpu
def test_rnn_config3_shape_runs_on_effnet_features():
    """BASELINE config 3 wiring (evaluate.py:143-192): trunk features (B,N,1280) -> LogicRNNLSTM(1280,512,2) -> (B,1)."""
    from deepfake_video_detection_b200 import FrameScorer
    from deepfake_video_detection_b200.synthetic import load_checkpoint, synth_crops
    from oracle import rnn_oracle as R
    m, sd = _weights()
    m = m.cuda().eval()
    crops, _ = synth_crops(21, 3, 16)
    feats = FrameScorer(load_checkpoint(0), "fp16", "cuda").features(torch.from_numpy(crops).cuda()).view(3, 16, 1280)
    with torch.no_grad():
        p = m(feats).cpu()
        ref = R.rnn_forward(sd, feats.cpu())
    assert (p - ref).abs().max().item() < 5e-3
