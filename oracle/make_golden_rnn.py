"""Freeze golden vectors of the UNMODIFIED reference `src/RNNModel.LogicRNNLSTM` (build container only).

    python -m oracle.make_golden_rnn
Weights: seeded draws (default nn.Linear init scaled x2 so gates leave the linear regime); inputs: seeded
feature-like tensors.  Writes tests/golden/rnn_ref_seed0.npz (inputs are regenerated from the seed in tests)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rnn_ref_seed0.npz")


def make_inputs(seed=0, B=6, T=16, IN=1280):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, IN, generator=g).abs() * 0.5            # pooled SiLU features are mostly positive, O(0.3)
    lengths = torch.tensor([16, 3, 9, 16, 1, 12])[:B]
    return x, lengths


def make_state_dict(seed=0, IN=1280, H=512, L=2):
    from src.RNNModel import LogicRNNLSTM
    torch.manual_seed(seed)
    m = LogicRNNLSTM(IN, H, L, dropout=0.5).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(2.0)
        # calibrate the last Linear so the pre-sigmoid logits spread O(1) over the golden inputs (otherwise every
        # probability is ~0.52 and a parity check on it would be vacuous)
        x, _ = make_inputs(seed)
        feats = []
        hook = m.classifier[2].register_forward_hook(lambda mod, i, o: feats.append(o))
        m(x); hook.remove()
        z = torch.nn.functional.linear(feats[0], m.classifier[3].weight)
        m.classifier[3].weight.mul_(1.5 / (z.std() + 1e-9))
        m.classifier[3].bias.copy_(-(z * (1.5 / (z.std() + 1e-9))).mean().reshape(1))
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    m, sd = make_state_dict()
    x, lengths = make_inputs()
    with torch.no_grad():
        p_plain = m(x)
        p_len = m(x, lengths)
    np.savez_compressed(GOLDEN, prob=p_plain.numpy(), prob_lengths=p_len.numpy())
    print("prob", p_plain.flatten().numpy(), "\nwith lengths (sorted order)", p_len.flatten().numpy())


if __name__ == "__main__":
    main()
