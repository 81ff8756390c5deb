"""Freeze golden vectors of the UNMODIFIED reference `src/models.ViTFeatureExtractor` (build container only).

    python -m oracle.make_golden_vit
The reference class is imported as-is from /root/reference over `oracle/timm_standin` (timm is not installed), loaded
strictly with the seeded synthetic weights and run on seeded images.  Writes tests/golden/vit_ref_seed0.npz
(features only; weights and inputs are regenerated from the seed in tests)."""
import os, sys
import numpy as np, torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "timm_standin"))
sys.path.insert(0, "/root/reference")
from oracle import vit_oracle as V   # noqa: E402
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden", "vit_ref_seed0.npz")
N_IMAGES = 4


def main():
    torch.set_num_threads(8)
    from src.models import ViTFeatureExtractor                     # the unmodified reference class
    m = ViTFeatureExtractor("vit_base_patch16_224", pretrained=False).eval()
    print("strict load:", m.load_state_dict(V.synth_state_dict(0), strict=True), "out_dim", m.out_dim)
    x = V.synth_images(0, N_IMAGES)
    with torch.no_grad():
        f = m(x)
    # DeepfakeModel (ViT + SimpleGCN + classifier, models.py:199-291) on 2 videos x 4 nodes, chain graph as app.py:2245-2250
    from src.models import DeepfakeModel
    from src.utils import normalize_adjacency
    dm = DeepfakeModel().eval()
    print("DeepfakeModel strict load:", dm.load_state_dict(V.synth_deepfake_state_dict(0), strict=True))
    A = np.zeros((4, 4), np.float32)
    for i in range(3):
        A[i, i + 1] = A[i + 1, i] = 1.0
    A_norm = torch.from_numpy(normalize_adjacency(A)).float().unsqueeze(0).repeat(2, 1, 1)
    assert torch.equal(A_norm[0], V.chain_adjacency(4))
    imgs = V.synth_images(2, 8).view(2, 4, 3, 224, 224)
    with torch.no_grad():
        dl = dm(imgs, A_norm)
    print("DeepfakeModel logits", dl)
    np.savez_compressed(GOLDEN, features=f.numpy(), deepfake_logits=dl.numpy())
    print("features", f.shape, "abs-mean", f.abs().mean().item(), "max", f.abs().max().item(),
          "row-to-row diff", (f[0] - f[1]).abs().mean().item())


if __name__ == "__main__":
    main()
