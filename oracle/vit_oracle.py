"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the ViT-B/16 frame encoder of BASELINE config 5.

Restates, in plain fp32 torch ops driven by the state_dict, what `src/models.py:88-107`
(`ViTFeatureExtractor.forward` -> `timm.create_model('vit_base_patch16_224', num_classes=0)`) computes.  timm is an
un-vendored dependency (`requirements.txt:12`, `timm>=0.9.0`, no lock file): the published architecture is restated
here — 16x16/16 patch conv, CLS token + learned position embedding, 12 pre-norm blocks (LayerNorm eps 1e-6,
12-head attention with qkv bias, MLP 3072 with exact-erf GELU), final LayerNorm, CLS row out.
Pinned by tests/test_oracle.py against (a) goldens frozen from the UNMODIFIED reference class over
`oracle/timm_standin` (oracle/make_golden_vit.py) and (b) `transformers.ViTModel`, an independent implementation.
Keys are the reference's state_dict names (`vit.` prefix = the attribute at src/models.py:93).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module."""
import torch
import torch.nn.functional as F

DIM, DEPTH, HEADS, PATCH, TOKENS = 768, 12, 12, 16, 197


def synth_state_dict(seed=0):
    """Seeded ViT-B/16 weights with the reference schema (152 tensors).  Larger than timm's trunc_normal(.02) init
    so that attention is not uniform and the residual branches matter (a parity check on near-identity blocks
    would be vacuous)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std: torch.randn(*s, generator=g) * std
    sd = {"vit.cls_token": r(1, 1, DIM, std=0.5), "vit.pos_embed": r(1, TOKENS, DIM, std=0.5),
          "vit.patch_embed.proj.weight": r(DIM, 3, PATCH, PATCH, std=0.03), "vit.patch_embed.proj.bias": r(DIM, std=0.1)}
    for i in range(DEPTH):
        p = f"vit.blocks.{i}."
        sd[p + "norm1.weight"] = 1 + r(DIM, std=0.1); sd[p + "norm1.bias"] = r(DIM, std=0.1)
        sd[p + "attn.qkv.weight"] = r(3 * DIM, DIM, std=0.06); sd[p + "attn.qkv.bias"] = r(3 * DIM, std=0.1)
        sd[p + "attn.proj.weight"] = r(DIM, DIM, std=0.03); sd[p + "attn.proj.bias"] = r(DIM, std=0.05)
        sd[p + "norm2.weight"] = 1 + r(DIM, std=0.1); sd[p + "norm2.bias"] = r(DIM, std=0.1)
        sd[p + "mlp.fc1.weight"] = r(4 * DIM, DIM, std=0.04); sd[p + "mlp.fc1.bias"] = r(4 * DIM, std=0.1)
        sd[p + "mlp.fc2.weight"] = r(DIM, 4 * DIM, std=0.02); sd[p + "mlp.fc2.bias"] = r(DIM, std=0.05)
    sd["vit.norm.weight"] = 1 + r(DIM, std=0.1); sd["vit.norm.bias"] = r(DIM, std=0.1)
    return sd


def synth_images(seed, n):
    """Normalised-crop-like inputs (B,3,224,224): smooth low-frequency content + noise, O(1) values."""
    g = torch.Generator().manual_seed(1000 + seed)
    low = F.interpolate(torch.randn(n, 3, 14, 14, generator=g), size=224, mode="bilinear", align_corners=False)
    return (low + 0.3 * torch.randn(n, 3, 224, 224, generator=g)).contiguous()


def vit_features(sd, x):
    """x (B,3,224,224) fp32 -> (B,768) fp32: the CLS row after the final norm (src/models.py:105-107)."""
    B = x.shape[0]
    t = F.conv2d(x, sd["vit.patch_embed.proj.weight"], sd["vit.patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["vit.cls_token"].expand(B, -1, -1), t), dim=1) + sd["vit.pos_embed"]
    hd = DIM // HEADS
    for i in range(DEPTH):
        p = f"vit.blocks.{i}."
        h = F.layer_norm(t, (DIM,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, TOKENS, 3, HEADS, hd).permute(2, 0, 3, 1, 4)
        a = ((qkv[0] * hd ** -0.5) @ qkv[1].transpose(-2, -1)).softmax(-1) @ qkv[2]
        t = t + F.linear(a.transpose(1, 2).reshape(B, TOKENS, DIM), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F.layer_norm(t, (DIM,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        t = t + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return F.layer_norm(t[:, 0], (DIM,), sd["vit.norm.weight"], sd["vit.norm.bias"], 1e-6)


# ---- DeepfakeModel = ViT frame encoder + SimpleGCN + classifier (src/models.py:177-291; caller app.py:2225-2255) ----------
def normalize_adjacency(A):
    """src/utils.py:95-104: D^-1/2 (A + I) D^-1/2 (numpy, fp32)."""
    import numpy as np
    A = A.copy().astype(np.float32)
    A = A + np.eye(A.shape[0], dtype=A.dtype)
    d = np.power(np.sum(A, axis=1), -0.5)
    d[np.isinf(d)] = 0.0
    return np.diag(d) @ A @ np.diag(d)


def chain_adjacency(n):
    """app.py:2245-2250: consecutive frames are linked."""
    import numpy as np
    A = np.zeros((n, n), np.float32)
    for i in range(n - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0
    return torch.from_numpy(normalize_adjacency(A)).float()


def synth_deepfake_state_dict(seed=0):
    """DeepfakeModel schema: the ViT under `vit.vit.*` (models.py:223 wraps ViTFeatureExtractor) + gcn + classifier."""
    sd = {"vit." + k: v for k, v in synth_state_dict(seed).items()}
    g = torch.Generator().manual_seed(100 + seed)
    r = lambda *s, std: torch.randn(*s, generator=g) * std
    sd.update({"gcn.fc1.weight": r(256, DIM, std=0.05), "gcn.fc1.bias": r(256, std=0.1),
               "gcn.fc2.weight": r(128, 256, std=0.08), "gcn.fc2.bias": r(128, std=0.1),
               "classifier.0.weight": r(64, 128, std=0.12), "classifier.0.bias": r(64, std=0.1),
               "classifier.3.weight": r(2, 64, std=0.3), "classifier.3.bias": r(2, std=0.1)})
    return sd


def gcn_head(sd, feats, A_norm):
    """feats (B,N,768), A_norm (B,N,N) -> logits (B,2): models.py:186-197 (SimpleGCN) + :283-291 (mean pool, classifier)."""
    h = torch.bmm(A_norm, feats)
    h = F.relu(F.linear(h, sd["gcn.fc1.weight"], sd["gcn.fc1.bias"]))
    h = F.relu(F.linear(h, sd["gcn.fc2.weight"], sd["gcn.fc2.bias"]))
    p = h.mean(dim=1)
    return F.linear(F.relu(F.linear(p, sd["classifier.0.weight"], sd["classifier.0.bias"])), sd["classifier.3.weight"], sd["classifier.3.bias"])


def deepfake_forward(sd, images, A_norm):
    """images (B,N,3,224,224), A_norm (B,N,N) -> logits (B,2)  (models.py:271-291)."""
    B, N = images.shape[:2]
    vit_sd = {k[4:]: v for k, v in sd.items() if k.startswith("vit.")}
    feats = vit_features(vit_sd, images.reshape(B * N, *images.shape[2:])).view(B, N, -1)
    return gcn_head(sd, feats, A_norm)
