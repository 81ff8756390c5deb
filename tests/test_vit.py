"""ViT-B/16 frame encoder (BASELINE config 5, src/models.py:88-107): oracle pins (CPU), CUDA path vs oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "vit_ref_seed0.npz")


def _hf_state_dict(sd):
    """timm -> transformers.ViTModel key map (independent implementation; qkv split into query/key/value)."""
    out = {"embeddings.cls_token": sd["vit.cls_token"], "embeddings.position_embeddings": sd["vit.pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["vit.patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["vit.patch_embed.proj.bias"],
           "layernorm.weight": sd["vit.norm.weight"], "layernorm.bias": sd["vit.norm.bias"]}
    for i in range(12):
        p, q = f"vit.blocks.{i}.", f"encoder.layer.{i}."
        for j, n in enumerate(("query", "key", "value")):
            out[q + f"attention.attention.{n}.weight"] = sd[p + "attn.qkv.weight"][j * 768:(j + 1) * 768]
            out[q + f"attention.attention.{n}.bias"] = sd[p + "attn.qkv.bias"][j * 768:(j + 1) * 768]
        for a, b in (("attention.output.dense", "attn.proj"), ("layernorm_before", "norm1"), ("layernorm_after", "norm2"),
                     ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2")):
            out[q + a + ".weight"], out[q + a + ".bias"] = sd[p + b + ".weight"], sd[p + b + ".bias"]
    return out


def test_vit_oracle_matches_reference_goldens_and_transformers():
    from oracle import vit_oracle as V
    sd, x = V.synth_state_dict(0), V.synth_images(0, 4)
    assert len(sd) == 4 + 12 * 12 + 2                               # timm vit_base_patch16_224, num_classes=0
    with torch.no_grad():
        f = V.vit_features(sd, x)
    g = np.load(GOLDEN)["features"]                                  # frozen from the unmodified reference class
    assert np.abs(f.numpy() - g).max() < 1e-4
    assert np.abs(g[0] - g[1]).mean() > 0.1                          # image-dependent: not vacuous
    from transformers import ViTConfig, ViTModel
    cfg = ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                    layer_norm_eps=1e-6, image_size=224, patch_size=16, qkv_bias=True)
    m = ViTModel(cfg, add_pooling_layer=False).eval()
    m.load_state_dict(_hf_state_dict(sd), strict=True)
    with torch.no_grad():
        h = m(pixel_values=x[:2]).last_hidden_state[:, 0]
    assert (h - f[:2]).abs().max().item() < 1e-4


def test_vit_dropin_contract():
    from deepfake_video_detection_b200.vit_model import ViTFeatureExtractor
    from oracle import vit_oracle as V
    sd = V.synth_state_dict(0)
    m = ViTFeatureExtractor("vit_base_patch16_224", pretrained=False)
    assert m.out_dim == 768 and set(m.state_dict()) == set(sd)       # the reference's schema (models.py:93)
    m.load_state_dict(sd, strict=True)
    x = V.synth_images(1, 1)
    m.train()
    with torch.no_grad():
        f = m(x)                                                     # eager training-mode graph = same function
        assert (f - V.vit_features(sd, x)).abs().max().item() < 1e-4
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16", 2e-2), ("bf16", 1.5e-1)])
def test_vit_cuda_matches_oracle(prec, tol):
    """16-bit GEMM operands, fp32 accumulate / residual stream / LayerNorm / softmax: north_star's 2e-2 on O(1) outputs."""
    from deepfake_video_detection_b200.vit_model import ViTFeatureExtractor
    from oracle import vit_oracle as V
    sd, x = V.synth_state_dict(0), V.synth_images(0, 4)
    m = ViTFeatureExtractor(precision=prec).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        f = m(x.cuda()).cpu()
        ref = V.vit_features(sd, x)
    g = torch.from_numpy(np.load(GOLDEN)["features"])
    err, err_g = (f - ref).abs().max().item(), (f - g).abs().max().item()
    print(f"vit {prec}: max |d| vs oracle {err:.3e}, vs reference goldens {err_g:.3e}, rel {((f - ref).norm() / ref.norm()).item():.3e}")
    assert err <= tol and err_g <= tol


@pytest.mark.gpu
def test_vit_cuda_batch_invariance_and_ragged_batches():
    from deepfake_video_detection_b200.vit_model import ViTFeatureExtractor
    from oracle import vit_oracle as V
    m = ViTFeatureExtractor().eval()
    m.load_state_dict(V.synth_state_dict(0), strict=True)
    x = V.synth_images(3, 7).cuda()
    with torch.no_grad():
        full = m(x)
        one = torch.cat([m(x[i:i + 1]) for i in range(7)])
        perm = torch.randperm(7, generator=torch.Generator().manual_seed(0))
        shuffled = m(x[perm.cuda()])
    assert torch.equal(full, one)                                    # bit-identical whatever the batch composition
    assert torch.equal(full[perm.cuda()], shuffled)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 192, 192, device="cuda"))
