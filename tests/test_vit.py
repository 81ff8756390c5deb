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


def test_deepfake_model_oracle_and_dropin_contract():
    """DeepfakeModel = ViT + SimpleGCN + classifier (models.py:199-291): restatement vs goldens of the unmodified class."""
    from deepfake_video_detection_b200.vit_model import DeepfakeModel
    from oracle import vit_oracle as V
    sd = V.synth_deepfake_state_dict(0)
    imgs = V.synth_images(2, 8).view(2, 4, 3, 224, 224)
    A = V.chain_adjacency(4).unsqueeze(0).repeat(2, 1, 1)
    with torch.no_grad():
        lg = V.deepfake_forward(sd, imgs, A)
    g = np.load(GOLDEN)["deepfake_logits"]
    assert np.abs(lg.numpy() - g).max() < 1e-4
    m = DeepfakeModel()
    assert set(m.state_dict()) == set(sd)                          # vit.vit.*, gcn.*, classifier.*
    m.load_state_dict(sd, strict=True)
    m.train(); m.gcn.dropout.p = 0.0; m.classifier[2].p = 0.0
    with torch.no_grad():
        assert (m(imgs[:1, :2], A[:1, :2, :2]) - V.deepfake_forward(sd, imgs[:1, :2], A[:1, :2, :2])).abs().max().item() < 1e-4
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(imgs, A)
    with pytest.raises(ValueError):
        DeepfakeModel(backbone="clip")


@pytest.mark.gpu
def test_deepfake_model_cuda_matches_oracle():
    from deepfake_video_detection_b200.vit_model import DeepfakeModel
    from oracle import vit_oracle as V
    sd = V.synth_deepfake_state_dict(0)
    imgs = V.synth_images(2, 8).view(2, 4, 3, 224, 224)
    A = V.chain_adjacency(4).unsqueeze(0).repeat(2, 1, 1)
    m = DeepfakeModel().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = m(imgs.cuda(), A.cuda()).cpu()
        ref = V.deepfake_forward(sd, imgs, A)
        # the fused head alone, on exact fp32 features: tight tolerance (only the summation order differs)
        vit_sd = {k[4:]: v for k, v in sd.items() if k.startswith("vit.")}
        feats = V.vit_features(vit_sd, imgs.reshape(8, 3, 224, 224)).view(2, 4, -1)
    g = torch.from_numpy(np.load(GOLDEN)["deepfake_logits"])
    print(f"deepfake model: max |dlogit| vs oracle {(out - ref).abs().max().item():.3e}, vs reference goldens {(out - g).abs().max().item():.3e}")
    assert (out - ref).abs().max().item() <= 2e-2 and (out - g).abs().max().item() <= 2e-2
    assert (out.argmax(1) == g.argmax(1)).all()
    import ctypes as C
    from deepfake_video_detection_b200 import _lib
    from deepfake_video_detection_b200.engine import _stream_ptr
    lib = _lib.load()
    for n_nodes in (4, 1, 3):                                       # ragged node counts incl. a single-frame graph
        f = feats[:, :n_nodes].contiguous().cuda(); a = V.chain_adjacency(n_nodes).unsqueeze(0).repeat(2, 1, 1).cuda()
        o = torch.empty((2, 2), device="cuda")
        assert lib.dfd_gcn_head(m._pack_head(f.device), f.data_ptr(), a.data_ptr(), 2, n_nodes, o.data_ptr(), _stream_ptr(f.device)) == 0
        assert (o.cpu() - V.gcn_head(sd, f.cpu(), a.cpu())).abs().max().item() < 1e-4
    assert lib.dfd_gcn_head(m._pack_head(f.device), f.data_ptr(), a.data_ptr(), 2, 65, o.data_ptr(), _stream_ptr(f.device)) != 0   # nodes > 64


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16", 2.5e-3), ("bf16", 2e-2)])
@pytest.mark.parametrize("images", [1, 3, 37, 160])
def test_vit_attention_tcgen05_kernel(prec, tol, images):
    """The tcgen05 / TMEM attention kernel (csrc/vit_attn_tc.cu) on random qkv against fp32 softmax(Q K^T / 8) V of the same
    16-bit-rounded operands (tolerance = rounding of P and of the output to the storage type: the mma.sync kernel this one
    replaced measured the same 1.9e-3 / 1.6e-2 on B200).  37 images = 888 tiles: several per CTA (both TMEM buffers, both stages,
    the P buffer reused), the last image's tiles read past the end of the tensor (zero fill); 160 images: > 12 tiles per CTA."""
    from deepfake_video_detection_b200 import _lib
    from deepfake_video_detection_b200.engine import PRECISIONS, TORCH_DTYPE, _stream_ptr
    lib = _lib.load()
    g = torch.Generator().manual_seed(images)
    qkv = (torch.randn(images * 197, 2304, generator=g) * 1.5).to(TORCH_DTYPE[prec]).cuda()
    o = torch.full((images * 197, 768), float("nan"), dtype=TORCH_DTYPE[prec], device="cuda")
    rc = lib.dfd_k_vit_attention(qkv.data_ptr(), o.data_ptr(), images, PRECISIONS[prec], _stream_ptr(qkv.device))
    assert rc == 0, lib.dfd_vit_last_error()
    torch.cuda.synchronize()
    out = o.float().cpu()
    x = qkv.float().cpu().view(images, 197, 3, 12, 64).permute(2, 0, 3, 1, 4)          # (3, B, heads, tokens, d)
    ref = torch.softmax(x[0] @ x[1].transpose(-1, -2) * 0.125, dim=-1) @ x[2]          # (B, heads, tokens, d)
    ref = ref.permute(0, 2, 1, 3).reshape(images * 197, 768)
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item()
    print(f"vit attention {prec} images {images}: max |err| {err:.2e}")
    assert err <= tol, err
