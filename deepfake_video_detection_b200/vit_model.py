"""Drop-in for the reference's `src/models.ViTFeatureExtractor` (models.py:88-107): ViT-B/16 frame encoder.

Same constructor arguments, `.vit` child with timm's `vit_base_patch16_224` parameter names (152-tensor state_dict,
strict loads work), `.out_dim`, and `forward(x[B,3,224,224]) -> [B,768]` (CLS row after the final norm).  In eval
mode the encoder runs in libdfd_b200.so (csrc/vit.cu: tcgen05 GEMMs + fused LayerNorm / attention kernels); training
mode keeps an eager graph for autograd.  No CPU path for inference."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .engine import DEFAULT_PRECISION, PRECISIONS, _stream_ptr

DIM, DEPTH, HEADS, PATCH, TOKENS = 768, 12, 12, 16, 197
MAX_IMAGES_PER_CALL = 512                       # workspace = 2.1 MB per image


class _PatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(3, DIM, PATCH, PATCH)


class _Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.qkv, self.proj = nn.Linear(DIM, 3 * DIM), nn.Linear(DIM, DIM)


class _Mlp(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(DIM, 4 * DIM), nn.Linear(4 * DIM, DIM)


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.norm1, self.attn = nn.LayerNorm(DIM, eps=1e-6), _Attention()
        self.norm2, self.mlp = nn.LayerNorm(DIM, eps=1e-6), _Mlp()


class VisionTransformerB16(nn.Module):
    """Parameter container with timm's names (`vit_base_patch16_224`, num_classes=0) + the eager training forward."""

    def __init__(self):
        super().__init__()
        self.num_features = self.embed_dim = DIM
        self.patch_embed = _PatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, DIM))
        self.pos_embed = nn.Parameter(torch.randn(1, TOKENS, DIM) * 0.02)
        self.blocks = nn.Sequential(*[_Block() for _ in range(DEPTH)])
        self.norm = nn.LayerNorm(DIM, eps=1e-6)

    def forward(self, x):                       # eager (training / autograd) path
        b = x.shape[0]
        t = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
        t = torch.cat((self.cls_token.expand(b, -1, -1), t), dim=1) + self.pos_embed
        for blk in self.blocks:
            h = blk.norm1(t)
            qkv = blk.attn.qkv(h).reshape(b, TOKENS, 3, HEADS, DIM // HEADS).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            t = t + blk.attn.proj(a.transpose(1, 2).reshape(b, TOKENS, DIM))
            t = t + blk.mlp.fc2(F.gelu(blk.mlp.fc1(blk.norm2(t))))
        return self.norm(t[:, 0])


class ViTFeatureExtractor(nn.Module):            # models.py:88-107
    def __init__(self, model_name="vit_base_patch16_224", pretrained=False, out_dim=768, precision: str = DEFAULT_PRECISION):
        super().__init__()
        if model_name != "vit_base_patch16_224":
            raise ValueError(f"ViTFeatureExtractor (B200 build): only vit_base_patch16_224 is built, got {model_name!r}")
        if pretrained:
            raise RuntimeError("ViTFeatureExtractor (B200 build): no pretrained download offline; load a state_dict instead")
        self.vit = VisionTransformerB16()
        self.out_dim = self.vit.num_features
        self.precision = precision
        self._handle, self._key = None, None

    def _pack(self, device):
        tensors = list(self.state_dict(keep_vars=True).items())
        key = (str(device), self.precision, tuple((t.data_ptr(), t._version) for _, t in tensors))
        if self._handle is not None and key == self._key:
            return self._handle
        lib = _lib.load()
        if self._handle is not None:
            lib.dfd_vit_free_weights(self._handle)
        keep = [(k.encode(), v.detach().to("cpu", torch.float32).contiguous()) for k, v in tensors]
        n = len(keep)
        names = (C.c_char_p * n)(*[k for k, _ in keep])
        data = (C.c_void_p * n)(*[v.data_ptr() for _, v in keep])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in keep])
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.dfd_vit_pack_weights(n, names, data, numel, PRECISIONS[self.precision], C.byref(h))
        if rc:
            raise RuntimeError(f"dfd_vit_pack_weights failed ({rc}): {lib.dfd_vit_last_error().decode()}")
        self._handle, self._key = h, key
        return h

    def __del__(self):                    # release the packed device weights with the module
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().dfd_vit_free_weights(self._handle)
                self._handle = None
        except Exception:
            pass

    def forward(self, x):
        if self.training:
            return self.vit(x)
        if x.device.type != "cuda":
            raise RuntimeError("ViTFeatureExtractor (B200 build): inference needs a CUDA tensor; there is no CPU fallback")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"ViTFeatureExtractor: expected (B,3,224,224), got {tuple(x.shape)}")
        lib = _lib.load()
        h = self._pack(x.device)
        x = x.contiguous().float()
        out = torch.empty((x.shape[0], DIM), dtype=torch.float32, device=x.device)
        ws, ws_bytes = None, 0
        for i0 in range(0, x.shape[0], MAX_IMAGES_PER_CALL):
            n = min(MAX_IMAGES_PER_CALL, x.shape[0] - i0)
            nbytes = C.c_size_t()
            if lib.dfd_vit_workspace_bytes(n, C.byref(nbytes)):
                raise RuntimeError(lib.dfd_vit_last_error().decode())
            if ws is None or nbytes.value > ws_bytes:
                ws, ws_bytes = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device), nbytes.value
            with torch.cuda.device(x.device):
                rc = lib.dfd_vit_features(h, x[i0:i0 + n].data_ptr(), n, out[i0:i0 + n].data_ptr(), ws.data_ptr(), ws_bytes,
                                          _stream_ptr(x.device))
            if rc:
                raise RuntimeError(f"dfd_vit_features failed ({rc}): {lib.dfd_vit_last_error().decode()}")
        return out


class SimpleGCN(nn.Module):                      # models.py:177-197 (parameter container + eager forward)
    def __init__(self, in_dim, hid_dim=256, out_dim=128, dropout=0.3):
        super().__init__()
        self.fc1, self.fc2, self.dropout = nn.Linear(in_dim, hid_dim), nn.Linear(hid_dim, out_dim), nn.Dropout(dropout)

    def forward(self, H, A_norm):
        H = self.dropout(F.relu(self.fc1(torch.bmm(A_norm, H))))
        return F.relu(self.fc2(H))


class DeepfakeModel(nn.Module):
    """Drop-in for `src/models.DeepfakeModel` (models.py:199-291) with the `timm_vit` backbone: ViT-B/16 frame features ->
    SimpleGCN over the frame graph -> mean pool -> classifier.  `forward(images[B,N,3,224,224], A_norm[B,N,N]) -> [B,C]`.
    Same children / state_dict keys (`vit.vit.*`, `gcn.*`, `classifier.*`).  Eval mode runs in libdfd_b200.so (ViT encoder
    + one fused head kernel per video); training mode keeps the eager graph.  CLIP / DINOv2 backbones are not built."""

    def __init__(self, vit_out=768, gcn_hid=256, gcn_out=128, num_classes=2, pretrained_vit=False,
                 vit_model_name="vit_base_patch16_224", vit_pretrained_path=None, backbone: str = "timm_vit",
                 precision: str = DEFAULT_PRECISION, **_unused):
        super().__init__()
        if (backbone or "timm_vit").lower() not in {"timm_vit", "vit", "timm"}:
            raise ValueError(f"DeepfakeModel (B200 build): only the timm_vit backbone is built, got {backbone!r}")
        if (vit_out, gcn_hid, gcn_out) != (768, 256, 128):
            raise ValueError("DeepfakeModel (B200 build): built for the reference's default sizes (768, 256, 128)")
        self.vit = ViTFeatureExtractor(model_name=vit_model_name, pretrained=pretrained_vit, out_dim=vit_out, precision=precision)
        self.vit_proj = nn.Identity()
        if vit_pretrained_path is not None:
            self.vit.load_state_dict(torch.load(vit_pretrained_path, map_location="cpu"))
        self.gcn = SimpleGCN(in_dim=vit_out, hid_dim=gcn_hid, out_dim=gcn_out)
        self.classifier = nn.Sequential(nn.Linear(gcn_out, 64), nn.ReLU(), nn.Dropout(0.3), nn.Linear(64, num_classes))
        self.num_classes = num_classes
        self._head, self._head_key = None, None

    def _pack_head(self, device):
        tensors = [(k, v) for k, v in self.state_dict(keep_vars=True).items() if not k.startswith("vit.")]
        key = (str(device), tuple((t.data_ptr(), t._version) for _, t in tensors))
        if self._head is not None and key == self._head_key:
            return self._head
        lib = _lib.load()
        if self._head is not None:
            lib.dfd_gcn_free_weights(self._head)
        keep = [(k.encode(), v.detach().to("cpu", torch.float32).contiguous()) for k, v in tensors]
        n = len(keep)
        names = (C.c_char_p * n)(*[k for k, _ in keep])
        data = (C.c_void_p * n)(*[v.data_ptr() for _, v in keep])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in keep])
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.dfd_gcn_pack_weights(n, names, data, numel, self.num_classes, C.byref(h))
        if rc:
            raise RuntimeError(f"dfd_gcn_pack_weights failed ({rc}): {lib.dfd_vit_last_error().decode()}")
        self._head, self._head_key = h, key
        return h

    def __del__(self):
        try:
            if getattr(self, "_head", None) is not None:
                _lib.load().dfd_gcn_free_weights(self._head)
                self._head = None
        except Exception:
            pass

    def forward(self, images, A_norm):
        B, N, Cc, H, W = images.shape
        x = images.reshape(B * N, Cc, H, W)
        if self.training:
            feats = self.vit_proj(self.vit(x)).view(B, N, -1)
            return self.classifier(self.gcn(feats, A_norm).mean(dim=1))
        feats = self.vit(x)                                        # (B*N, 768) fp32, raises on CPU tensors
        lib = _lib.load()
        h = self._pack_head(images.device)
        adj = A_norm.to(device=images.device, dtype=torch.float32).contiguous()
        if tuple(adj.shape) != (B, N, N):
            raise ValueError(f"DeepfakeModel: A_norm must be ({B},{N},{N}), got {tuple(adj.shape)}")
        out = torch.empty((B, self.num_classes), dtype=torch.float32, device=images.device)
        with torch.cuda.device(images.device):
            rc = lib.dfd_gcn_head(h, feats.data_ptr(), adj.data_ptr(), B, N, out.data_ptr(), _stream_ptr(images.device))
        if rc:
            raise RuntimeError(f"dfd_gcn_head failed ({rc}): {lib.dfd_vit_last_error().decode()}")
        return out
