"""Engine of the `resnet50` ensemble member (reference src/pretrained_detector.py:38-41 + :103-143): packs a member's
state_dict (torchvision trunk keys under `backbone.*` + the pool/head keys) and scores videos through csrc/resnet.cu.
No CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import PRECISIONS, _stream_ptr, normalize_key

FEATURE_DIM = 2048


class ResNet50Scorer:
    def __init__(self, state_dict, precision: str, device, use_temporal_attention: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ResNet50Scorer: a CUDA device is required; there is no CPU fallback")
        self.use_temporal_attention = use_temporal_attention
        lib = _lib.load()
        keep = [(normalize_key(k).encode(), v.detach().to("cpu", torch.float32).contiguous()) for k, v in state_dict.items()
                if torch.is_tensor(v) and v.is_floating_point()]
        n = len(keep)
        names = (C.c_char_p * n)(*[k for k, _ in keep])
        data = (C.c_void_p * n)(*[v.data_ptr() for _, v in keep])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in keep])
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.dfd_resnet50_pack_weights(n, names, data, numel, PRECISIONS[precision], C.byref(self.handle))
        if rc:
            raise RuntimeError(f"dfd_resnet50_pack_weights failed ({rc}): {lib.dfd_resnet_last_error().decode()}")

    def free(self):
        if self.handle:
            _lib.load().dfd_resnet50_free_weights(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def score(self, x_flat: torch.Tensor, offsets: torch.Tensor, max_frames: int, use_temporal_attention=None):
        """x_flat fp32 (F,3,224,224) on the device, offsets int32 (V+1) -> logits (V,2), frame_scores (F)."""
        if x_flat.device.type != "cuda" or tuple(x_flat.shape[1:]) != (3, 224, 224):
            raise ValueError("ResNet50Scorer.score: expected a CUDA tensor (F,3,224,224); there is no CPU fallback")
        lib = _lib.load()
        x_flat = x_flat.contiguous().float()
        frames, videos = x_flat.shape[0], offsets.numel() - 1
        att = self.use_temporal_attention if use_temporal_attention is None else use_temporal_attention
        logits = torch.empty((videos, 2), dtype=torch.float32, device=x_flat.device)
        scores = torch.empty((frames,), dtype=torch.float32, device=x_flat.device)
        nbytes = C.c_size_t()
        if lib.dfd_resnet50_workspace_bytes(frames, C.byref(nbytes)):
            raise RuntimeError(lib.dfd_resnet_last_error().decode())
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x_flat.device)
        with torch.cuda.device(x_flat.device):
            rc = lib.dfd_resnet50_score_videos(self.handle, x_flat.data_ptr(), offsets.contiguous().data_ptr(), videos, frames, int(max_frames),
                                               1 if att else 0, logits.data_ptr(), scores.data_ptr(), None, ws.data_ptr(), nbytes.value,
                                               _stream_ptr(x_flat.device))
        if rc:
            raise RuntimeError(f"dfd_resnet50_score_videos failed ({rc}): {lib.dfd_resnet_last_error().decode()}")
        return logits, scores
