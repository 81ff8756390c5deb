"""Crop + resize of face boxes on the GPU, bit-exact with the reference's `pil.crop(box).resize((S, S))`
(app.py:1964-1978, src/data_prepare.py:54-56; Pillow BICUBIC).  Output feeds `FrameScorer.score` directly."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib
from .engine import _stream_ptr


def clamp_boxes(boxes: Sequence[Sequence[float]], width: int, height: int):
    """The reference's box handling (app.py:1964-1975): int() truncation, clamp to the frame, drop empty boxes.
    boxes rows = (frame_index, x1, y1, x2, y2) -> (kept rows as ints, indices of the kept rows)."""
    kept, idx = [], []
    for i, (f, x1, y1, x2, y2) in enumerate(boxes):
        x1, y1, x2, y2 = int(x1), int(y1), int(x2), int(y2)
        x1 = max(0, min(width, x1)); x2 = max(0, min(width, x2))
        y1 = max(0, min(height, y1)); y2 = max(0, min(height, y2))
        if x2 <= x1 or y2 <= y1:
            continue
        kept.append((int(f), x1, y1, x2, y2)); idx.append(i)
    return kept, idx


def crop_resize(frames: torch.Tensor, boxes: Sequence[Sequence[float]], size: int = 224) -> Tuple[torch.Tensor, list]:
    """frames uint8 (N,H,W,3) on a CUDA device, boxes rows (frame_index, x1, y1, x2, y2) -> (faces uint8 (M,size,size,3),
    indices of the boxes that survived the reference's clamping)."""
    if frames.device.type != "cuda" or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("crop_resize: frames must be a uint8 (N,H,W,3) CUDA tensor; there is no CPU fallback")
    frames = frames.contiguous()
    n_frames, h, w, _ = frames.shape
    kept, idx = clamp_boxes(boxes, w, h)
    out = torch.empty((len(kept), size, size, 3), dtype=torch.uint8, device=frames.device)
    if not kept:
        return out, idx
    arr = (_lib.CropBox * len(kept))()
    for j, (f, x1, y1, x2, y2) in enumerate(kept):
        if not 0 <= f < n_frames:
            raise ValueError(f"crop_resize: frame index {f} out of range")
        arr[j] = _lib.CropBox(f * h * w * 3, w, h, x1, y1, x2, y2)
    lib = _lib.load()
    nbytes = C.c_size_t()
    if lib.dfd_crop_resize_workspace_bytes(arr, len(kept), size, C.byref(nbytes)):
        raise RuntimeError(lib.dfd_resize_last_error().decode())
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        rc = lib.dfd_crop_resize_u8(frames.data_ptr(), arr, len(kept), size, out.data_ptr(), ws.data_ptr(), nbytes.value,
                                    _stream_ptr(frames.device))
    if rc:
        raise RuntimeError(f"dfd_crop_resize_u8 failed ({rc}): {lib.dfd_resize_last_error().decode()}")
    return out, idx
