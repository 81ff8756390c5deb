"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the crop + resize step in front of the scoring path.

Reference call sites: `app.py:1964-1978` and `src/data_prepare.py:54-56` — `pil.crop((x1, y1, x2, y2)).resize((S, S))`.
The arithmetic lives in Pillow (third-party, `requirements.txt:17`: `pillow>=11.0.0`, no lock file; this image has 12.2.0), whose
`Image.resize` default for RGB images is BICUBIC.  Restated here from Pillow's published algorithm
(src/libImaging/Resample.c: `precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`,
`ImagingResampleVertical_8bpc`):
  * separable: horizontal pass first, 8-bit (rounded, clipped) intermediate image, then the vertical pass;
  * antialiased: support = 2.0 * max(scale, 1), scale = in / out; window of output pixel xx is
    [int(center - support + 0.5), int(center + support + 0.5)) clipped to the crop, center = (xx + 0.5) * scale;
  * bicubic kernel with a = -0.5, weights normalised by their sum in double precision, then converted to 22-bit
    fixed point with round-half-away-from-zero; accumulation starts from 1 << 21 and ends with >> 22 and a clip to [0, 255].
Pinned bit-exactly against Pillow itself (tests/test_resize.py), which is installed here and on the GPU box.
Only tests/ may import this module."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int):
    """-> bounds int32 (out,2) = (xmin, count), kk int32 (out, ksize): Pillow's precompute_coeffs + normalize_coeffs_8bpc."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample axis 0 of img (n, ..., 3) uint8 to out_size."""
    bounds, kk = coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[x0 + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def crop_resize(frame: np.ndarray, box, out_size: int = 224) -> np.ndarray:
    """frame (H,W,3) uint8, box (x1,y1,x2,y2) already clamped as at app.py:1968-1973 -> (S,S,3) uint8."""
    x1, y1, x2, y2 = box
    crop = frame[y1:y2, x1:x2]
    tmp = _pass(np.ascontiguousarray(crop.transpose(1, 0, 2)), out_size).transpose(1, 0, 2)     # horizontal pass first
    return _pass(np.ascontiguousarray(tmp), out_size)                                             # then vertical
