"""Synthetic weights and face crops for benchmarks, smoke runs and tests (no arithmetic of the scoring path).

The checkpoint named by the task (`checkpoints/pretrained_dfdc200_20260125/checkpoint_best_efficientnet_b0.pt`) is an
absent Git-LFS blob and PyTorch default init collapses the trunk output to ~1e-12 (F4), so
benchmarks and parity checks use a CALIBRATED synthetic checkpoint with the reference's exact state_dict
schema (366 tensors):

  * conv / linear weights: seeded `torch.Generator` draws (bit-reproducible on any CPU);
  * BatchNorm running statistics and the rescaled head: computed once by `oracle/synth_checkpoint.py`
    (layer-by-layer calibration on synthetic crops, i.e. what a trained network's running stats look like)
    and frozen in `tests/golden/synth_calib_seed{seed}.npz`; this module only reads that file.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (repeats, kernel, stride, expand, out_channels): timm efficientnet_b0 stage spec
B0_STAGES = ((1, 3, 1, 1, 16), (2, 3, 2, 6, 24), (2, 5, 2, 6, 40), (3, 3, 2, 6, 80),
             (3, 5, 1, 6, 112), (4, 5, 2, 6, 192), (1, 3, 1, 6, 320))


def block_specs():
    """[(key_prefix, cin, mid, cout, k, stride, rd, has_expand, has_skip)] for the 16 MBConv blocks."""
    out, cin = [], 32
    for s, (r, k, st, e, cout) in enumerate(B0_STAGES):
        for b in range(r):
            stride = st if b == 0 else 1
            out.append((f"backbone.2.{s}.{b}", cin, cin * e, cout, k, stride,
                        max(1, round(cin * 0.25)), e != 1, stride == 1 and cin == cout))
            cin = cout
    return out

# --------------------------------------------------------------------------- synthetic crops
def synth_crops(seed: int, n_videos: int, frames_per_video, size: int = 224) -> tuple[np.ndarray, np.ndarray]:
    """Smooth, face-crop-like uint8 RGB frames.  Returns crops (F,size,size,3) uint8 and offsets (V+1,) int32.

    numpy PCG64 draws + integer-exact upsampling, so the bytes are identical on every machine.
    `frames_per_video` is an int or a per-video sequence (ragged videos)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if np.isscalar(frames_per_video):
        frames_per_video = [int(frames_per_video)] * n_videos
    offsets = np.zeros(n_videos + 1, np.int32)
    offsets[1:] = np.cumsum(frames_per_video)
    out = np.empty((int(offsets[-1]), size, size, 3), np.uint8)
    assert size % 28 == 0

    def up(a, f):      # nearest-neighbour block upsample (integer exact)
        return np.repeat(np.repeat(a, f, axis=0), f, axis=1)

    for v in range(n_videos):
        base_lo = rng.integers(40, 216, (7, 7, 3)).astype(np.int32)          # video-level low-frequency field
        base_mid = rng.integers(-40, 41, (28, 28, 3)).astype(np.int32)
        for t in range(frames_per_video[v]):
            lo = base_lo + rng.integers(-12, 13, (7, 7, 3))
            mid = base_mid + rng.integers(-10, 11, (28, 28, 3))
            img = up(lo, size // 7) + up(mid, size // 28) + rng.integers(-16, 17, (size, size, 3))
            # cheap integer box blur (3 taps each way) so edges are not block-aligned
            img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0)) // 3
            img = (img + np.roll(img, 1, 1) + np.roll(img, -1, 1)) // 3
            out[offsets[v] + t] = np.clip(img, 0, 255).astype(np.uint8)
    return out, offsets


# --------------------------------------------------------------------------- weights
def _schema():
    """[(key, shape)] of the reference state_dict, in the reference's order."""
    keys = [("backbone.0.weight", (32, 3, 3, 3))]

    def bn(p, c):
        return [(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())]

    keys += bn("backbone.1", 32)
    for (p, cin, mid, cout, k, stride, rd, has_expand, has_skip) in block_specs():
        if has_expand:
            keys += [(p + ".conv_pw.weight", (mid, cin, 1, 1))] + bn(p + ".bn1", mid)
            keys += [(p + ".conv_dw.weight", (mid, 1, k, k))] + bn(p + ".bn2", mid)
        else:
            keys += [(p + ".conv_dw.weight", (mid, 1, k, k))] + bn(p + ".bn1", mid)
        keys += [(p + ".se.conv_reduce.weight", (rd, mid, 1, 1)), (p + ".se.conv_reduce.bias", (rd,)),
                 (p + ".se.conv_expand.weight", (mid, rd, 1, 1)), (p + ".se.conv_expand.bias", (mid,))]
        if has_expand:
            keys += [(p + ".conv_pwl.weight", (cout, mid, 1, 1))] + bn(p + ".bn3", cout)
        else:
            keys += [(p + ".conv_pw.weight", (cout, mid, 1, 1))] + bn(p + ".bn2", cout)
    keys += [("backbone.3.weight", (1280, 320, 1, 1))] + bn("backbone.4", 1280)
    keys += [("temporal_attention.0.weight", (64, 1280)), ("temporal_attention.0.bias", (64,)),
             ("temporal_attention.2.weight", (1, 64)), ("temporal_attention.2.bias", (1,)),
             ("fc1.weight", (256, 1280)), ("fc1.bias", (256,)), ("fc2.weight", (2, 256)), ("fc2.bias", (2,))]
    return keys


def schema():
    return _schema()


def seeded_weights(seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in _schema():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(1000, dtype=torch.int64)
        elif key.endswith("running_mean"):
            sd[key] = torch.zeros(shape)
        elif key.endswith("running_var"):
            sd[key] = torch.ones(shape)
        elif ".bn" in key or key.startswith("backbone.1.") or key.startswith("backbone.4."):
            is_linear_bn = key.rsplit(".", 2)[-2] in ("bn3",) or (".0.0.bn2" in key)   # project BNs (no act)
            if key.endswith(".weight"):
                lo, hi = (0.4, 0.9) if is_linear_bn else (0.6, 1.4)
                sd[key] = lo + (hi - lo) * torch.rand(shape, generator=g)
            else:
                sd[key] = 0.25 * torch.randn(shape, generator=g)
        elif "se.conv_reduce.weight" in key:
            sd[key] = torch.randn(shape, generator=g) * (2.0 / shape[1]) ** 0.5
        elif "se.conv_reduce.bias" in key:
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif "se.conv_expand.weight" in key:
            sd[key] = torch.randn(shape, generator=g) * (2.0 / shape[1]) ** 0.5
        elif "se.conv_expand.bias" in key:
            sd[key] = 0.3 * torch.randn(shape, generator=g)
        elif key.endswith(".bias"):                       # linear biases
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        else:                                             # conv / linear weights: fan-in scaled
            fan_in = int(np.prod(shape[1:]))
            sd[key] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def frozen_path(seed: int) -> str:
    return os.path.join(GOLDEN_DIR, f"synth_calib_seed{seed}.npz")


def apply_frozen(sd: dict, seed: int) -> dict:
    """Overlay the frozen data-dependent tensors (BN running stats, calibrated head) onto seeded weights."""
    path = frozen_path(seed)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python -m oracle.make_golden` in the build container")
    z = np.load(path)
    for k in z.files:
        if k == "__fc1_scale__":
            sd["fc1.weight"] = (sd["fc1.weight"].double() * float(z[k])).float()
        else:
            sd[k] = torch.from_numpy(z[k].copy())
    return {k: sd[k].contiguous() for k, _ in _schema()}


def load_checkpoint(seed: int = 0) -> dict:
    """The calibrated synthetic state_dict (fp32, reference schema and key order)."""
    return apply_frozen(seeded_weights(seed), seed)
