"""Calibrated synthetic checkpoint + synthetic face-crop generator.  TEST INFRASTRUCTURE, NOT PRODUCT.

Why: the checkpoint named by the task (`checkpoints/pretrained_dfdc200_20260125/checkpoint_best_efficientnet_b0.pt`)
is an absent Git-LFS blob, and PyTorch default init collapses the trunk output to ~1e-12
(SURVEY.md F4), which would make any parity check vacuous.  This module builds a 366-key state_dict with
the reference's exact schema whose activations stay O(1) through all 81 conv layers:

  * conv / linear weights: seeded `torch.Generator` draws (bit-reproducible on any CPU);
  * BatchNorm running statistics: measured layer by layer on a calibration batch of synthetic crops and
    then perturbed, i.e. what a trained network's running stats look like.  They depend on conv outputs,
    which are not bit-identical across CPUs, so they are frozen once in
    `tests/golden/synth_bn_stats_seed{seed}.npz` (written by `oracle/make_golden.py`) and re-read from there;
  * head: `fc2` rescaled so the logit margin over the calibration videos has std ≈ 2 (verdicts are O(1) decisions).
"""
from __future__ import annotations

import os
import numpy as np
import torch

from . import effnet_b0_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


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
    for (p, cin, mid, cout, k, stride, rd, has_expand, has_skip) in O.block_specs():
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


def _seeded_weights(seed: int) -> dict:
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


def _calibrate_bn(sd: dict, seed: int) -> None:
    """Set every BN's running stats to (perturbed) batch statistics of a calibration batch."""
    crops, _ = synth_crops(seed + 1000, 6, 4)
    x = O.prep_u8_hwc(crops)
    g = torch.Generator().manual_seed(seed + 1)

    def hook(p, pre):
        mean = pre.mean((0, 2, 3))
        var = pre.var((0, 2, 3), unbiased=False)
        sd[p + ".running_mean"] = (mean + 0.1 * var.sqrt() * torch.randn(mean.shape, generator=g)).contiguous()
        sd[p + ".running_var"] = (var * (0.8 + 0.45 * torch.rand(var.shape, generator=g)) + 1e-4).contiguous()

    with torch.no_grad():
        O.trunk_features(sd, x, bn_hook=hook)


def _calibrate_head(sd: dict, seed: int) -> None:
    crops, offs = synth_crops(seed + 2000, 24, 4)
    with torch.no_grad():
        feats = O.trunk_features(sd, O.prep_u8_hwc(crops)).view(24, 4, -1)
        # centre fc1 pre-activations so roughly half the units are live
        pooled = feats.mean(1)
        pre = torch.nn.functional.linear(pooled, sd["fc1.weight"])
        sd["fc1.weight"] = sd["fc1.weight"] / (pre.std() + 1e-6)
        sd["fc1.bias"] = -(pre / (pre.std() + 1e-6)).mean(0)
        # attention MLP: make the pre-sigmoid score spread O(1) across frames
        a = torch.relu(torch.nn.functional.linear(feats, sd["temporal_attention.0.weight"], sd["temporal_attention.0.bias"]))
        s = torch.nn.functional.linear(a, sd["temporal_attention.2.weight"])
        sd["temporal_attention.2.weight"] = sd["temporal_attention.2.weight"] * (1.5 / (s.std() + 1e-6))
        sd["temporal_attention.2.bias"] = -(s * (1.5 / (s.std() + 1e-6))).mean().reshape(1)
        logits, _ = O.attention_pool_head(sd, feats)
        margin = logits[:, 1] - logits[:, 0]
        scale = 2.0 / (margin.std() + 1e-6)
        sd["fc2.weight"] = sd["fc2.weight"] * scale
        logits, _ = O.attention_pool_head(sd, feats)
        margin = logits[:, 1] - logits[:, 0]
        sd["fc2.bias"] = sd["fc2.bias"] + torch.tensor([0.5, -0.5]) * margin.median()


_HEAD_KEYS = ("temporal_attention.0.weight", "temporal_attention.0.bias", "temporal_attention.2.weight",
              "temporal_attention.2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


def frozen_path(seed: int) -> str:
    return os.path.join(GOLDEN_DIR, f"synth_calib_seed{seed}.npz")


def make_checkpoint(seed: int = 0, use_frozen: bool = True, freeze: bool = False) -> dict:
    """The calibrated synthetic state_dict (fp32, reference schema, reference key order).

    With `use_frozen` the data-dependent pieces (BN running stats, rescaled head) come from the committed
    fixture so every machine gets the same bytes; `freeze=True` (make_golden.py) recomputes and writes it."""
    sd = _seeded_weights(seed)
    path = frozen_path(seed)
    if use_frozen and not freeze and os.path.exists(path):
        z = np.load(path)
        for k in z.files:
            sd[k] = torch.from_numpy(z[k].copy())
    else:
        _calibrate_bn(sd, seed)
        _calibrate_head(sd, seed)
        if freeze:
            blob = {k: v.numpy() for k, v in sd.items()
                    if k.endswith("running_mean") or k.endswith("running_var")}
            blob.update({k: sd[k].numpy() for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
                                                     "temporal_attention.2.weight", "temporal_attention.2.bias")})
            # fc1.weight is a pure rescale of seeded draws; store only its scalar to keep the fixture small
            ratio = float((sd["fc1.weight"].flatten()[0] / _seeded_weights(seed)["fc1.weight"].flatten()[0]))
            del blob["fc1.weight"]
            blob["__fc1_scale__"] = np.float64(ratio)
            np.savez_compressed(path, **blob)
            return make_checkpoint(seed, use_frozen=True, freeze=False)
    if "__fc1_scale__" in sd:
        ratio = float(sd.pop("__fc1_scale__"))
        sd["fc1.weight"] = (_seeded_weights(seed)["fc1.weight"].double() * ratio).float()
    return {k: sd[k].contiguous() for k, _ in _schema()}
