"""CPU oracle for the EfficientNet-B0 frame-scoring path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.
The shipped path (`deepfake_video_detection_b200/`) never does; it fails loudly when the CUDA
library is missing instead of falling back to anything here.

What is restated (reference = /root/reference, file:line):

  * tensor prep              app.py:2084-2086  (`from_numpy().permute(0,3,1,2).float()/255`)
  * imagenet_normalize       app.py:1772-1780
  * trunk                    src/pretrained_detector.py:42-49,116  — the arithmetic lives in the
                             un-vendored `timm` (requirements.txt:12, `timm>=0.9.0`, no lock file):
                             timm `efficientnet_b0` = conv_stem/bn1(SiLU)/7 stages/conv_head/bn2(SiLU)/avg-pool,
                             symmetric padding k//2, BN eps 1e-5, SE on the block-input/4 width.
  * temporal attention pool  src/pretrained_detector.py:65-71,123-131 (+ mean mode :132-135)
  * classification head      src/pretrained_detector.py:74-76,138-141   (dropout = identity in eval)
  * decision rule            app.py:2090-2094 (softmax, fake index 1), :2096-2112 (threshold 0.5 default,
                             extreme-threshold guard), :2174,:2193 (abstain margin / confidence)

Pinning (the reference ships no tests or golden vectors for this path — SURVEY.md §4, F8):
  1. `oracle/make_golden.py` imports the UNMODIFIED reference class in the build container (over
     `oracle/timm_standin`) and freezes its outputs into `tests/golden/`; `tests/test_oracle.py` checks
     this restatement against those vectors.
  2. The trunk is additionally checked against `torchvision.models.efficientnet_b0`, an independent
     implementation of the same published architecture (key map: SURVEY.md App. C).
Because the real checkpoint is an absent Git-LFS blob, parity is pinned on the
calibrated synthetic checkpoint of `oracle/synth_checkpoint.py`.

Everything is plain fp32 ATen on the CPU, driven by a reference-schema state_dict (366 keys).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # app.py:1774
IMAGENET_STD = (0.229, 0.224, 0.225)    # app.py:1775
BN_EPS = 1e-5

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


# --------------------------------------------------------------------------- prep (app.py:2084-2086)
def prep_u8_hwc(faces_u8) -> torch.Tensor:
    """uint8 (T,H,W,3) RGB → fp32 (T,3,H,W), `/255` then ImageNet mean/std (app.py:2084-2085)."""
    t = torch.as_tensor(np.asarray(faces_u8)) if not torch.is_tensor(faces_u8) else faces_u8
    x = t.permute(0, 3, 1, 2).float() / 255.0
    return imagenet_normalize(x)


def imagenet_normalize(frames: torch.Tensor) -> torch.Tensor:
    """app.py:1772-1780, same shape rules and error."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=frames.dtype)
    std = torch.tensor(IMAGENET_STD, dtype=frames.dtype)
    if frames.dim() == 4:
        return (frames - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    if frames.dim() == 5:
        return (frames - mean.view(1, 1, 3, 1, 1)) / std.view(1, 1, 3, 1, 1)
    raise ValueError(f"Unsupported frames shape for normalization: {tuple(frames.shape)}")


# --------------------------------------------------------------------------- trunk (timm efficientnet_b0)
def _bn(x, sd, p, act, hook=None):
    if hook is not None:          # used only by oracle/synth_checkpoint.py to calibrate BN statistics
        hook(p, x)
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                     sd[p + ".bias"], False, 0.0, BN_EPS)
    return F.silu(y) if act else y


def _se(x, sd, p):
    s = x.mean((2, 3), keepdim=True)
    s = F.silu(F.conv2d(s, sd[p + ".conv_reduce.weight"], sd[p + ".conv_reduce.bias"]))
    s = F.conv2d(s, sd[p + ".conv_expand.weight"], sd[p + ".conv_expand.bias"])
    return x * torch.sigmoid(s)


def trunk_features(sd, x: torch.Tensor, taps: dict | None = None, bn_hook=None) -> torch.Tensor:
    """x fp32 (F,3,H,W) normalised → pooled features (F,1280).  `taps` collects per-layer activations."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    y = F.conv2d(x, sd["backbone.0.weight"], None, 2, 1)
    y = tap("stem", _bn(y, sd, "backbone.1", True, bn_hook))
    for (p, cin, mid, cout, k, stride, rd, has_expand, has_skip) in block_specs():
        inp = y
        if has_expand:
            y = tap(p + ".expand", _bn(F.conv2d(y, sd[p + ".conv_pw.weight"]), sd, p + ".bn1", True, bn_hook))
            y = F.conv2d(y, sd[p + ".conv_dw.weight"], None, stride, k // 2, 1, mid)
            y = tap(p + ".dw", _bn(y, sd, p + ".bn2", True, bn_hook))
            y = _se(y, sd, p + ".se")
            y = _bn(F.conv2d(y, sd[p + ".conv_pwl.weight"]), sd, p + ".bn3", False, bn_hook)
        else:
            y = F.conv2d(y, sd[p + ".conv_dw.weight"], None, stride, k // 2, 1, mid)
            y = tap(p + ".dw", _bn(y, sd, p + ".bn1", True, bn_hook))
            y = _se(y, sd, p + ".se")
            y = _bn(F.conv2d(y, sd[p + ".conv_pw.weight"]), sd, p + ".bn2", False, bn_hook)
        if has_skip:
            y = y + inp
        tap(p + ".out", y)
    y = tap("head", _bn(F.conv2d(y, sd["backbone.3.weight"]), sd, "backbone.4", True, bn_hook))
    return y.mean((2, 3))


# --------------------------------------------------------------------------- pool + head (pretrained_detector.py:112-141)
def attention_pool_head(sd, features: torch.Tensor, use_temporal_attention: bool = True):
    """features (B,T,1280) → logits (B,2), frame_scores (B,T)."""
    B, T, _ = features.shape
    if use_temporal_attention:
        a = F.relu(F.linear(features, sd["temporal_attention.0.weight"], sd["temporal_attention.0.bias"]))
        a = torch.sigmoid(F.linear(a, sd["temporal_attention.2.weight"], sd["temporal_attention.2.bias"]))
        w = F.softmax(a.squeeze(-1), dim=1)
        pooled = (features * w.unsqueeze(-1)).sum(dim=1)
        frame_scores = w
    else:
        pooled = features.mean(dim=1)
        frame_scores = torch.ones(B, T) / T
    h = F.relu(F.linear(pooled, sd["fc1.weight"], sd["fc1.bias"]))
    logits = F.linear(h, sd["fc2.weight"], sd["fc2.bias"])
    return logits, frame_scores


def detector_forward(sd, x: torch.Tensor, use_temporal_attention: bool = True):
    """PretrainedBackboneDetector.forward (pretrained_detector.py:103-143): x (B,T,3,H,W) fp32."""
    B, T, c, h, w = x.shape
    with torch.no_grad():
        feats = trunk_features(sd, x.reshape(B * T, c, h, w)).view(B, T, -1)
        return attention_pool_head(sd, feats, use_temporal_attention)


def score_ragged(sd, crops_u8, offsets, use_temporal_attention: bool = True):
    """Per-video scoring of ragged videos = one reference B=1 call per video.

    crops_u8 (F,H,W,3) uint8; offsets (V+1,) frame offsets.  Returns logits (V,2), frame_scores (F,)."""
    logits, scores = [], []
    for v in range(len(offsets) - 1):
        a, b = int(offsets[v]), int(offsets[v + 1])
        x = prep_u8_hwc(crops_u8[a:b]).unsqueeze(0)
        lg, fs = detector_forward(sd, x, use_temporal_attention)
        logits.append(lg[0])
        scores.append(fs[0])
    return torch.stack(logits), torch.cat(scores)


# --------------------------------------------------------------------------- decision (app.py:2090-2112, 2174, 2193)
def decide(logits: torch.Tensor, threshold: float = 0.5, fake_idx: int = 1,
           abstain_conf: float = 0.60, abstain_margin: float = 0.0,
           allow_extreme_threshold: bool = False):
    """Per-video verdicts as app.py computes them.  Returns a list of dicts."""
    thr = float(threshold)
    if not allow_extreme_threshold and (thr < 0.05 or thr > 0.95):   # app.py:2106-2109
        thr = 0.5
    probs = torch.softmax(logits.float(), dim=1)
    out = []
    for v in range(logits.shape[0]):
        prob_fake = float(probs[v, fake_idx])
        prob_real = float(probs[v, 1 - fake_idx])
        is_fake = bool(prob_fake >= thr)
        conf = prob_fake if is_fake else prob_real
        abstained = (abs(prob_fake - thr) <= abstain_margin and abstain_margin > 0) or conf < abstain_conf
        out.append(dict(pred_class=1 if is_fake else 0, is_fake=is_fake, prob_fake=prob_fake,
                        prob_real=prob_real, confidence=conf, threshold=thr, abstained=bool(abstained)))
    return out


# --------------------------------------------------------------------------- BN folding used by tests
def fold_bn(w: torch.Tensor, sd, bn_prefix: str):
    """w' = w·γ/√(σ²+eps), b' = β − μ·γ/√(σ²+eps)  (fp32)."""
    scale = sd[bn_prefix + ".weight"] / torch.sqrt(sd[bn_prefix + ".running_var"] + BN_EPS)
    return w * scale.view(-1, *([1] * (w.dim() - 1))), sd[bn_prefix + ".bias"] - sd[bn_prefix + ".running_mean"] * scale


def score_ragged_batched(sd, crops_u8, offsets, use_temporal_attention: bool = True, chunk: int = 256):
    """`score_ragged` for large samples: the trunk (per-frame independent, pretrained_detector.py:116) runs over chunks of
    frames instead of one call per video; pool + head stay one B=1 call per video (:123-141).  Pinned against `score_ragged`
    in tests/test_oracle.py."""
    F_total = int(offsets[-1])
    feats = []
    with torch.no_grad():
        for a in range(0, F_total, chunk):
            feats.append(trunk_features(sd, prep_u8_hwc(crops_u8[a:min(F_total, a + chunk)])))
        feats = torch.cat(feats) if feats else torch.zeros(0, 1280)
        logits, scores = [], []
        for v in range(len(offsets) - 1):
            a, b = int(offsets[v]), int(offsets[v + 1])
            lg, fs = attention_pool_head(sd, feats[a:b].unsqueeze(0), use_temporal_attention)
            logits.append(lg[0])
            scores.append(fs[0])
    return torch.stack(logits), torch.cat(scores), feats
