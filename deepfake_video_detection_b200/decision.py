"""Host-side pieces of the reference's inference glue that sit directly around the model call.

  imagenet_normalize   app.py:1772-1780 (same shape rules / error), kept for callers that already hold
                       float frames; the throughput path feeds uint8 crops and fuses this into the stem.
  decide               app.py:2090-2112, 2174, 2193: softmax -> prob_fake -> threshold -> verdict/abstain.
"""
from __future__ import annotations

from typing import List

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def imagenet_normalize(frames: torch.Tensor) -> torch.Tensor:
    mean = torch.tensor(IMAGENET_MEAN, device=frames.device, dtype=frames.dtype)
    std = torch.tensor(IMAGENET_STD, device=frames.device, dtype=frames.dtype)
    if frames.dim() == 4:
        return (frames - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    if frames.dim() == 5:
        return (frames - mean.view(1, 1, 3, 1, 1)) / std.view(1, 1, 3, 1, 1)
    raise ValueError(f"Unsupported frames shape for normalization: {tuple(frames.shape)}")


def decide(logits: torch.Tensor, threshold: float = 0.5, fake_idx: int = 1, abstain_conf: float = 0.60,
           abstain_margin: float = 0.0, allow_extreme_threshold: bool = False) -> List[dict]:
    """Per-video verdict dicts with the reference's field names (app.py:2212-2223)."""
    thr = float(threshold)
    if not allow_extreme_threshold and (thr < 0.05 or thr > 0.95):
        thr = 0.5
    probs = torch.softmax(logits.detach().float().cpu(), dim=1)
    out = []
    for v in range(probs.shape[0]):
        prob_fake, prob_real = float(probs[v, fake_idx]), float(probs[v, 1 - fake_idx])
        is_fake = prob_fake >= thr
        conf = prob_fake if is_fake else prob_real
        abstained = (abstain_margin > 0.0 and abs(prob_fake - thr) <= abstain_margin) or conf < abstain_conf
        out.append({"prediction": "Uncertain" if abstained else ("Deepfake" if is_fake else "Real"),
                    "verdict_yes_no": "Unsure" if abstained else ("Yes" if is_fake else "No"),
                    "pred_class": None if abstained else int(is_fake), "is_fake": bool(is_fake),
                    "confidence": conf, "prob_real": prob_real, "prob_fake": prob_fake, "threshold": thr,
                    "abstained": bool(abstained)})
    return out
