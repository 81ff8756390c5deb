"""Drop-in for the reference's `src/pretrained_detector.py` (EfficientNet-B0 path + the `resnet50` ensemble member).

Same class names, constructor keywords, attributes, `forward(x) -> (logits, frame_scores)` contract and
state_dict schema (366 keys) as the reference (`src/pretrained_detector.py:15-143`,
`:146-218`), so `app.load_model` (`app.py:1691-1761`) and `InferenceAgent` (`agent_system.py:82-109`) work
unchanged.  In eval mode the arithmetic runs in libdfd_b200.so (hand-written sm_100a kernels); there is
no CPU or eager fallback for inference — a CPU tensor or a missing library raises.

The parameter containers below mirror timm>=0.9 `efficientnet_b0` child order and names (the reference
builds its trunk with `timm.create_model`, `:43`, which is not vendored); their eager `forward` is used
ONLY in training mode (`model.train()`: the reference's fine-tuning helpers need autograd), never for scoring.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import DEFAULT_PRECISION, FrameScorer

_STAGES = ((1, 3, 1, 1, 16), (2, 3, 2, 6, 24), (2, 5, 2, 6, 40), (3, 3, 2, 6, 80),
           (3, 5, 1, 6, 112), (4, 5, 2, 6, 192), (1, 3, 1, 6, 320))


class _BNAct(nn.BatchNorm2d):
    """BatchNorm2d (+SiLU) with nn.BatchNorm2d's parameter names (timm BatchNormAct2d)."""

    def __init__(self, c: int, act: bool):
        super().__init__(c, eps=1e-5, momentum=0.1)
        self.apply_act = act

    def forward(self, x):
        y = super().forward(x)
        return F.silu(y) if self.apply_act else y


class _SE(nn.Module):
    def __init__(self, c: int, rd: int):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, rd, 1)
        self.conv_expand = nn.Conv2d(rd, c, 1)

    def forward(self, x):
        s = self.conv_expand(F.silu(self.conv_reduce(x.mean((2, 3), keepdim=True))))
        return x * torch.sigmoid(s)


class _DSConv(nn.Module):          # timm DepthwiseSeparableConv (stage 0)
    def __init__(self, cin, cout, k, stride, rd):
        super().__init__()
        self.skip = stride == 1 and cin == cout
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = _BNAct(cin, True)
        self.se = _SE(cin, rd)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = _BNAct(cout, False)

    def forward(self, x):
        y = self.bn2(self.conv_pw(self.se(self.bn1(self.conv_dw(x)))))
        return x + y if self.skip else y


class _MBConv(nn.Module):          # timm InvertedResidual (stages 1-6)
    def __init__(self, cin, cout, k, stride, expand, rd):
        super().__init__()
        mid = cin * expand
        self.skip = stride == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = _BNAct(mid, True)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = _BNAct(mid, True)
        self.se = _SE(mid, rd)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = _BNAct(cout, False)

    def forward(self, x):
        y = self.bn2(self.conv_dw(self.bn1(self.conv_pw(x))))
        y = self.bn3(self.conv_pwl(self.se(y)))
        return x + y if self.skip else y


class _AvgPoolFlatten(nn.Module):
    def forward(self, x):
        return x.mean((2, 3))


def _efficientnet_b0_trunk() -> nn.Sequential:
    """children()[:-1] of timm efficientnet_b0: conv_stem, bn1, blocks, conv_head, bn2, global_pool."""
    stages, cin = [], 32
    for (r, k, s, e, cout) in _STAGES:
        blocks = []
        for b in range(r):
            rd = max(1, round(cin * 0.25))
            st = s if b == 0 else 1
            blocks.append(_DSConv(cin, cout, k, st, rd) if e == 1 else _MBConv(cin, cout, k, st, e, rd))
            cin = cout
        stages.append(nn.Sequential(*blocks))
    return nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), _BNAct(32, True), nn.Sequential(*stages),
                         nn.Conv2d(320, 1280, 1, bias=False), _BNAct(1280, True), _AvgPoolFlatten())


class PretrainedBackboneDetector(nn.Module):
    """Reference: src/pretrained_detector.py:15-143.  Extra keyword: `precision` ("fp16" | "bf16")."""

    def __init__(self, backbone_name: str = "efficientnet_b0", pretrained: bool = True, num_classes: int = 2,
                 dropout_rate: float = 0.5, freeze_backbone: bool = False, use_temporal_attention: bool = True,
                 precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.backbone_name = backbone_name
        self.num_classes = num_classes
        self.use_temporal_attention = use_temporal_attention
        self.precision = precision
        if backbone_name not in ("efficientnet_b0", "resnet50"):
            # resnet18/34, vit* and the other efficientnets of the reference (:38-56) are outside this path
            raise ValueError(f"Unsupported backbone: {backbone_name} (the B200 path implements efficientnet_b0 and resnet50)")
        if num_classes != 2:
            raise ValueError("the B200 head kernel implements num_classes=2 (the reference's only use)")
        # `pretrained=True` means "download ImageNet weights" in the reference; there is no network here and
        # every reference caller loads a checkpoint afterwards, so the flag only selects the init below.
        if backbone_name == "resnet50":                                       # :38-41 (the ensemble's default second member)
            import torchvision                                                # the reference builds this trunk from torchvision itself
            self.backbone = nn.Sequential(*list(torchvision.models.resnet50(weights=None).children())[:-1])
            self.feature_dim = 2048
        else:
            self.backbone = _efficientnet_b0_trunk()
            self.feature_dim = 1280                                           # :49
        if freeze_backbone:                                                   # :58-61
            for p in self.backbone.parameters():
                p.requires_grad = False
        if use_temporal_attention:                                            # :64-71
            self.temporal_attention = nn.Sequential(nn.Linear(self.feature_dim, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        self.dropout = nn.Dropout(dropout_rate)                               # :74-76
        self.fc1 = nn.Linear(self.feature_dim, 256)
        self.fc2 = nn.Linear(256, num_classes)
        self._init_head_weights()
        self._scorer: Optional[FrameScorer] = None
        self._scorer_key = None
        self._generation = 0             # bumped by _apply() (.to / .half / .cuda replace tensor storage)
        self._flat = None                # cached [tensor, ...] of the state_dict, rebuilt when the generation changes

    def _init_head_weights(self):                                             # :80-85
        nn.init.kaiming_normal_(self.fc1.weight, mode="fan_out", nonlinearity="relu")
        nn.init.constant_(self.fc1.bias, 0)
        nn.init.normal_(self.fc2.weight, 0, 0.01)
        nn.init.constant_(self.fc2.bias, 0)

    def unfreeze_backbone(self, num_blocks: int = 2):                         # :87-101 (stage granularity)
        stages = list(self.backbone.children()) if self.backbone_name == "resnet50" else list(self.backbone[2])
        for stage in stages[-num_blocks:]:
            for p in stage.parameters():
                p.requires_grad = True

    # ---- CUDA engine plumbing ----------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        self._generation += 1
        self._flat = None
        return super()._apply(fn, *args, **kwargs)

    def _engine(self, device: torch.device) -> FrameScorer:
        # Cheap change detection (the reference calls forward once per request, app.py:2089): in-place updates (load_state_dict,
        # optimiser steps, `.add_`) bump the tensors' version counters, storage moves go through _apply().  Walking the
        # module tree (state_dict) only happens when the key changes; replacing a submodule by assignment needs `refresh()`.
        if self._flat is None or self._flat[0] != self._generation:
            self._flat = (self._generation, [t for t in self.state_dict(keep_vars=True).values()])
        key = (str(device), self.precision, self._generation, sum(t._version for t in self._flat[1]))
        if self._scorer is None or key != self._scorer_key:
            if self._scorer is not None:
                (self._scorer if self.backbone_name == "resnet50" else self._scorer.weights).free()
            tensors = list(self.state_dict(keep_vars=True).items())
            self._flat = (self._generation, [t for _, t in tensors])
            key = (str(device), self.precision, self._generation, sum(t._version for t in self._flat[1]))
            sd = {k: v for k, v in tensors}
            if not self.use_temporal_attention:     # mean mode has no attention MLP; the packer wants the keys
                z = torch.zeros
                sd.update({"temporal_attention.0.weight": z(64, self.feature_dim), "temporal_attention.0.bias": z(64),
                           "temporal_attention.2.weight": z(1, 64), "temporal_attention.2.bias": z(1)})
            if self.backbone_name == "resnet50":
                from .resnet_model import ResNet50Scorer
                self._scorer = ResNet50Scorer(sd, self.precision, device, self.use_temporal_attention)
            else:
                self._scorer = FrameScorer(sd, self.precision, device, self.use_temporal_attention)
            self._scorer_key = key
        return self._scorer

    def refresh(self) -> None:
        """Force a repack of the weights on the next forward (after replacing a submodule or a Parameter object)."""
        self._generation += 1
        self._flat = None

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """x (B,T,C,H,W) float -> logits (B,num_classes), frame_scores (B,T)   (:103-143)."""
        batch_size, num_frames, c, h, w = x.shape
        if self.training:
            return self._forward_eager(x)
        if num_frames < 1 or num_frames > 1024:
            raise ValueError(f"forward: between 1 and 1024 frames per video, got {num_frames} (the reference serves at most 64, app.py:2053)")
        if x.device.type != "cuda":
            raise RuntimeError("PretrainedBackboneDetector (B200 build): inference needs a CUDA tensor; "
                               "there is no CPU fallback (move model and input to cuda)")
        eng = self._engine(x.device)
        x_flat = x.contiguous().view(batch_size * num_frames, c, h, w)        # :115
        if x_flat.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x_flat = x_flat.float()
        if x_flat.dtype != torch.float32 and x_flat.dtype != {"fp16": torch.float16, "bf16": torch.bfloat16}[self.precision]:
            x_flat = x_flat.float()
        offsets = torch.arange(0, (batch_size + 1) * num_frames, num_frames, dtype=torch.int32, device=x.device)
        if self.backbone_name == "resnet50":
            logits, scores = eng.score(x_flat.float(), offsets, num_frames, self.use_temporal_attention)
            return logits, scores.view(batch_size, num_frames)
        logits, scores = eng.score(x_flat, offsets, self.use_temporal_attention)
        return logits, scores.view(batch_size, num_frames)

    def _forward_eager(self, x):
        """Training-mode path (autograd): same graph as the reference's forward, stock PyTorch ops."""
        b, t, c, h, w = x.shape
        feats = self.backbone(x.view(b * t, c, h, w)).view(b, t, -1)
        if self.use_temporal_attention:
            a = F.softmax(self.temporal_attention(feats).squeeze(-1), dim=1)
            pooled, scores = (feats * a.unsqueeze(-1)).sum(dim=1), a
        else:
            pooled, scores = feats.mean(dim=1), torch.ones(b, t, device=x.device) / t
        y = self.dropout(F.relu(self.fc1(self.dropout(pooled))))
        return self.fc2(y), scores


class EnsembleDetector(nn.Module):
    """Reference: src/pretrained_detector.py:146-218.  Members: `efficientnet_b0` and `resnet50` (the reference's default
    pair, app.py:1597)."""

    def __init__(self, backbone_names: List[str], pretrained: bool = True, num_classes: int = 2,
                 dropout_rate: float = 0.5, ensemble_method: str = "average", precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.models = nn.ModuleList([
            PretrainedBackboneDetector(backbone_name=n, pretrained=pretrained, num_classes=num_classes,
                                       dropout_rate=dropout_rate, use_temporal_attention=True, precision=precision)
            for n in backbone_names])
        self.ensemble_method = ensemble_method
        if ensemble_method == "weighted":
            self.weights = nn.Parameter(torch.ones(len(backbone_names)) / len(backbone_names))

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        outs = [m(x) for m in self.models]
        logits = torch.stack([o[0] for o in outs], dim=0)
        scores = torch.stack([o[1] for o in outs], dim=0)
        if self.ensemble_method == "average":
            return logits.mean(dim=0), scores.mean(dim=0)
        if self.ensemble_method == "weighted":
            wts = F.softmax(self.weights, dim=0).view(-1, 1, 1)
            return (logits * wts).sum(dim=0), (scores * wts).sum(dim=0)
        if self.ensemble_method == "voting":
            pred = torch.mode(logits.argmax(dim=-1), dim=0)[0]
            return F.one_hot(pred, num_classes=2).float(), scores.mean(dim=0)
        raise ValueError(f"Unknown ensemble method: {self.ensemble_method}")
