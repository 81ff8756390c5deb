"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the `resnet50` ensemble member.

Reference: `src/pretrained_detector.py:38-41` builds the trunk as `nn.Sequential(*list(torchvision.models.resnet50().children())[:-1])`
(conv1, bn1, relu, maxpool, layer1..4, avgpool -> (N, 2048, 1, 1)), `:103-143` pools over time and classifies, and
`EnsembleDetector.forward` (`:179-218`) combines the members.  torchvision is installed, so the UNMODIFIED reference
class runs as ground truth (oracle/make_golden_resnet.py); this module restates the trunk functionally from the
state_dict (keys `backbone.{0,1,4,5,6,7}.*`) so that the 16-bit storage points of the CUDA path can be emulated.
Only tests/ may import this module."""
import torch
import torch.nn.functional as F

LAYERS = ((4, 3, 64, 1), (5, 4, 128, 2), (6, 6, 256, 2), (7, 3, 512, 2))        # (Sequential index, blocks, width, stride)


def _r16(t, dt):
    return t if dt is None else t.to(dt).float()


def _conv_bn(sd, conv, bn, x, stride, pad, dt=None):
    """conv (no bias) + BatchNorm (eval) with the BN folded into the weights in fp32, as the CUDA packer does."""
    w = sd[conv + ".weight"]
    scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
    shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
    return F.conv2d(_r16(x, dt), _r16(w * scale.view(-1, 1, 1, 1), dt), shift, stride, pad)


def trunk_features(sd, x, dt=None):
    """x (N,3,H,W) fp32 -> (N,2048) fp32.  dt = torch.float16 / bfloat16 emulates the 16-bit storage of every activation."""
    y = _r16(F.relu(_conv_bn(sd, "backbone.0", "backbone.1", x, 2, 3, dt)), dt)
    y = F.max_pool2d(y, 3, 2, 1)
    for idx, blocks, width, stride in LAYERS:
        for b in range(blocks):
            p = f"backbone.{idx}.{b}."
            s = stride if b == 0 else 1
            o = _r16(F.relu(_conv_bn(sd, p + "conv1", p + "bn1", y, 1, 0, dt)), dt)
            o = _r16(F.relu(_conv_bn(sd, p + "conv2", p + "bn2", o, s, 1, dt)), dt)      # torchvision: stride on the 3x3
            o = _conv_bn(sd, p + "conv3", p + "bn3", o, 1, 0, dt)
            idn = _r16(_conv_bn(sd, p + "downsample.0", p + "downsample.1", y, s, 0, dt), dt) if (p + "downsample.0.weight") in sd else y
            y = _r16(F.relu(o + idn), dt)
    return y.mean((2, 3))


def pool_head(sd, feats, use_attention=True):
    """feats (B,T,D) -> logits (B,2), frame_scores (B,T): pretrained_detector.py:123-141."""
    if use_attention:
        a = torch.sigmoid(F.linear(F.relu(F.linear(feats, sd["temporal_attention.0.weight"], sd["temporal_attention.0.bias"])),
                                   sd["temporal_attention.2.weight"], sd["temporal_attention.2.bias"])).squeeze(-1)
        a = F.softmax(a, dim=1)
        pooled = (feats * a.unsqueeze(-1)).sum(dim=1)
    else:
        a = torch.ones(feats.shape[:2]) / feats.shape[1]
        pooled = feats.mean(dim=1)
    return F.linear(F.relu(F.linear(pooled, sd["fc1.weight"], sd["fc1.bias"])), sd["fc2.weight"], sd["fc2.bias"]), a


def synth_state_dict(seed=0, calib_frames=None, frozen=None):
    """Seeded resnet50 member with the reference schema (326 tensors).  Conv / linear weights and BN affine parameters are
    seeded draws (bit-reproducible); the data-dependent parts — BN running statistics = perturbed batch statistics of a
    calibration batch (activations stay O(1)), fc2 rescaled to a logit margin of std ~1 — are computed once from
    `calib_frames` by oracle/make_golden_resnet.py and frozen in tests/golden/resnet50_ref_seed0.npz (`frozen`), because conv
    outputs are not bit-identical across CPUs.  bn3 gains are small (x0.25), as in trained residual networks: with unit
    gains the random trunk amplifies a 16-bit rounding ~20x and no 16-bit implementation could meet 2e-2."""
    import torchvision
    torch.manual_seed(seed)
    net = torchvision.models.resnet50(weights=None)
    g = torch.Generator().manual_seed(seed + 1)
    trunk = torch.nn.Sequential(*list(net.children())[:-1])
    bns = [(n, m) for n, m in trunk.named_modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for n, m in bns:                                          # affine parameters first: the statistics below depend on them
            gain = 0.6 + 0.5 * torch.rand(m.weight.shape, generator=g)
            m.weight.copy_(gain * (0.25 if n.endswith("bn3") else 1.0))
            m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
    if frozen is not None:
        with torch.no_grad():
            for n, m in bns:
                m.running_mean.copy_(torch.from_numpy(frozen["backbone." + n + ".running_mean"]))
                m.running_var.copy_(torch.from_numpy(frozen["backbone." + n + ".running_var"]))
    elif calib_frames is not None:
        for _, m in bns:
            m.momentum = 1.0                                      # running stats := batch stats of the calibration batch
        trunk.train()
        with torch.no_grad():
            trunk(calib_frames)
        trunk.eval()
        with torch.no_grad():
            for _, m in bns:
                m.running_mean += 0.1 * m.running_var.sqrt() * torch.randn(m.running_mean.shape, generator=g)
                m.running_var.mul_(0.8 + 0.45 * torch.rand(m.running_var.shape, generator=g)).add_(1e-4)
    trunk.eval()
    sd = {"backbone." + k: v.detach().clone() for k, v in trunk.state_dict().items()}
    g2 = torch.Generator().manual_seed(seed + 2)
    r = lambda *s, std: torch.randn(*s, generator=g2) * std
    sd.update({"temporal_attention.0.weight": r(64, 2048, std=0.03), "temporal_attention.0.bias": r(64, std=0.1),
               "temporal_attention.2.weight": r(1, 64, std=0.6), "temporal_attention.2.bias": r(1, std=0.1),
               "fc1.weight": r(256, 2048, std=0.03), "fc1.bias": r(256, std=0.1),
               "fc2.weight": r(2, 256, std=0.1), "fc2.bias": r(2, std=0.1)})
    if frozen is not None:
        sd["fc2.weight"] = sd["fc2.weight"] * float(frozen["__fc2_scale__"])
        sd["fc2.bias"] = torch.from_numpy(frozen["fc2.bias"]).clone()
    return sd
