"""CPU study of the 16-bit path's rounding points (not a test, not product).

Emulates the CUDA path's storage roundings on the fp32 oracle graph and ablates one class of rounding
points at a time, to find which of them spend the 2e-2 logit budget.  Worst case = single-frame videos
(no averaging over T in the attention pool), so the sample is N videos of T = 1.

    python tools/numerics_study.py [--n 96] [--seed 5] [--dtype fp16|bf16]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import effnet_b0_oracle as O                                  # noqa: E402
from deepfake_video_detection_b200.synthetic import load_checkpoint, synth_crops    # noqa: E402


def make_round(dtype):
    t = torch.float16 if dtype == "fp16" else torch.bfloat16
    return lambda x: x.to(t).float()


def silu_tanh_approx(x, noise):
    """x * (0.5 * tanh.approx(0.5 x) + 0.5); tanh.approx.f32 has a maximum relative error of 2^-11 (PTX ISA), modelled
    as a uniform relative perturbation of that size when `noise` is set."""
    t = torch.tanh(0.5 * x)
    if noise:
        t = t * (1.0 + (torch.rand_like(t) * 2 - 1) * 2.0 ** -11)
    return x * (0.5 * t + 0.5)


def silu_tanh_f16x2(x):
    """tanh.approx.f16x2 form of the SiLU: h = x/2 rounded to fp16, t = tanh(h) with a maximum ABSOLUTE error of 2^-10.987 (PTX ISA,
    modelled uniform) rounded to fp16, result h + h t in one fp16 FMA."""
    h = (0.5 * x).half().float()
    t = torch.tanh(h) + (torch.rand_like(h) * 2 - 1) * 2.0 ** -10.987
    t = t.half().float()
    return (h + h * t).half().float()


def trunk(sd, x, rnd, cfg):
    """cfg keys (all default True = the shipped path's rounding): w (pointwise weights), stem, expand, dw, gate (gated A
    re-rounded), out (block outputs), skip32 (False; True keeps the skip operand in fp32), silu_noise, head_in."""
    state = {"on": True}
    clean = cfg.get("clean_blocks", ())                    # block indices computed without any 16-bit rounding
    r = lambda key, t: rnd(t) if (cfg.get(key, True) and state["on"]) else t
    noisy = cfg.get("silu_noise", True)
    silu = lambda t: silu_tanh_approx(t, True) if (noisy and state["on"]) else F.silu(t)
    silu_e = silu if cfg.get("silu_noise_expand", True) else F.silu       # SiLU of the stem / expand epilogues
    silu_x = silu_tanh_f16x2 if cfg.get("expand_f16tanh", False) else silu_e   # SiLU of the expand epilogues only
    silu_d = silu if cfg.get("silu_noise_dw", True) else F.silu           # SiLU of the depthwise kernels
    w, b = O.fold_bn(sd["backbone.0.weight"], sd, "backbone.1")
    y = r("stem", silu_e(F.conv2d(x, w, b, 2, 1)))          # uint8 inputs are exact, the stem weights are hi+lo split
    y32 = y
    for bi, (p, cin, mid, cout, k, stride, rd, has_expand, has_skip) in enumerate(O.block_specs()):
        state["on"] = bi not in clean
        inp = y32 if cfg.get("skip32", False) else y
        if has_expand:
            w, b = O.fold_bn(sd[p + ".conv_pw.weight"], sd, p + ".bn1")
            y = r("expand", (silu_x if bi in cfg.get("f16tanh_blocks", range(16)) else silu_e)(F.conv2d(y, r("w", w), b)))
            w, b = O.fold_bn(sd[p + ".conv_dw.weight"], sd, p + ".bn2")
        else:
            w, b = O.fold_bn(sd[p + ".conv_dw.weight"], sd, p + ".bn1")
        d32 = silu_d(F.conv2d(y, w, b, stride, k // 2, 1, mid))
        s = d32.mean((2, 3), keepdim=True)
        y = r("dw", d32)
        s = F.silu(F.conv2d(s, sd[p + ".se.conv_reduce.weight"], sd[p + ".se.conv_reduce.bias"]))
        g = torch.sigmoid(F.conv2d(s, sd[p + ".se.conv_expand.weight"], sd[p + ".se.conv_expand.bias"]))
        a = r("gate", y * g)
        if has_expand:
            w, b = O.fold_bn(sd[p + ".conv_pwl.weight"], sd, p + ".bn3")
        else:
            w, b = O.fold_bn(sd[p + ".conv_pw.weight"], sd, p + ".bn2")
        y32 = F.conv2d(a, r("w", w), b)
        if has_skip:
            y32 = y32 + inp
        y = r("out", y32)
    state["on"] = True
    w, b = O.fold_bn(sd["backbone.3.weight"], sd, "backbone.4")
    y = F.silu(F.conv2d(y, r("w", w), b))
    return y.mean((2, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=96)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    torch.manual_seed(0)
    sd = load_checkpoint(0)
    crops, _ = synth_crops(args.seed, args.n, [1] * args.n)
    rnd = make_round(args.dtype)
    variants = {
        "shipped": {},
        "exact_silu": {"silu_noise": False},
        "exact_silu_dw": {"silu_noise_dw": False},
        "exact_silu_expand": {"silu_noise_expand": False},
        "expand_f16tanh_all": {"expand_f16tanh": True},
        "expand_f16tanh_b123": {"expand_f16tanh": True, "f16tanh_blocks": (1, 2, 3)},
        "no_w": {"w": False},
        "no_expand": {"expand": False},
        "no_dw": {"dw": False},
        "no_gate": {"gate": False},
        "no_out": {"out": False},
        "no_stem": {"stem": False},
        "skip32": {"skip32": True},
        "skip32+exact_silu": {"skip32": True, "silu_noise": False},
        "skip32+exact_silu+no_w": {"skip32": True, "silu_noise": False, "w": False},
        "skip32+exact_silu+no_w+no_gate": {"skip32": True, "silu_noise": False, "w": False, "gate": False},
        "clean_0-4": {"clean_blocks": range(0, 5)},
        "clean_5-10": {"clean_blocks": range(5, 11)},
        "clean_11-15": {"clean_blocks": range(11, 16)},
        "clean_5-15": {"clean_blocks": range(5, 16)},
        "clean_0-2": {"clean_blocks": range(0, 3)},
        "none": {k: False for k in ("w", "stem", "expand", "dw", "gate", "out", "silu_noise")},
    }
    if args.variants:
        variants = {k: v for k, v in variants.items() if k in args.variants.split(",")}
    with torch.no_grad():
        x = O.prep_u8_hwc(crops)
        f32 = O.trunk_features(sd, x)
        ref = O.attention_pool_head(sd, f32[:, None])[0]
        print(f"{args.dtype}: {args.n} single-frame videos, |logit| max {ref.abs().max():.2f}, margin std {(ref[:, 1] - ref[:, 0]).std():.2f}")
        out = {}
        for name, cfg in variants.items():
            f = trunk(sd, x, rnd, cfg)
            lg = O.attention_pool_head(sd, f[:, None])[0]
            dl = (lg - ref).abs().max(dim=1)[0]
            out[name] = dict(feat_rel=float((f - f32).norm() / f32.norm()), dlogit_max=float(dl.max()),
                             dlogit_p90=float(dl.quantile(0.9)), dlogit_rms=float(dl.pow(2).mean().sqrt()))
            print(f"{name:36s} feat_rel {out[name]['feat_rel']:.2e}  dlogit max {out[name]['dlogit_max']:.4f}  p90 {out[name]['dlogit_p90']:.4f}  rms {out[name]['dlogit_rms']:.4f}", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
