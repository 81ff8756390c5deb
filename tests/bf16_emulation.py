"""CPU emulation of the CUDA path's rounding points (bf16 storage, fp32 accumulate).

Test helper only: predicts how far a *correct* bf16 pipeline may sit from the fp32 oracle, and gives
the per-kernel GPU tests a tight expectation (same inputs, same rounding points).
"""
import torch
import torch.nn.functional as F

from oracle import effnet_b0_oracle as O


def bf(x):
    return x.to(torch.bfloat16).float()


def trunk_features_bf16(sd, x, taps=None, gate_round=True):
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t
    x = bf(x)                                              # K1 output is bf16
    w, b = O.fold_bn(sd["backbone.0.weight"], sd, "backbone.1")
    y = tap("stem", bf(F.silu(F.conv2d(x, bf(w), b, 2, 1))))
    for (p, cin, mid, cout, k, stride, rd, has_expand, has_skip) in O.block_specs():
        inp = y
        if has_expand:
            w, b = O.fold_bn(sd[p + ".conv_pw.weight"], sd, p + ".bn1")
            y = tap(p + ".expand", bf(F.silu(F.conv2d(y, bf(w), b))))
            w, b = O.fold_bn(sd[p + ".conv_dw.weight"], sd, p + ".bn2")
        else:
            w, b = O.fold_bn(sd[p + ".conv_dw.weight"], sd, p + ".bn1")
        d32 = F.silu(F.conv2d(y, w, b, stride, k // 2, 1, mid))        # dw weights stay fp32
        s = d32.mean((2, 3), keepdim=True)                                # squeeze from fp32 values
        y = tap(p + ".dw", bf(d32))
        s = F.silu(F.conv2d(s, sd[p + ".se.conv_reduce.weight"], sd[p + ".se.conv_reduce.bias"]))
        g = torch.sigmoid(F.conv2d(s, sd[p + ".se.conv_expand.weight"], sd[p + ".se.conv_expand.bias"]))
        tap(p + ".gate", g)
        a = y * g
        if gate_round:
            a = bf(a)                                                     # gated A operand re-rounded for the MMA
        if has_expand:
            w, b = O.fold_bn(sd[p + ".conv_pwl.weight"], sd, p + ".bn3")
        else:
            w, b = O.fold_bn(sd[p + ".conv_pw.weight"], sd, p + ".bn2")
        y = F.conv2d(a, bf(w), b)
        if has_skip:
            y = y + inp
        y = tap(p + ".out", bf(y))
    w, b = O.fold_bn(sd["backbone.3.weight"], sd, "backbone.4")
    y = F.silu(F.conv2d(y, bf(w), b))                                     # head stays fp32 into the pool
    return y.mean((2, 3))
