"""Sweep the depthwise kernel's channel block (channels per CTA) over the 12 network shapes in ONE process.
    python tools/sweep_dw.py [--frames 1024]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib

SHAPES = {(32, 3, 1, 112): [16, 32], (96, 3, 2, 112): [16, 32, 48], (144, 3, 1, 56): [16, 48], (144, 5, 2, 56): [16, 48, 72],
          (240, 5, 1, 28): [16, 48, 80, 120], (240, 3, 2, 28): [48, 80, 120, 240], (480, 3, 1, 14): [32, 96, 160, 240],
          (480, 5, 1, 14): [32, 96, 160, 240], (672, 5, 1, 14): [32, 96, 112, 224], (672, 5, 2, 14): [96, 224, 336],
          (1152, 5, 1, 7): [64, 128, 192, 384], (1152, 3, 1, 7): [64, 128, 192, 384]}
ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=1024); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
total = 0.0
for (C, k, s, H), cbs in SHAPES.items():
    OH = (H + 2 * (k // 2) - k) // s + 1
    x = torch.randn(a.frames, H, H, C, device="cuda").half()
    w = torch.randn(k * k, C, device="cuda"); b = torch.randn(C, device="cuda")
    out = torch.empty(a.frames, OH, OH, C, device="cuda", dtype=torch.half)
    parts = torch.empty(a.frames, lib.dfd_k_dw_num_partials(OH, OH, C, k, s), C, device="cuda")
    nbytes = (x.numel() + out.numel()) * 2
    res = {}
    for cb in [0] + cbs:
        lib.dfd_k_set_dw_channel_block(cb)
        def run():
            _lib.check(lib.dfd_k_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), parts.data_ptr(), a.frames, H, H, C, k, s, 1, st))
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[cb] = min(ts) * 1e3
    lib.dfd_k_set_dw_channel_block(0)
    best = min(res, key=res.get)
    total += res[0]
    print(f"C={C:5d} k={k} s={s} H={H:4d}: " + "  ".join(f"cb{cb}={t:.0f}us" for cb, t in res.items()) + f"   best cb{best} {nbytes / res[best] / 1e3:.0f} GB/s (default {nbytes / res[0] / 1e3:.0f})")
    del x, out, parts
print(f"default total (16 launches of the network): {total:.0f} us per pass of the 12 shapes")
