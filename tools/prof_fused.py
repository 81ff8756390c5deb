"""Time / profile one fused expand + depthwise block (mbconv_fused.cu) through the kernel-level C ABI."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib

SHAPES = {1: (16, 96, 112, 3, 2), 2: (24, 144, 56, 3, 1), 3: (24, 144, 56, 5, 2)}
ap = argparse.ArgumentParser()
ap.add_argument("--block", type=int, default=1); ap.add_argument("--frames", type=int, default=512); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
cin, mid, H, k, s = SHAPES[a.block]
lib = _lib.load()
OH = (H + 2 * (k // 2) - k) // s + 1
x = torch.randn(a.frames, H, H, cin, device="cuda").half()
we = (torch.randn(mid, cin, device="cuda") / cin ** 0.5).half(); be = torch.randn(mid, device="cuda") * 0.3
w = (torch.randn(k * k, mid, device="cuda") / k).contiguous(); b = torch.randn(mid, device="cuda") * 0.2
out = torch.empty(a.frames, OH, OH, mid, device="cuda", dtype=torch.half)
parts = torch.empty(a.frames, lib.dfd_k_dw_num_partials(OH, OH, mid, k, s), mid, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.dfd_k_mbconv_fused(x.data_ptr(), we.data_ptr(), be.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), parts.data_ptr(),
                                      a.frames, H, H, cin, mid, k, s, 1, st))
run(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
nbytes = a.frames * (H * H * cin + OH * OH * mid) * 2
mufu = a.frames * (H * H * mid + OH * OH * mid)
print(f"block {a.block} ({cin}->{mid} @{H} k{k} s{s}) frames={a.frames}: best {min(ts)*1e3:.1f} us  {nbytes / min(ts) / 1e6:.0f} GB/s algorithmic, "
      f"MUFU floor {mufu / (148 * 16 * 1.965e9) * 1e6:.1f} us")
