"""Time / profile one depthwise shape through the kernel-level C ABI."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=144); ap.add_argument("--k", type=int, default=3); ap.add_argument("--s", type=int, default=1)
ap.add_argument("--H", type=int, default=56); ap.add_argument("--frames", type=int, default=512); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
lib = _lib.load()
OH = (a.H + 2 * (a.k // 2) - a.k) // a.s + 1
x = torch.randn(a.frames, a.H, a.H, a.C, device="cuda").half()
w = torch.randn(a.k * a.k, a.C, device="cuda"); b = torch.randn(a.C, device="cuda")
out = torch.empty(a.frames, OH, OH, a.C, device="cuda", dtype=torch.half)
nparts = lib.dfd_k_dw_num_partials(OH, OH, a.C, a.k, a.s)
parts = torch.empty(a.frames, nparts, a.C, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.dfd_k_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), parts.data_ptr(), a.frames, a.H, a.H, a.C, a.k, a.s, 1, st))
run(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
nbytes = (x.numel() + out.numel()) * 2
print(f"dw C={a.C} k={a.k} s={a.s} H={a.H} frames={a.frames}: best {min(ts)*1e3:.1f} us  {nbytes / min(ts) / 1e6:.0f} GB/s  ({out.numel() / min(ts) / 1e6:.1f} Gelem/s)")
