"""Time the squeeze-excite gate kernel on one layer shape through the kernel-level C ABI."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=1152); ap.add_argument("--rd", type=int, default=48); ap.add_argument("--nparts", type=int, default=1)
ap.add_argument("--frames", type=int, default=2048); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = _lib.load()
parts = torch.randn(a.frames, a.nparts, a.C, device="cuda")
w1 = torch.randn(a.rd, a.C, device="cuda") * 0.1; b1 = torch.randn(a.rd, device="cuda"); w2t = torch.randn(a.rd, a.C, device="cuda") * 0.3; b2 = torch.randn(a.C, device="cuda")
gate = torch.empty(a.frames, a.C, device="cuda")
st = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.dfd_k_se(parts.data_ptr(), a.nparts, C.c_float(1.0 / 49), w1.data_ptr(), b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), gate.data_ptr(), a.frames, a.C, a.rd, st))
run(); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(a.iters):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"se C={a.C} rd={a.rd} nparts={a.nparts} frames={a.frames}: best {min(ts)*1e3:.1f} us")
