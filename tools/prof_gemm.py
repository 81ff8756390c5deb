"""Time / profile one pointwise-GEMM shape through the kernel-level C ABI."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=144); ap.add_argument("--N", type=int, default=24)
ap.add_argument("--HW", type=int, default=3136); ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--gate", type=int, default=1); ap.add_argument("--res", type=int, default=1); ap.add_argument("--act", type=int, default=0)
ap.add_argument("--iters", type=int, default=5); ap.add_argument("--pool", type=int, default=0)
a = ap.parse_args()
lib = _lib.load()
M = a.frames * a.HW
A = torch.randn(M, a.K, device="cuda").half(); W = (torch.randn(a.N, a.K, device="cuda") / a.K ** 0.5).half()
bias = torch.randn(a.N, device="cuda"); G = torch.rand(a.frames, a.K, device="cuda") if a.gate else None
R = torch.randn(M, a.N, device="cuda").half() if a.res else None
D = torch.empty(M, a.N, device="cuda", dtype=torch.half)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
feat = torch.empty(a.frames, a.N, device="cuda")
def run():
    if a.pool:
        _lib.check(lib.dfd_k_gemm_pool(A.data_ptr(), W.data_ptr(), bias.data_ptr(), feat.data_ptr(), M, a.K, a.N, a.HW, 1, 0, st))
        return
    _lib.check(lib.dfd_k_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), G.data_ptr() if a.gate else None, R.data_ptr() if a.res else None,
                              D.data_ptr(), M, a.K, a.N, a.HW, a.act, 1, 0, st))
run(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
nbytes = M * (a.K + a.N + (a.N if a.res else 0)) * 2
flops = 2.0 * M * a.K * a.N
print(f"K={a.K} N={a.N} HW={a.HW} frames={a.frames} gate={a.gate} res={a.res} act={a.act}: best {min(ts)*1e3:.1f} us  {nbytes / min(ts) / 1e6:.0f} GB/s  {flops / min(ts) / 1e9:.0f} TFLOP/s")
