"""Time the CTA-pair GEMM (csrc/gemm_pair.cu) against the single-CTA kernel (csrc/gemm_tc.cu) on the ViT-B/16 shapes."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfake_video_detection_b200 import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--images", type=int, default=512); ap.add_argument("--iters", type=int, default=10); ap.add_argument("--impls", type=int, nargs="+", default=[3, 0])
a = ap.parse_args()
lib = _lib.load()
M = a.images * 197
st = int(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, K, N, act in [("qkv", 768, 2304, 0), ("proj", 768, 768, 0), ("fc1", 768, 3072, 2), ("fc2", 3072, 768, 0)]:
    A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).half(); b = torch.randn(N, device="cuda")
    D = torch.empty(M, N, dtype=torch.float16, device="cuda")
    for impl in a.impls:
        run = lambda: _lib.check(lib.dfd_k_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr(), None, None, D.data_ptr(), M, K, N, 1, act, 1, impl, st))
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"{name:5s} M {M} K {K} N {N} impl {impl} ({'CTA pairs' if impl == 3 else 'single CTA'}): median {t*1e3:.1f} us  {2.0*M*K*N/t/1e9:.0f} TFLOP/s", flush=True)
