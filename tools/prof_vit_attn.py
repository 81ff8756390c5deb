"""Time the ViT attention kernel (tcgen05 / TMEM, csrc/vit_attn_tc.cu) on random qkv through the kernel-level C ABI."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--images", type=int, default=512); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
lib = _lib.load()
qkv = (torch.randn(a.images * 197, 2304, device="cuda") * 1.5).half()
o = torch.empty(a.images * 197, 768, device="cuda", dtype=torch.half)
st = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.dfd_k_vit_attention(qkv.data_ptr(), o.data_ptr(), a.images, 1, st))
run(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
flops = 4.0 * 197 * 197 * 64 * 12 * a.images
print(f"attention (tcgen05), {a.images} images: best {min(ts)*1e3:.1f} us  {flops / min(ts) / 1e9:.0f} TFLOP/s")
