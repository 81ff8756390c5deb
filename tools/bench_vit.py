"""ViT-B/16 frame encoder at BASELINE config 5 (batch 512): device time, TFLOP/s against the measured bf16 peak."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200.vit_model import ViTFeatureExtractor

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512); ap.add_argument("--iters", type=int, default=5); ap.add_argument("--precision", default="fp16")
a = ap.parse_args()
torch.manual_seed(0)
m = ViTFeatureExtractor(precision=a.precision).eval().cuda()
x = torch.randn(a.batch, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
flops = 35.1e9 * a.batch                       # 17.56 GMAC per image
peak = 1387.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", peak)
except Exception:
    pass
print(json.dumps({"workload": f"ViT-B/16 forward, batch {a.batch}, {a.precision}", "ms": round(ms, 3), "images_per_s": round(a.batch / ms * 1e3, 1),
                  "TFLOPs": round(flops / ms / 1e9, 1), "peak_TFLOPs": peak, "frac": round(flops / ms / 1e9 / peak, 4)}))
