"""`resnet50` ensemble member (and the weighted ensemble with efficientnet_b0): device time per batch of videos.
Round 1 took one short run (profiles/r01_bench_resnet.json); a longer one belongs in the next evidence run."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import EnsembleDetector, PretrainedBackboneDetector

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=16); ap.add_argument("--frames", type=int, default=32); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
torch.manual_seed(0)
x = torch.randn(a.videos, a.frames, 3, 224, 224, device="cuda")
out = {}
for name, m in (("resnet50", PretrainedBackboneDetector("resnet50", pretrained=False)),
                ("ensemble_effnet_b0+resnet50", EnsembleDetector(["efficientnet_b0", "resnet50"], pretrained=False, ensemble_method="weighted"))):
    m = m.eval().cuda()
    with torch.no_grad():
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    F = a.videos * a.frames
    out[name] = {"ms": round(ms, 3), "frames_per_s": round(F / ms * 1e3, 1)}
    if name == "resnet50":
        out[name]["TFLOPs"] = round(8.2e9 * F / ms / 1e9, 1)        # 4.1 GMAC per 224x224 frame
print(json.dumps({"workload": f"{a.videos} videos x {a.frames} frames, fp32 NCHW input resident on the device", **out}))
