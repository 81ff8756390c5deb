"""Host->device bandwidth of the bench's per-step input (308 MB of uint8 crops from pinned memory), alone and while the scoring
kernels run: tells whether bench.py's `e2e` (H2D inside the timed region) is bound by the copy or by the kernels."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import FrameScorer, make_offsets
from deepfake_video_detection_b200.synthetic import load_checkpoint

F = 2048
host = torch.empty((F, 224, 224, 3), dtype=torch.uint8, pin_memory=True).random_(0, 256)
dev = torch.empty_like(host, device="cuda")
cs = torch.cuda.Stream()
def copy_ms(n=5):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(cs):
            e0.record(cs); dev.copy_(host, non_blocking=True); e1.record(cs)
        cs.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
alone = copy_ms()
sc = FrameScorer(load_checkpoint(0), device="cuda")
crops = torch.randint(0, 256, (F, 224, 224, 3), dtype=torch.uint8, device="cuda")
off = make_offsets([32] * 64, "cuda")
for _ in range(3):
    sc.score(crops, off)
torch.cuda.synchronize()
for _ in range(12):                      # keep the GPU busy while the copies are timed
    sc.score(crops, off)
busy = copy_ms()
torch.cuda.synchronize()
print(json.dumps({"bytes": host.numel(), "h2d_ms_alone": round(alone, 3), "h2d_GBps_alone": round(host.numel() / alone / 1e6, 1),
                  "h2d_ms_under_load": round(busy, 3), "h2d_GBps_under_load": round(host.numel() / busy / 1e6, 1)}))
