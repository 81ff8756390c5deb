"""BASELINE config 3: EfficientNet-B0 features + LogicRNNLSTM temporal head over 16-frame sequences, batch 256 videos."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import FrameScorer
from deepfake_video_detection_b200.rnn_model import LogicRNNLSTM
from deepfake_video_detection_b200.synthetic import load_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=256); ap.add_argument("--frames", type=int, default=16); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
torch.manual_seed(0)
sc = FrameScorer(load_checkpoint(0), "fp16", "cuda")
rnn = LogicRNNLSTM(1280, 512, 2).eval().cuda()
F = a.videos * a.frames
crops = torch.randint(0, 256, (F, 224, 224, 3), dtype=torch.uint8, device="cuda")
def step():
    feats = sc.features(crops)                                    # (F, 1280) fp32
    return rnn(feats.view(a.videos, a.frames, 1280)), feats
with torch.no_grad():
    for _ in range(2): step()
    torch.cuda.synchronize()
    t_all, t_rnn = [], []
    for _ in range(a.iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); feats = sc.features(crops); e1.record(); rnn(feats.view(a.videos, a.frames, 1280)); e2.record()
        torch.cuda.synchronize(); t_all.append(e0.elapsed_time(e2)); t_rnn.append(e1.elapsed_time(e2))
ms, ms_rnn = sorted(t_all)[len(t_all) // 2], sorted(t_rnn)[len(t_rnn) // 2]
print(json.dumps({"workload": f"EfficientNet-B0 features + LogicRNNLSTM head, {a.videos} videos x {a.frames} frames", "ms": round(ms, 3),
                  "ms_rnn_head": round(ms_rnn, 3), "frames_per_s": round(F / ms * 1e3, 1), "videos_per_s": round(a.videos / ms * 1e3, 1)}))
