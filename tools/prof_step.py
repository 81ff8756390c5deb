"""Profiling driver: N forward passes of the scoring path on synthetic crops (for ncu / quick timing)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import FrameScorer, make_offsets
from deepfake_video_detection_b200.synthetic import load_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=16)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="fp16")
a = ap.parse_args()
sc = FrameScorer(load_checkpoint(0), a.precision, "cuda")
F = a.videos * a.frames
crops = torch.randint(0, 256, (F, 224, 224, 3), dtype=torch.uint8, device="cuda")
off = make_offsets([a.frames] * a.videos, "cuda")
for i in range(a.iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc.score(crops, off)
    torch.cuda.synchronize(); print(f"iter {i}: {1e3 * (time.perf_counter() - t0):.3f} ms, {sc.last_launch_count} launches, {F / (time.perf_counter() - t0):.0f} frames/s")
