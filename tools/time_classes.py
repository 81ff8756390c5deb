"""Per-kernel-class device time of one scoring step (CUDA events around every launch, dfd_profile_*), median of N profiled passes.
Quick A/B of library variants: DFD_LIB_PATH=build/variants/libdfd_<name>.so python tools/time_classes.py [--videos 64 --frames 32]"""
import argparse, ctypes as C, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import FrameScorer, make_offsets, _lib
from deepfake_video_detection_b200.synthetic import load_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=64)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--only", default="")
a = ap.parse_args()
lib = _lib.load()
sc = FrameScorer(load_checkpoint(0), "fp16", "cuda")
F = a.videos * a.frames
crops = torch.randint(0, 256, (F, 224, 224, 3), dtype=torch.uint8, device="cuda")
off = make_offsets([a.frames] * a.videos, "cuda")
for _ in range(3): sc.score(crops, off)
torch.cuda.synchronize()
acc = {}
for _ in range(a.iters):
    lib.dfd_profile_enable(1)
    sc.score(crops, off)
    entries = (_lib.ProfileEntry * 16)(); n = C.c_int()
    _lib.check(lib.dfd_profile_collect(entries, 16, C.byref(n)), "profile_collect")
    lib.dfd_profile_enable(0)
    for e in entries[: n.value]:
        if e.launches: acc.setdefault(e.name.decode(), []).append((e.ms, e.bytes))
tot = 0.0
out = []
for k, v in acc.items():
    ms = statistics.median(x[0] for x in v); tot += ms
    if not a.only or a.only in k: out.append(f"{k} {ms:.4f} ms {v[0][1] / ms / 1e6:.0f} GB/s")
print(os.environ.get("DFD_LIB_PATH", "default"), "|", " | ".join(out), "| sum", round(tot, 3))
