"""BASELINE config 4 (single GPU part): frames/s of the scoring path vs batch size F (videos of 32 crops; F<32 -> one short video),
device-resident inputs, plus the end-to-end host path for a few chunk sizes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepfake_video_detection_b200 import FrameScorer, make_offsets
from deepfake_video_detection_b200.synthetic import load_checkpoint

sc = FrameScorer(load_checkpoint(0), "fp16", "cuda")
res = {}
for F in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    lens = [32] * (F // 32) if F >= 32 else [F]
    crops = torch.randint(0, 256, (F, 224, 224, 3), dtype=torch.uint8, device="cuda")
    off = make_offsets(lens, "cuda")
    for _ in range(3): sc.score(crops, off)
    torch.cuda.synchronize()
    n = 20 if F <= 256 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): sc.score(crops, off)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[F] = {"ms": round(ms, 3), "frames_per_s": round(F / ms * 1e3)}
    if F <= 256:                                   # same step as one CUDA-graph launch
        gs = sc.capture(lens)
        for _ in range(3): gs.run(crops)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): gs.run(crops)
        e1.record(); torch.cuda.synchronize()
        msg = e0.elapsed_time(e1) / n
        res[F].update(graph_ms=round(msg, 3), graph_frames_per_s=round(F / msg * 1e3))
        del gs
    print(F, res[F], flush=True)
    del crops
host = torch.empty((2048, 224, 224, 3), dtype=torch.uint8, pin_memory=True).random_(0, 256)
for cv in (4, 8, 16, 32, 64):
    for _ in range(2): sc.score_host(host, [32] * 64, chunk_videos=cv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sc.score_host(host, [32] * 64, chunk_videos=cv)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("e2e chunk_videos", cv, round(ms, 2), "ms", round(2048 / ms * 1e3), "frames/s", flush=True)
    res[f"e2e_chunk{cv}"] = round(2048 / ms * 1e3)
json.dump(res, open("gpurun_out/sweep_batch.json", "w"), indent=1)
