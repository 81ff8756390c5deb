"""Batch scorer for the reference's on-disk face-crop format.

    python -m deepfake_video_detection_b200.score_npz --data_dir crops/ --checkpoint checkpoint_best_efficientnet_b0.pt \\
           --out_csv preds.csv [--max_frames 32] [--threshold 0.5] [--batch_videos 64]

Input: `.npz` files written by the reference's `src/data_prepare.py:279-281` — `faces` uint8 (N,224,224,3) RGB and an
optional `label`; label inference from the file name follows `src/dataset.py:43-49`.  Output: the reference's
prediction CSV (`file,label,pred,prob`, `src/evaluate.py:469-475`; the reference has this CLI only for its gcn/rnn
models).  Videos are batched, uploaded as uint8 and scored by the CUDA path (`FrameScorer.score_host`)."""
from __future__ import annotations

import argparse
import csv
import os
import sys
from pathlib import Path

import numpy as np
import torch


def infer_label(name: str) -> int:
    s = name.lower()                       # dataset.py:43-49
    if "fake" in s or "deepfake" in s:
        return 1
    if "real" in s or "original" in s:
        return 0
    return -1


def load_video(path: Path, max_frames: int):
    data = np.load(path)
    faces = data["faces"]
    if faces.ndim != 4 or faces.shape[-1] != 3 or faces.dtype != np.uint8:
        raise ValueError(f"{path}: `faces` must be uint8 (N,H,W,3), got {faces.dtype} {faces.shape}")
    label = int(np.array(data["label"]).item()) if "label" in data else infer_label(path.name)
    if len(faces) > max_frames:            # evenly spaced subset, like the reference's linspace sampling (evaluate.py:74-83)
        idx = np.linspace(0, len(faces) - 1, max_frames).round().astype(int)
        faces = faces[idx]
    return np.ascontiguousarray(faces), label


def load_state_dict(path: str) -> dict:
    ckpt = torch.load(path, map_location="cpu")
    if isinstance(ckpt, dict):
        for k in ("model_state", "state_dict"):      # app.py:1337
            if k in ckpt and isinstance(ckpt[k], dict):
                return ckpt[k]
    return ckpt


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--data_dir", required=True)
    ap.add_argument("--checkpoint", required=True, help="raw state_dict / {'model_state': ...} .pt file")
    ap.add_argument("--out_csv", default=None)
    ap.add_argument("--max_frames", type=int, default=32)
    ap.add_argument("--batch_videos", type=int, default=64)
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args(argv)

    from . import FrameScorer, decide
    files = sorted(Path(a.data_dir).rglob("*.npz"))
    if not files:
        print(f"no .npz files under {a.data_dir}", file=sys.stderr)
        return 1
    scorer = FrameScorer(load_state_dict(a.checkpoint), a.precision, a.device)
    rows = []
    for b0 in range(0, len(files), a.batch_videos):
        batch = files[b0:b0 + a.batch_videos]
        vids = [load_video(p, max(1, min(64, a.max_frames))) for p in batch]      # app.py:2053 clamps to 1..64
        lens = [len(f) for f, _ in vids]
        host = torch.from_numpy(np.concatenate([f for f, _ in vids])).pin_memory()
        logits, _ = scorer.score_host(host, lens)
        for p, (_, label), d in zip(batch, vids, decide(logits, threshold=a.threshold, abstain_conf=0.0)):
            rows.append([os.path.relpath(p, a.data_dir), label, int(d["is_fake"]), f"{d['prob_fake']:.6f}"])
    if a.out_csv:
        with open(a.out_csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["file", "label", "pred", "prob"])
            w.writerows(rows)
    labelled = [(r[1], r[2]) for r in rows if r[1] in (0, 1)]
    if labelled:
        acc = sum(int(l == p) for l, p in labelled) / len(labelled)
        print(f"{len(rows)} videos scored, accuracy on {len(labelled)} labelled: {acc:.4f}")
    else:
        print(f"{len(rows)} videos scored")
    return 0


if __name__ == "__main__":
    sys.exit(main())
