"""GPU diagnostic (writes the record the large parity test asserts on): error of the CUDA path against the fp32 oracle over
the full BASELINE configs[1] batch (64 videos x 32 crops) and 256 ragged videos with T in 1..64.

    python tools/parity_diag.py [--precision fp16] [--out gpurun_out/parity_diag.json]

The oracle is the checker here (tools/ is not product); nothing in the package imports it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL = 2e-2


def ragged_lengths():
    """256 videos: every T of 1..64 occurs, short videos (no averaging over T, the worst case) over-represented."""
    lens = [1] * 64 + [2] * 32 + [3] * 16 + [4] * 16
    lens += [5 + (i * 7) % 12 for i in range(64)]
    lens += [17 + (i * 5) % 16 for i in range(48)]
    lens += [33 + (i * 11) % 32 for i in range(15)] + [64]
    return lens


def sample_sets():
    from deepfake_video_detection_b200.synthetic import synth_crops
    return {"c2_64x32": synth_crops(11, 64, 32), "ragged_256": synth_crops(12, 256, ragged_lengths())}


def compare(logits, ref_logits, lens, tol=TOL):
    """Per-video record: max-abs logit error, verdict agreement outside the |prob_fake - 0.5| < 2 tol band."""
    from oracle import effnet_b0_oracle as O
    dl = (logits - ref_logits).abs().max(dim=1)[0]
    ours, ref = O.decide(logits), O.decide(ref_logits)
    in_band = [abs(r["prob_fake"] - r["threshold"]) < 2 * tol for r in ref]
    flips = [o["is_fake"] != r["is_fake"] for o, r in zip(ours, ref)]
    lens = np.asarray(lens)
    by_T = {}
    for lo, hi in ((1, 1), (2, 2), (3, 4), (5, 16), (17, 32), (33, 64)):
        m = torch.from_numpy((lens >= lo) & (lens <= hi))
        if m.any():
            by_T[f"T{lo}-{hi}"] = {"videos": int(m.sum()), "max": float(dl[m].max()), "rms": float(dl[m].pow(2).mean().sqrt())}
    return {"videos": len(lens), "frames": int(lens.sum()), "dlogit_max": float(dl.max()), "dlogit_rms": float(dl.pow(2).mean().sqrt()),
            "dlogit_p99": float(dl.quantile(0.99)), "over_tol": int((dl > tol).sum()), "by_T": by_T,
            "flips_outside_band": int(sum(f and not b for f, b in zip(flips, in_band))), "flips_in_band": int(sum(f and b for f, b in zip(flips, in_band))),
            "videos_in_band": int(sum(in_band)), "ref_logit_absmax": float(ref_logits.abs().max()),
            "ref_margin_std": float((ref_logits[:, 1] - ref_logits[:, 0]).std())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16,bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_diag.json"))
    args = ap.parse_args()
    from deepfake_video_detection_b200 import FrameScorer, make_offsets
    from deepfake_video_detection_b200.synthetic import load_checkpoint
    from oracle import effnet_b0_oracle as O
    torch.set_num_threads(os.cpu_count() or 8)
    sd = load_checkpoint(0)
    out = {"tol": TOL, "switches": {k: v for k, v in os.environ.items() if k.startswith("DFD_")}}
    sets = sample_sets()
    refs = {}
    for name, (crops, offs) in sets.items():
        t0 = time.time()
        refs[name] = O.score_ragged_batched(sd, crops, offs)
        print(f"oracle {name}: {int(offs[-1])} frames in {time.time() - t0:.1f} s", flush=True)
    for prec in args.precision.split(","):
        sc = FrameScorer(sd, prec, "cuda")
        out[prec] = {}
        for name, (crops, offs) in sets.items():
            lens = np.diff(offs)
            logits, scores, feat = sc.score(torch.from_numpy(crops).cuda(), make_offsets(lens, "cuda"), return_features=True)
            ref_logits, ref_scores, ref_feat = refs[name]
            rec = compare(logits.cpu(), ref_logits, lens)
            rec["feat_rel"] = float((feat.cpu() - ref_feat).norm() / ref_feat.norm())
            rec["frame_scores_max_abs"] = float((scores.cpu() - ref_scores).abs().max())
            out[prec][name] = rec
            print(prec, name, json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
