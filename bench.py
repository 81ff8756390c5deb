#!/usr/bin/env python
"""Benchmark of the hot path: per-video EfficientNet-B0 real/fake scoring of uint8 face crops.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 64 videos x 32 uint8 224x224 crops per GPU -> fused prep + trunk +
temporal-attention pool + head -> per-video logits.  One "step" = one pass over that batch.  With N > 1 the
driver launches one rank per GPU (torchrun): every rank scores its own 64 videos (weak scaling) and the step
ends with the single all-gather of per-video logits.  Prints ONE JSON line on rank 0.

  value     frames/s, inputs resident in HBM, CUDA events over exactly K steps, max over ranks
  e2e       same metric through the public API with HOST (pinned) crops: H2D copy + score + D2H of logits per step
  roofline  dominant kernel class of one extra profiled step (CUDA events around every launch on the launch
            stream): algorithmic bytes / measured time vs the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle (port of the reference's CPU path) on the host cores, bounded sample
--impl reference: the reference's CPU path (oracle port; the reference itself is pure Python over an
un-vendored timm and cannot travel to the GPU box) timed on the host cores with all threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VIDEOS, FRAMES_PER_VIDEO, SIZE = 64, 32, 224
METRIC = "frames/sec EfficientNet-B0 inference at 1/2/4/8 B200, % roofline, vs CPU ref"
WORKLOAD = "per-video scoring: 64 videos x 32 uint8 224x224 face crops per GPU -> fused preprocess + EfficientNet-B0 + attention pool + head"


_saved_stdout_fd = None


def stdout_to_stderr():
    """Anything C libraries (NCCL's version banner, ...) write to fd 1 while the bench runs goes to stderr, so that stdout
    carries exactly the one JSON line."""
    global _saved_stdout_fd
    sys.stdout.flush()
    _saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)


def restore_stdout():
    global _saved_stdout_fd
    if _saved_stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(_saved_stdout_fd, 1)
        os.close(_saved_stdout_fd)
        _saved_stdout_fd = None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1387.1))), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80}
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_reference_fps(sd, threads: int, runs: int, frames: int = 32, budget_s: float = 0.0):
    """Reference CPU path (oracle port): one video of `frames` crops, B=1 call, fp32, no_grad.
    With budget_s > 0 the number of timed runs is chosen from the warm-up time so that the sample takes about that long."""
    from oracle import effnet_b0_oracle as O          # the oracle is the checker/baseline here, never the measured product
    from deepfake_video_detection_b200.synthetic import synth_crops
    crops, _ = synth_crops(123, 1, frames)
    torch.set_num_threads(threads)
    x = O.prep_u8_hwc(crops).unsqueeze(0)
    t0 = time.perf_counter()
    O.detector_forward(sd, x)                     # warm-up
    if budget_s > 0:
        runs = max(runs, min(60, int(budget_s / max(time.perf_counter() - t0, 1e-3))))
    ts = []
    for _ in range(runs):
        t0 = time.perf_counter()
        x = O.prep_u8_hwc(crops).unsqueeze(0)     # the prep is part of the path (app.py:2084-2086)
        O.detector_forward(sd, x)
        ts.append(time.perf_counter() - t0)
    return frames / statistics.median(ts), ts


def run_reference(args, rank, world):
    if rank != 0:
        return
    from deepfake_video_detection_b200.synthetic import load_checkpoint, synth_crops
    from oracle import effnet_b0_oracle as O          # reference arm = the oracle port of the reference's CPU path
    sd = load_checkpoint(0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    crops, _ = synth_crops(123, 1, FRAMES_PER_VIDEO)
    step = lambda: O.detector_forward(sd, O.prep_u8_hwc(crops).unsqueeze(0))
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = args.steps * FRAMES_PER_VIDEO / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 video x 32 crops per step (B=1 reference call)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps of 1 video x 32 crops, oracle port of the reference CPU path, fp32"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp16", "bf16"])
    ap.add_argument("--videos", type=int, default=VIDEOS)
    ap.add_argument("--frames", type=int, default=FRAMES_PER_VIDEO)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from deepfake_video_detection_b200 import DEFAULT_PRECISION, FrameScorer, _lib, make_offsets
    from deepfake_video_detection_b200.sharding import gather_video_logits

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    stdout_to_stderr()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision or DEFAULT_PRECISION

    # weights: calibrated synthetic checkpoint with the reference's state_dict schema (the real one is an absent LFS blob)
    from deepfake_video_detection_b200.synthetic import load_checkpoint
    sd = load_checkpoint(0)
    scorer = FrameScorer(sd, precision, dev)
    V, T = args.videos, args.frames
    F = V * T
    g = torch.Generator(device=dev).manual_seed(rank)
    crops = torch.randint(0, 256, (F, SIZE, SIZE, 3), dtype=torch.uint8, device=dev, generator=g)   # 308 MB > L2
    offsets = make_offsets([T] * V, dev)
    total_videos = V * world

    def step():
        logits, _ = scorer.score(crops, offsets)
        if world > 1:
            logits = gather_video_logits(logits, total_videos)
        return logits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    sampler.stop_flag.set()
    ms = e0.elapsed_time(e1)
    launches_per_step = scorer.last_launch_count
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * F * args.steps / (ms / 1e3)

    # ---- e2e: host (pinned) crops -> H2D -> score -> D2H logits, through the public API ---------------
    host = torch.empty((F, SIZE, SIZE, 3), dtype=torch.uint8, pin_memory=True)
    host.copy_(crops)
    host_logits = torch.empty((V, 2), dtype=torch.float32, pin_memory=True)

    lens = [T] * V

    def e2e_step():
        lg, _ = scorer.score_host(host, lens)          # public API: chunked H2D overlapped with scoring
        if world > 1:
            lg = gather_video_logits(lg, total_videos)[rank * V:(rank + 1) * V]
        host_logits.copy_(lg, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    # the same K steps as the resident measurement: the first step's copy cannot overlap anything (the pipeline was drained by
    # the barrier), so a short run would mostly measure that fill
    n_e2e = max(2, args.steps)
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {"value": world * F * n_e2e / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": host.numel(),
           "d2h_bytes_per_step": host_logits.numel() * 4, "steps": n_e2e}

    # ---- one profiled step: per-kernel-class device time vs algorithmic bytes ----------------------------
    lib = _lib.load()
    hbm_gbs, tf_peak, peak_kind = measured_peaks()
    lib.dfd_profile_enable(1)
    scorer.score(crops, offsets)
    entries = (_lib.ProfileEntry * 16)()
    n = C.c_int()
    _lib.check(lib.dfd_profile_collect(entries, 16, C.byref(n)), "profile_collect")
    lib.dfd_profile_enable(0)
    kernels = {}
    for e in entries[: n.value]:
        if e.launches:
            kernels[e.name.decode()] = {"launches": e.launches, "ms": round(e.ms, 4), "GBps": round(e.bytes / e.ms / 1e6, 1),
                                        "TFLOPs": round(e.flops / e.ms / 1e9, 2), "MB": round(e.bytes / 1e6, 1)}
    # DRAM traffic per launch of each kernel class from the committed ncu capture of the same workload (profiles/)
    traffic = {}
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
        if files and (V, T) == (VIDEOS, FRAMES_PER_VIDEO):
            tj = json.load(open(files[-1]))
            traffic = {k: v["dram_bytes_per_launch"] for k, v in tj["classes"].items()}
            traffic["__file__"] = os.path.basename(files[-1])
    except Exception:
        traffic = {}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    d = kernels[dom]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["GBps"], "peak": hbm_gbs, "unit": "GB/s",
                "frac": round(d["GBps"] / hbm_gbs, 4), "traffic": traffic.get(dom), "traffic_source": traffic.get("__file__"), "peak_kind": f"of {peak_kind}",
                "avg_launch_ms": round(d["ms"] / d["launches"], 4), "algorithmic_bytes_per_launch": d["MB"] * 1e6 / d["launches"],
                "share_of_step": round(d["ms"] / sum(k["ms"] for k in kernels.values()), 3)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "videos_per_gpu": V, "frames_per_video": T, "crop": SIZE, "l2": "inputs (308 MB/GPU) and every activation tensor exceed the 126 MB L2",
                       "weights": "calibrated synthetic checkpoint, reference state_dict schema (366 tensors)",
                       "chunk_frames": int(os.environ.get("DFD_CHUNK_FRAMES", "2048")), "parallelism": f"videos sharded over {world} GPU(s), one all-gather of logits",
                       "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("DFD_")}},
            "clocks": sampler.result(), "e2e": e2e, "gpu_launches": launches_per_step * args.steps + (0),
            "gpu_launches_per_step": launches_per_step, "roofline": roofline, "kernels": kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fps_all, ts_all = cpu_reference_fps(sd, cores, 3, budget_s=10.0)
            fps_1, ts_1 = cpu_reference_fps(sd, 1, 2, budget_s=4.0)
            line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"1 video x 32 crops (BASELINE configs[0]), median of {len(ts_all)} runs ({sum(ts_all):.1f} s) after 1 warm-up, "
                                              "oracle port of the reference CPU path, fp32, all host threads",
                                    "value_1_thread": fps_1, "sample_1_thread": f"median of {len(ts_1)} runs ({sum(ts_1):.1f} s), torch.set_num_threads(1) as the reference deploys (app.py:5-8)"}
        restore_stdout()
        print(json.dumps(line), flush=True)
        stdout_to_stderr()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
