#!/usr/bin/env python
"""Benchmark of the hot path: per-video EfficientNet-B0 real/fake scoring of uint8 face crops.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|5|ensemble]

--config 2 (default; BASELINE.json configs[1], the headline): 64 videos x 32 uint8 224x224 crops per GPU -> fused prep +
  trunk + temporal-attention pool + head -> per-video logits.  One "step" = one pass over that batch.  With N > 1 the driver
  launches one rank per GPU (torchrun): every rank scores its own 64 videos (weak scaling) and the step ends with the single
  all-gather of per-video logits.  The N > 1 line also carries a `strong` record (BASELINE configs[3]): one FIXED 64 x 32
  batch, and one fixed ragged 128-frame batch, sharded over the N ranks with `score_videos_sharded`.
--config 3: EfficientNet-B0 features + LogicRNNLSTM temporal head, 256 videos x 16 frames (evaluate.py:143-192 wiring).
--config 5: ViT-B/16 frame encoder (models.py:88-107) forward at batch 512, the dense-contraction stress case (tensor roofline).
--config ensemble: EnsembleDetector([efficientnet_b0, resnet50], 'weighted') (pretrained_detector.py:179-218) on 8 videos x 32 frames.
Prints ONE JSON line on rank 0.

  value     metric units per second, inputs resident in HBM, CUDA events over exactly K steps, max over ranks
  steady    the same step repeated for >= 1 s with one event pair per step: median / p10 / p90 (a 0.3 s timed region moves by
            percents with one power-cap excursion; the median does not)
  e2e       same metric through the public API with HOST (pinned) inputs: H2D copy + step + D2H of the result per step
  roofline  dominant kernel: algorithmic bytes (or flops) / measured time vs the measured peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle (port of the reference's CPU path) on the host cores, bounded sample (config 2, N = 1)
--impl reference: the reference's CPU path (oracle port; the reference itself is pure Python over an un-vendored timm and
cannot travel to the GPU box) timed on the host cores with all threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
STRONG_SMALL_LENS = [32, 16, 16, 8, 8, 8, 8, 8, 8, 4, 4, 4, 4]          # 13 ragged videos, 128 frames (uneven shards at every N)

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


# ------------------------------------------------------------------------------------------------ CPU reference arm
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


# ------------------------------------------------------------------------------------------------ timing helpers
class Timer:
    """CUDA-event timing on the current stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms: float) -> float:
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(self, step, n: int) -> float:
        """ms for EXACTLY n back-to-back steps."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def steady(self, step, est_ms: float, min_s: float = 1.0, max_steps: int = 400):
        """>= min_s of steps, one event pair per step: median / p10 / p90 over this rank's steps (rank 0 reports)."""
        n = int(min(max_steps, max(10, min_s * 1e3 / max(est_ms, 1e-3))))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        self.barrier()
        ev[0].record()
        for i in range(n):
            step()
            ev[i + 1].record()
        self.barrier()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
        return {"steps": n, "median_ms": round(ts[n // 2], 4), "p10_ms": round(ts[n // 10], 4), "p90_ms": round(ts[(9 * n) // 10], 4),
                "seconds": round(sum(ts) / 1e3, 3)}


def source_digest() -> str:
    """Digest of the kernel sources: a committed ncu traffic record is only quoted for the code it was captured on."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "deepfake_video_detection_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


def committed_traffic(workload_is_default: bool):
    """DRAM bytes per launch of each kernel class from the newest ncu capture committed under profiles/ — quoted only when
    that capture was taken on exactly these kernel sources (else null: a stale number is worse than none)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files or not workload_is_default:
        return {}, None
    try:
        tj = json.load(open(files[-1]))
        if tj.get("source_digest") != source_digest():
            return {}, os.path.basename(files[-1]) + " (stale: kernels changed since the capture)"
        return {k: v["dram_bytes_per_launch"] for k, v in tj["classes"].items()}, os.path.basename(files[-1])
    except Exception:
        return {}, None


def profile_classes(lib, _lib, run_once):
    """One extra profiled pass: per-kernel-class device time (CUDA events around every launch, on the launch stream) vs
    algorithmic bytes / flops."""
    lib.dfd_profile_enable(1)
    run_once()
    entries = (_lib.ProfileEntry * 16)()
    n = C.c_int()
    _lib.check(lib.dfd_profile_collect(entries, 16, C.byref(n)), "profile_collect")
    lib.dfd_profile_enable(0)
    kernels = {}
    for e in entries[: n.value]:
        if e.launches:
            kernels[e.name.decode()] = {"launches": e.launches, "ms": round(e.ms, 4), "GBps": round(e.bytes / e.ms / 1e6, 1),
                                        "TFLOPs": round(e.flops / e.ms / 1e9, 2), "MB": round(e.bytes / 1e6, 1)}
    return kernels


# fused expand + depthwise kernels (blocks 2.1.0 / 2.1.1 / 2.2.0 at 224x224): bytes per frame the two kernels they replace
# would move (x + 2 * expanded + dw_out, 16-bit) and the SiLU evaluations they contain (expanded + dw_out elements)
FUSED_UNFUSED_MB_PER_FRAME = (112 * 112 * (16 + 2 * 96) + 56 * 56 * 96 + 56 * 56 * (24 + 2 * 144) + 56 * 56 * 144
                              + 56 * 56 * (24 + 2 * 144) + 28 * 28 * 144) * 2 / 1e6
FUSED_SILU_PER_FRAME = 112 * 112 * 96 + 56 * 56 * 96 + 2 * 56 * 56 * 144 + 56 * 56 * 144 + 28 * 28 * 144
MUFU_PER_S = 15.9 * 148 * 1.965e9              # measured tanh.approx rate per SM and clock (tools/mufu_probe.cu) x SMs x max clock


def hbm_roofline(kernels, traffic, traffic_src, hbm_gbs, peak_kind, frames=None):
    """`roofline` of the class with the largest share of the step, every class's fraction beside it (`by_class`).  All classes
    are measured against the HBM roof on their algorithmic bytes; the fused expand + depthwise class is NOT HBM-bound (the
    fusion removed 78 % of its traffic): its record says what limits it and what the kernels it replaces would cost."""
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    d = kernels[dom]
    total = sum(k["ms"] for k in kernels.values())
    out = {"kernel": dom, "bound": "hbm", "achieved": d["GBps"], "peak": hbm_gbs, "unit": "GB/s", "frac": round(d["GBps"] / hbm_gbs, 4),
           "traffic": traffic.get(dom), "traffic_source": traffic_src, "peak_kind": f"of {peak_kind}",
           "avg_launch_ms": round(d["ms"] / d["launches"], 4), "algorithmic_bytes_per_launch": d["MB"] * 1e6 / d["launches"],
           "share_of_step": round(d["ms"] / total, 3),
           "by_class": {k: {"share": round(v["ms"] / total, 3), "frac": round(v["GBps"] / hbm_gbs, 4)} for k, v in kernels.items()},
           "whole_step": {"algorithmic_GB": round(sum(k["MB"] for k in kernels.values()) / 1e3, 3), "ms": round(total, 4),
                          "GBps": round(sum(k["MB"] for k in kernels.values()) / total, 1),
                          "frac": round(sum(k["MB"] for k in kernels.values()) / total / hbm_gbs, 4)}}
    f = kernels.get("expand_dwconv_fused")
    if f and frames:
        unfused_mb = FUSED_UNFUSED_MB_PER_FRAME * frames
        own_mb = sum(k["MB"] for k in kernels.values())
        # the same step on the bytes of the REFERENCE's op graph (every op reads and writes its tensors once): the fused
        # kernels counted as the expand GEMM + depthwise pair they replace.  `frac` above stays on the bytes the kernels really need.
        out["whole_step"]["reference_graph_GB"] = round((own_mb - f["MB"] + unfused_mb) / 1e3, 3)
        out["whole_step"]["reference_graph_frac"] = round((own_mb - f["MB"] + unfused_mb) / total / hbm_gbs, 4)
        out["fused_class"] = {"limiter": "latency (12 warps per SM behind a per-row barrier; fewer instructions or fewer barriers measured no gain) with MUFU at 0.4 of its rate; not HBM",
                              "silu_per_s": round(FUSED_SILU_PER_FRAME * frames / (f["ms"] / 1e3), 0), "mufu_peak_per_s": round(MUFU_PER_S, 0),
                              "frac_of_mufu": round(FUSED_SILU_PER_FRAME * frames / (f["ms"] / 1e3) / MUFU_PER_S, 4),
                              "replaces_unfused_MB": round(unfused_mb, 1), "unfused_hbm_floor_ms": round(unfused_mb / hbm_gbs, 4),
                              "unfused_equivalent_GBps": round(unfused_mb / f["ms"], 1)}
    return out


# ------------------------------------------------------------------------------------------------ config 2 (headline)
def bench_config2(args, rank, world, dev, timer, sampler_cls, local_rank):
    import torch.distributed as dist
    from deepfake_video_detection_b200 import DEFAULT_PRECISION, FrameScorer, _lib, make_offsets
    from deepfake_video_detection_b200.sharding import gather_video_logits, score_videos_sharded, shard_bounds
    from deepfake_video_detection_b200.synthetic import load_checkpoint
    precision = args.precision or DEFAULT_PRECISION
    # weights: calibrated synthetic checkpoint with the reference's state_dict schema (the real one is an absent LFS blob)
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

    for _ in range(args.warmup):
        step()
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = timer.timed(step, args.steps)
    sampler.stop_flag.set()
    launches_per_step = scorer.last_launch_count
    value = world * F * args.steps / (ms / 1e3)
    # one profiled step right behind the timed ones (the state `value` was measured in, before the >= 1 s steady run heats the
    # part up): per-kernel-class device time vs algorithmic bytes
    lib = _lib.load()
    kernels = profile_classes(lib, _lib, lambda: scorer.score(crops, offsets))
    steady = timer.steady(step, ms / args.steps)

    # ---- e2e: host (pinned) crops -> H2D -> score -> D2H logits, through the public API ---------------
    host = torch.empty((F, SIZE, SIZE, 3), dtype=torch.uint8, pin_memory=True)
    host.copy_(crops)
    host_logits = torch.empty((V, 2), dtype=torch.float32, pin_memory=True)
    lens = [T] * V

    def e2e_step():
        lg, _ = scorer.score_host(host, lens)          # public API: H2D on a side stream overlapped with the previous call's scoring
        if world > 1:
            lg = gather_video_logits(lg, total_videos)[rank * V:(rank + 1) * V]
        host_logits.copy_(lg, non_blocking=True)

    for _ in range(2):
        e2e_step()
    # the same K steps as the resident measurement: the first step's copy cannot overlap anything (the pipeline was drained by
    # the barrier), so a short run would mostly measure that fill
    n_e2e = max(2, args.steps)
    ms_e2e = timer.timed(e2e_step, n_e2e)
    e2e = {"value": world * F * n_e2e / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": host.numel(),
           "d2h_bytes_per_step": host_logits.numel() * 4, "steps": n_e2e}

    # ---- strong scaling (BASELINE configs[3]): one FIXED batch sharded over the ranks ------------------
    strong = None
    if world > 1:
        gg = torch.Generator(device=dev).manual_seed(1234)                      # every rank builds the same global batch, scores its shard
        gl_crops = torch.randint(0, 256, (VIDEOS * FRAMES_PER_VIDEO, SIZE, SIZE, 3), dtype=torch.uint8, device=dev, generator=gg)
        strong = {}
        for name, lens_g in (("64x32", [FRAMES_PER_VIDEO] * VIDEOS), ("ragged_128_frames", STRONG_SMALL_LENS)):
            off_g = [0]
            for t in lens_g:
                off_g.append(off_g[-1] + t)
            lo, hi = shard_bounds(len(lens_g), world, rank)
            my_off = make_offsets(lens_g[lo:hi], dev)
            my_crops = gl_crops[off_g[lo]:off_g[hi]]

            def score_fn(a, b, my_crops=my_crops, my_off=my_off):
                return scorer.score(my_crops, my_off)[0]

            def sstep(lens_g=lens_g, score_fn=score_fn):
                return score_videos_sharded(score_fn, len(lens_g))

            for _ in range(3):
                out = sstep()
            n = max(args.steps, 10)
            ms_s = timer.timed(sstep, n)
            ref_rank0 = scorer.score(gl_crops[: off_g[-1]], make_offsets(lens_g, dev))[0]       # unsharded, same GPU: must be the same bits
            strong[name] = {"videos": len(lens_g), "frames": off_g[-1], "ms_per_step": round(ms_s / n, 4),
                            "frames_per_s": round(off_g[-1] * n / (ms_s / 1e3), 1), "shard_videos": [shard_bounds(len(lens_g), world, r)[1] - shard_bounds(len(lens_g), world, r)[0] for r in range(world)],
                            "bit_identical_to_unsharded": bool(torch.equal(out, ref_rank0))}
        lg_local = torch.zeros((V, 2), device=dev)
        n = 200
        ms_g = timer.timed(lambda: gather_video_logits(lg_local, total_videos), n)
        strong["allgather_us"] = round(1e3 * ms_g / n, 2)
        strong["scaling"] = "strong: total work fixed as N grows (BASELINE configs[3]); the headline `value` is weak scaling"

    # ---- roofline of the profiled step ------------------------------------------------------------------
    hbm_gbs, tf_peak, peak_kind = measured_peaks()
    traffic, traffic_src = committed_traffic((V, T) == (VIDEOS, FRAMES_PER_VIDEO))
    roofline = hbm_roofline(kernels, traffic, traffic_src, hbm_gbs, peak_kind, frames=V * T)

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": precision, "data": "synthetic",
        "config": {"workload": WORKLOAD, "videos_per_gpu": V, "frames_per_video": T, "crop": SIZE, "l2": "inputs (308 MB/GPU) and every activation tensor exceed the 126 MB L2",
                   "weights": "calibrated synthetic checkpoint, reference state_dict schema (366 tensors)",
                   "chunk_frames": int(os.environ.get("DFD_CHUNK_FRAMES", "2048")), "parallelism": f"videos sharded over {world} GPU(s), one all-gather of logits",
                   "host_affinity": getattr(args, "host_affinity", None),
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("DFD_")}},
        "clocks": sampler.result(), "steady": steady, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step, "roofline": roofline, "kernels": kernels,
    }
    if strong:
        line["strong"] = strong
    if world == 1 and not args.no_cpu_baseline and rank == 0:
        cores = os.cpu_count() or 1
        fps_all, ts_all = cpu_reference_fps(sd, cores, 3, budget_s=10.0)
        fps_1, ts_1 = cpu_reference_fps(sd, 1, 2, budget_s=4.0)
        line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"1 video x 32 crops (BASELINE configs[0]), median of {len(ts_all)} runs ({sum(ts_all):.1f} s) after 1 warm-up, "
                                          "oracle port of the reference CPU path, fp32, all host threads",
                                "value_1_thread": fps_1, "sample_1_thread": f"median of {len(ts_1)} runs ({sum(ts_1):.1f} s), torch.set_num_threads(1) as the reference deploys (app.py:5-8)"}
    return line


# ------------------------------------------------------------------------------------------------ configs 3 / 5 / ensemble
def generic_line(args, world, metric, unit, value, ms, steady, e2e, launches, roofline, config, sampler, precision, extra=None):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision,
            "data": "synthetic", "config": config, "clocks": sampler.result(), "steady": steady, "e2e": e2e,
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches, "roofline": roofline}
    if extra:
        line.update(extra)
    return line


def bench_config3(args, rank, world, dev, timer, sampler_cls, local_rank):
    """EfficientNet-B0 features + LogicRNNLSTM(1280, 512, 2) over 16-frame sequences, 256 videos per GPU (evaluate.py:143-192)."""
    from deepfake_video_detection_b200 import DEFAULT_PRECISION, FrameScorer, _lib
    from deepfake_video_detection_b200.rnn_model import LogicRNNLSTM
    from deepfake_video_detection_b200.synthetic import load_checkpoint
    precision = args.precision or DEFAULT_PRECISION
    V, T = 256, 16
    F = V * T
    torch.manual_seed(0)
    scorer = FrameScorer(load_checkpoint(0), precision, dev)
    rnn = LogicRNNLSTM(1280, 512, 2, precision=precision).eval().to(dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    crops = torch.randint(0, 256, (F, SIZE, SIZE, 3), dtype=torch.uint8, device=dev, generator=g)
    lib = _lib.load()
    counts = {}

    def step(x=crops):
        with torch.no_grad():
            feats = scorer.features(x)
            counts["trunk"] = scorer.last_launch_count
            prob = rnn(feats.view(V, T, 1280))
            counts["rnn"] = lib.dfd_last_launch_count()
            return prob

    for _ in range(args.warmup):
        step()
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = timer.timed(step, args.steps)
    sampler.stop_flag.set()
    steady = timer.steady(step, ms / args.steps)
    host = torch.empty((F, SIZE, SIZE, 3), dtype=torch.uint8, pin_memory=True)
    host.copy_(crops)
    stage = torch.empty_like(crops)
    host_out = torch.empty((V, 1), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        stage.copy_(host, non_blocking=True)
        host_out.copy_(step(stage), non_blocking=True)

    e2e_step()
    n = max(2, args.steps)
    ms_e = timer.timed(e2e_step, n)
    # head alone
    feats = scorer.features(crops).view(V, T, 1280)
    with torch.no_grad():
        ms_rnn = timer.timed(lambda: rnn(feats), 10) / 10
    hbm_gbs, tf_peak, peak_kind = measured_peaks()
    kernels = profile_classes(lib, _lib, lambda: scorer.features(crops))
    roofline = hbm_roofline(kernels, {}, None, hbm_gbs, peak_kind)
    roofline["rnn_head"] = {"ms": round(ms_rnn, 4), "flops_per_video": 0.302e9, "TFLOPs": round(0.302e9 * V / ms_rnn / 1e9, 2),
                            "note": "weight-bandwidth / launch-latency bound at this batch (SURVEY.md 8d): 64 gate GEMMs + cell kernels per sequence batch"}
    cfg = {"workload": "EfficientNet-B0 features + RNNModel temporal head over 16-frame sequences, batch 256 videos (BASELINE configs[2])",
           "videos_per_gpu": V, "frames_per_video": T, "rnn": "LogicRNNLSTM(1280, 512, 2), lengths=None, seeded random init",
           "l2": "inputs (617 MB/GPU) exceed the 126 MB L2", "parallelism": f"{world} independent replica(s): no collective on this path"}
    return generic_line(args, world, "frames/sec EfficientNet-B0 features + LogicRNNLSTM head (BASELINE configs[2])", "frames/s",
                        world * F * args.steps / (ms / 1e3), ms, steady,
                        {"value": world * F * n / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": host.numel(), "d2h_bytes_per_step": V * 4, "steps": n},
                        counts["trunk"] + counts["rnn"], roofline, cfg, sampler, precision, {"kernels": kernels, "videos_per_s": world * V * args.steps / (ms / 1e3)})


def bench_config5(args, rank, world, dev, timer, sampler_cls, local_rank):
    """ViT-B/16 (timm vit_base_patch16_224, num_classes=0; models.py:88-107) forward at batch 512: tensor roofline."""
    from deepfake_video_detection_b200 import DEFAULT_PRECISION, _lib
    from deepfake_video_detection_b200.vit_model import ViTFeatureExtractor
    precision = args.precision or DEFAULT_PRECISION
    B = 512
    torch.manual_seed(0)
    m = ViTFeatureExtractor(precision=precision).eval().to(dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn((B, 3, SIZE, SIZE), device=dev, generator=g)
    lib = _lib.load()

    def step(inp=x):
        with torch.no_grad():
            return m(inp)

    for _ in range(args.warmup):
        step()
    launches = lib.dfd_last_launch_count()
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = timer.timed(step, args.steps)
    sampler.stop_flag.set()
    steady = timer.steady(step, ms / args.steps)
    host = torch.empty((B, 3, SIZE, SIZE), dtype=torch.float32, pin_memory=True)
    host.copy_(x)
    stage = torch.empty_like(x)
    host_out = torch.empty((B, 768), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        stage.copy_(host, non_blocking=True)
        host_out.copy_(step(stage), non_blocking=True)

    e2e_step()
    n = max(2, args.steps)
    ms_e = timer.timed(e2e_step, n)
    hbm_gbs, tf_peak, peak_kind = measured_peaks()
    flops = 35.1e9 * B                                   # 17.56 GMAC per image
    tfl = flops / (steady["median_ms"] / 1e3) / 1e12
    roofline = {"kernel": "whole ViT-B/16 forward (12 x [qkv, attention, proj, fc1, fc2] CTA-pair tcgen05 GEMMs + tcgen05 attention)", "bound": "tensor",
                "achieved": round(tfl, 1), "peak": tf_peak, "unit": "TFLOP/s", "frac": round(tfl / tf_peak, 4), "traffic": None,
                "peak_kind": f"sustained bf16, {peak_kind}", "flops_per_step": flops, "timed_on": "median step of the steady run"}
    cfg = {"workload": "ViT frame encoder (ViT-B/16, 224x224) forward at batch 512 (BASELINE configs[4])", "images_per_gpu": B,
           "weights": "seeded random init, timm vit_base_patch16_224 schema", "l2": "inputs (308 MB/GPU) and activations exceed the 126 MB L2",
           "parallelism": f"{world} independent replica(s): no collective on this path"}
    return generic_line(args, world, "images/sec ViT-B/16 frame encoder forward (BASELINE configs[4])", "images/s",
                        world * B * args.steps / (ms / 1e3), ms, steady,
                        {"value": world * B * n / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": B * 768 * 4, "steps": n},
                        launches, roofline, cfg, sampler, precision)


def bench_ensemble(args, rank, world, dev, timer, sampler_cls, local_rank):
    """EnsembleDetector(['efficientnet_b0', 'resnet50'], 'weighted') — the reference's default pair (app.py:1597)."""
    from deepfake_video_detection_b200 import DEFAULT_PRECISION, EnsembleDetector, PretrainedBackboneDetector, _lib
    precision = args.precision or DEFAULT_PRECISION
    V, T = 8, 32
    F = V * T
    torch.manual_seed(0)
    ens = EnsembleDetector(["efficientnet_b0", "resnet50"], pretrained=False, ensemble_method="weighted", precision=precision).eval().to(dev)
    rn = ens.models[1]
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn((V, T, 3, SIZE, SIZE), device=dev, generator=g)
    lib = _lib.load()
    counts = {}

    def step(inp=x):
        with torch.no_grad():
            return ens(inp)[0]

    def rn_step():
        with torch.no_grad():
            return rn(x)[0]

    for _ in range(args.warmup):
        step()
    with torch.no_grad():
        ens.models[0](x); counts["effnet"] = ens.models[0]._scorer.last_launch_count
        rn(x); counts["resnet"] = lib.dfd_last_launch_count()
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = timer.timed(step, args.steps)
    sampler.stop_flag.set()
    steady = timer.steady(step, ms / args.steps)
    ms_rn = timer.timed(rn_step, 10) / 10
    host = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(x)
    stage = torch.empty_like(x)
    host_out = torch.empty((V, 2), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        stage.copy_(host, non_blocking=True)
        host_out.copy_(step(stage), non_blocking=True)

    e2e_step()
    n = max(2, args.steps)
    ms_e = timer.timed(e2e_step, n)
    hbm_gbs, tf_peak, peak_kind = measured_peaks()
    tfl = 8.2e9 * F / (ms_rn / 1e3) / 1e12               # 4.1 GMAC per 224x224 frame
    roofline = {"kernel": "resnet50 member (53 convolutions as tcgen05 GEMMs, stride-1 3x3 implicit)", "bound": "tensor", "achieved": round(tfl, 1),
                "peak": tf_peak, "unit": "TFLOP/s", "frac": round(tfl / tf_peak, 4), "traffic": None, "peak_kind": f"sustained bf16, {peak_kind}",
                "member_ms": round(ms_rn, 4), "share_of_step": round(ms_rn / (ms / args.steps), 3),
                "note": "whole-member figure; resnet50 at 224x224 is HBM-bound overall (about 75-140 FLOP/B against a ridge of 215), only its 3x3 convolutions are tensor-bound"}
    cfg = {"workload": "EnsembleDetector([efficientnet_b0, resnet50], weighted) on 8 videos x 32 frames, fp32 NCHW input (reference forward contract)",
           "videos_per_gpu": V, "frames_per_video": T, "weights": "seeded random init", "parallelism": f"{world} independent replica(s)"}
    return generic_line(args, world, "frames/sec EnsembleDetector efficientnet_b0 + resnet50 (SURVEY 8f-1)", "frames/s",
                        world * F * args.steps / (ms / 1e3), ms, steady,
                        {"value": world * F * n / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": V * 2 * 4, "steps": n},
                        counts["effnet"] + counts["resnet"], roofline, cfg, sampler, precision)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=["2", "3", "5", "ensemble"])
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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    stdout_to_stderr()
    torch.cuda.set_device(local_rank)
    from deepfake_video_detection_b200.sharding import bind_host_to_gpu
    args.host_affinity = bind_host_to_gpu(local_rank) if world > 1 else None     # one process per GPU: pinned buffers on the GPU's own NUMA node
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    timer = Timer(dev, world)
    fn = {"2": bench_config2, "3": bench_config3, "5": bench_config5, "ensemble": bench_ensemble}[args.config]
    line = fn(args, rank, world, dev, timer, ClockSampler, local_rank)
    if rank == 0:
        restore_stdout()
        print(json.dumps(line), flush=True)
        stdout_to_stderr()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
