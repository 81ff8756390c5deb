"""Multi-GPU scoring: videos are independent, so whole videos are sharded over ranks (never split — the
softmax over T in pretrained_detector.py:127 is per video) and the only exchange is one all-gather of the
per-video logits (8 bytes per video).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests of the host logic)."""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_videos: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: rank r owns videos [lo, hi); sizes differ by at most one."""
    base, rem = divmod(num_videos, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_video_logits(local_logits: torch.Tensor, num_videos: int, group=None) -> torch.Tensor:
    """All-gather per-video logits of a `shard_bounds` partition back into global video order.

    Ranks may own different counts, so every rank pads to the maximum shard size and the padding is
    dropped after the collective (all_gather needs equal shapes)."""
    world = dist.get_world_size(group)
    max_local = (num_videos + world - 1) // world
    n, c = local_logits.shape
    if n == max_local:
        buf = local_logits.contiguous()                       # full shard: no padding copy
    else:
        buf = local_logits.new_zeros((max_local, c))
        buf[:n] = local_logits
    out = local_logits.new_empty((world * max_local, c))      # concatenation layout (what gloo and NCCL both accept)
    dist.all_gather_into_tensor(out, buf, group=group)        # ONE collective, no per-rank list copies
    if num_videos == world * max_local:
        return out                                            # equal shards: rank order is global video order
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(num_videos, world, r)
        parts.append(out[r * max_local: r * max_local + hi - lo])
    return torch.cat(parts, dim=0)


def score_videos_sharded(score_fn: Callable[[int, int], torch.Tensor], num_videos: int, group=None) -> torch.Tensor:
    """`score_fn(lo, hi)` scores this rank's videos [lo, hi) and returns their logits (hi-lo, 2)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(num_videos, world, rank)
    local = score_fn(lo, hi)
    if local.shape[0] != hi - lo:
        raise RuntimeError(f"score_fn returned {local.shape[0]} videos for shard [{lo},{hi})")
    return gather_video_logits(local, num_videos, group)


def frames_of_shard(offsets: Sequence[int], lo: int, hi: int) -> Tuple[int, int]:
    """Frame range [a, b) covered by videos [lo, hi) of a global offsets array."""
    return int(offsets[lo]), int(offsets[hi])


def bind_host_to_gpu(device_index: int):
    """Pin the calling process to the CPU cores of the NUMA node its GPU hangs off, BEFORE it allocates pinned host buffers:
    page-locked memory is placed on the node of the allocating thread, and with one process per GPU feeding 26 GB/s of crops each
    (`FrameScorer.score_host`), buffers on the far socket make eight ranks share one inter-socket link.  Returns a description
    of what was done, or None when the topology cannot be read or the affinity cannot be changed (nothing is changed then)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"cuda:{device_index} -> NUMA node {node}, {len(allowed)} cores"
    except Exception:
        return None
