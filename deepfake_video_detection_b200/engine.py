"""Host-side engine over the C ABI: packed weights, workspace handling and the scoring entry points.

`FrameScorer` is the high-throughput surface (uint8 crops in, per-video logits out); the drop-in
`PretrainedBackboneDetector` in `pretrained_detector.py` sits on top of it.  torch is used for device
memory, streams and the caching allocator only — every arithmetic op runs in libdfd_b200.so.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Mapping, Optional, Sequence, Tuple

import torch

from . import _lib

PRECISIONS = {"fp16": _lib.DTYPE_FP16, "bf16": _lib.DTYPE_BF16}
TORCH_DTYPE = {"fp16": torch.float16, "bf16": torch.bfloat16}
DEFAULT_PRECISION = "fp16"      # bf16 storage misses the 2e-2 logit bar, fp16 meets it


def normalize_key(key: str) -> str:
    """Strip checkpoint wrappers the way the reference's `_normalize_state_dict_keys` does (app.py:1413-1432): 'module.',
    'model.' and 'net.' prefixes are removed repeatedly until none is left ('model.module.backbone…' -> 'backbone…')."""
    changed = True
    while changed:
        changed = False
        for prefix in ("module.", "model.", "net."):
            if key.startswith(prefix):
                key = key[len(prefix):]
                changed = True
    return key


def check_offsets(offsets_host: Sequence[int], frames: int) -> None:
    """offsets must start at 0, never decrease, end at `frames` and describe videos of 1..1024 frames (the pool kernel's range)."""
    off = [int(o) for o in offsets_host]
    if not off or off[0] != 0 or off[-1] != frames:
        raise ValueError(f"offsets must run from 0 to the number of frames ({frames}), got {off[:1]}..{off[-1:]}")
    for a, b in zip(off, off[1:]):
        if b - a < 1 or b - a > 1024:
            raise ValueError("every video needs between 1 and 1024 frames")


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class PackedWeights:
    """BN-folded, repacked weights resident on one GPU (opaque `dfd_weights_t*`)."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], precision: str = DEFAULT_PRECISION,
                 device: torch.device | str | int = "cuda"):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError("deepfake_video_detection_b200 runs on CUDA devices only (no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        names, keep = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked") or not torch.is_tensor(v):
                continue
            names.append(normalize_key(k).encode())
            keep.append(v.detach().to("cpu", torch.float32).contiguous())
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_data = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
        c_numel = (C.c_int64 * n)(*[t.numel() for t in keep])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.dfd_pack_weights(n, c_names, c_data, c_numel, PRECISIONS[precision], C.byref(handle)),
                       "pack_weights")
        self._handle = handle
        self._lib = lib

    @property
    def handle(self) -> C.c_void_p:
        if self._handle is None:
            raise RuntimeError("PackedWeights already freed")
        return self._handle

    def free(self) -> None:
        if getattr(self, "_handle", None) is not None:
            self._lib.dfd_free_weights(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FrameScorer:
    """Batched EfficientNet-B0 frame scoring on one GPU.

    score(crops, offsets)   uint8 (F,H,W,3) crops of V ragged videos -> logits (V,2), frame_scores (F,)
    features(frames)        frames -> pooled trunk features (F,1280) fp32
    Each video is scored exactly as one reference `model(faces.unsqueeze(0))` call (app.py:2086-2089).
    """

    def __init__(self, state_dict: Mapping[str, torch.Tensor], precision: str = DEFAULT_PRECISION,
                 device: torch.device | str | int = "cuda", use_temporal_attention: bool = True):
        self.weights = PackedWeights(state_dict, precision, device)
        self.device = self.weights.device
        self.precision = precision
        self.use_temporal_attention = bool(use_temporal_attention)
        self._lib = _lib.load()
        self.last_launch_count = 0
        self._ws = {}                       # stream handle -> persistent workspace (grown on demand, reused across calls)
        self._ws_lock = threading.Lock()

    def _workspace(self, nbytes: int) -> torch.Tensor:
        """Persistent per-(device, stream) workspace: calls on one stream run in stream order, so they can share it; calls on
        different streams (re-entrant use from several request threads) get their own.  A regrow frees the old
        block through the caching allocator on the stream that used it, i.e. behind the kernels still reading it."""
        if torch.cuda.is_current_stream_capturing():        # a captured call owns its workspace (graph-private memory pool)
            return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        key = _stream_ptr(self.device)
        with self._ws_lock:
            ws = self._ws.get(key)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                if len(self._ws) > 16:
                    self._ws.clear()
                self._ws[key] = ws
        return ws

    # ---- helpers -----------------------------------------------------------------------------
    def _classify_input(self, frames: torch.Tensor) -> Tuple[int, int, int, int]:
        if frames.device != self.device:
            raise RuntimeError(f"frames are on {frames.device}, scorer is on {self.device} (no implicit copies, no CPU path)")
        if not frames.is_contiguous():
            raise ValueError("frames must be contiguous")
        if frames.dim() != 4:
            raise ValueError(f"expected 4-D frames, got shape {tuple(frames.shape)}")
        if frames.dtype == torch.uint8:
            F, H, W, c = frames.shape
            kind = _lib.IN_U8_HWC
        else:
            F, c, H, W = frames.shape
            if frames.dtype == torch.float32:
                kind = _lib.IN_F32_NCHW
            elif frames.dtype == TORCH_DTYPE[self.precision]:
                kind = _lib.IN_H16_NCHW
            else:
                raise ValueError(f"unsupported frame dtype {frames.dtype} for precision {self.precision}")
        if c != 3:
            raise ValueError(f"frames must have 3 channels, got {c}")
        return kind, int(F), int(H), int(W)

    def preprocess(self, crops_u8: torch.Tensor) -> torch.Tensor:
        """K1: uint8 (F,H,W,3) -> normalised 16-bit (F,3,H,W)  (app.py:2084-2085)."""
        kind, F, H, W = self._classify_input(crops_u8)
        if kind != _lib.IN_U8_HWC:
            raise ValueError("preprocess expects uint8 (F,H,W,3) crops")
        out = torch.empty((F, 3, H, W), dtype=TORCH_DTYPE[self.precision], device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dfd_preprocess_u8hwc_to_nchw(crops_u8.data_ptr(), out.data_ptr(), F, H, W,
                                                             PRECISIONS[self.precision], _stream_ptr(self.device)), "preprocess")
        return out

    def features(self, frames: torch.Tensor) -> torch.Tensor:
        kind, F, H, W = self._classify_input(frames)
        feat = torch.empty((F, _lib.FEATURE_DIM), dtype=torch.float32, device=self.device)
        if F == 0:
            return feat
        nbytes = C.c_size_t()
        _lib.check(self._lib.dfd_workspace_bytes(F, H, W, C.byref(nbytes)), "workspace_bytes")
        ws = self._workspace(nbytes.value)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dfd_effnet_b0_features(self.weights.handle, frames.data_ptr(), kind, F, H, W,
                                                        feat.data_ptr(), ws.data_ptr(), nbytes.value,
                                                        _stream_ptr(self.device)), "effnet_b0_features")
        self.last_launch_count = self._lib.dfd_last_launch_count()
        return feat

    def pool_head(self, features: torch.Tensor, offsets: torch.Tensor,
                  use_temporal_attention: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if features.device != self.device or offsets.device != self.device:
            raise RuntimeError("features/offsets must live on the scorer's device")
        if offsets.dtype != torch.int32 or offsets.dim() != 1 or offsets.numel() < 1:
            raise ValueError("offsets must be a 1-D int32 tensor of V+1 frame offsets")
        features = features.contiguous()
        V, F = offsets.numel() - 1, features.shape[0]
        att = self.use_temporal_attention if use_temporal_attention is None else bool(use_temporal_attention)
        logits = torch.empty((V, 2), dtype=torch.float32, device=self.device)
        scores = torch.empty((F,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dfd_attn_pool_head(self.weights.handle, features.data_ptr(), offsets.data_ptr(), V, F,
                                                    int(att), logits.data_ptr(), scores.data_ptr(),
                                                    _stream_ptr(self.device)), "attn_pool_head")
        return logits, scores

    def score(self, frames: torch.Tensor, offsets: torch.Tensor, use_temporal_attention: Optional[bool] = None,
              return_features: bool = False):
        """Whole path for V ragged videos.  `offsets` int32 (V+1,) on the device, validated by the caller
        (see `make_offsets`)."""
        kind, F, H, W = self._classify_input(frames)
        if offsets.device != self.device or offsets.dtype != torch.int32 or offsets.dim() != 1:
            raise ValueError("offsets must be a 1-D int32 tensor on the scorer's device")
        V = offsets.numel() - 1
        att = self.use_temporal_attention if use_temporal_attention is None else bool(use_temporal_attention)
        logits = torch.empty((V, 2), dtype=torch.float32, device=self.device)
        scores = torch.empty((F,), dtype=torch.float32, device=self.device)
        feat = torch.empty((F, _lib.FEATURE_DIM), dtype=torch.float32, device=self.device) if return_features else None
        if V == 0 or F == 0:
            return (logits, scores, feat) if return_features else (logits, scores)
        nbytes = C.c_size_t()
        _lib.check(self._lib.dfd_score_workspace_bytes(F, H, W, C.byref(nbytes)), "score_workspace_bytes")
        ws = self._workspace(nbytes.value)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dfd_score_videos(self.weights.handle, frames.data_ptr(), kind, offsets.data_ptr(), V, F,
                                                  H, W, int(att), logits.data_ptr(), scores.data_ptr(), _ptr(feat),
                                                  ws.data_ptr(), nbytes.value, _stream_ptr(self.device)), "score_videos")
        self.last_launch_count = self._lib.dfd_last_launch_count()
        return (logits, scores, feat) if return_features else (logits, scores)


    def score_host(self, host_crops: torch.Tensor, frames_per_video: Sequence[int], chunk_videos: Optional[int] = None,
                   use_temporal_attention: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """End-to-end scoring of HOST crops: uint8 (F,H,W,3) in (ideally pinned) host memory -> logits (V,2) and
        frame_scores (F,) on the device, asynchronously (nothing here waits for the GPU).

        The H2D copy runs on a side stream into one of two persistent device staging buffers, so the transfer of
        one call (or chunk) overlaps the scoring of the previous one; back-to-back calls therefore run at
        max(copy, compute) per batch instead of copy + compute.  `host_crops` must stay unmodified until the
        returned tensors are consumed (usual async-copy contract).  `chunk_videos` splits a call into chunks
        (smaller staging buffers; per-video results do not depend on the chunking)."""
        if host_crops.device.type != "cpu" or host_crops.dtype != torch.uint8 or host_crops.dim() != 4:
            raise ValueError("score_host expects uint8 (F,H,W,3) crops in host memory")
        lens = [int(t) for t in frames_per_video]
        if sum(lens) != host_crops.shape[0]:
            raise ValueError("frames_per_video does not add up to the number of crops")
        V = len(lens)
        logits = torch.empty((V, 2), dtype=torch.float32, device=self.device)
        scores = torch.empty((host_crops.shape[0],), dtype=torch.float32, device=self.device)
        if V == 0:
            return logits, scores
        main = torch.cuda.current_stream(self.device)
        chunk_videos = V if not chunk_videos else int(chunk_videos)
        chunks, v0, f0 = [], 0, 0
        while v0 < V:
            v1 = min(V, v0 + chunk_videos)
            nf = sum(lens[v0:v1])
            chunks.append((v0, v1, f0, f0 + nf))
            v0, f0 = v1, f0 + nf
        need = max(b - a for _, _, a, b in chunks) * host_crops[0].numel()
        st = getattr(self, "_staging", None)
        if st is None or st["bytes"] < need:
            # The staging buffers come from the caching allocator on the MAIN stream, so they may alias a block that earlier
            # main-stream work (a previous score(), the old staging buffers of a smaller call) is still using: the copy
            # stream's first write into each slot must wait for everything queued on main so far.
            seed = torch.cuda.Event()
            cs = st["copy_stream"] if st is not None else torch.cuda.Stream(self.device)
            st = {"bytes": need, "copy_stream": cs, "turn": 0,
                  "buf": [torch.empty(need, dtype=torch.uint8, device=self.device) for _ in range(2)],
                  "freed": [seed, seed]}
            seed.record(main)
            self._staging = st
        cs = st["copy_stream"]
        frame_shape = tuple(host_crops.shape[1:])
        for (va, vb, fa, fb) in chunks:
            slot = st["turn"] & 1
            st["turn"] += 1
            dst = st["buf"][slot][: (fb - fa) * host_crops[0].numel()].view((fb - fa,) + frame_shape)
            with torch.cuda.stream(cs):
                cs.wait_event(st["freed"][slot])                   # the batch that last used this buffer has been scored
                dst.copy_(host_crops[fa:fb], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(cs)
            main.wait_event(ready)
            key = tuple(lens[va:vb])
            cache = self.__dict__.setdefault("_offsets_cache", {})
            if key not in cache:                                   # small H2D upload, once per distinct length pattern
                if len(cache) > 64:
                    cache.clear()
                cache[key] = make_offsets(key, self.device)
            lg, sc = self.score(dst, cache[key], use_temporal_attention)
            logits[va:vb].copy_(lg)
            scores[fa:fb].copy_(sc)
            ev = torch.cuda.Event()
            ev.record(main)
            st["freed"][slot] = ev
        return logits, scores


    def capture(self, frames_per_video: Sequence[int], height: int = 224, width: int = 224) -> "GraphedScorer":
        """Capture the whole scoring step for a fixed batch shape into a CUDA graph (one launch instead of ~70):
        the latency path for small batches (BASELINE config 4, F = 1 ... 64)."""
        return GraphedScorer(self, frames_per_video, height, width)


class GraphedScorer:
    """A `FrameScorer.score` call frozen into a CUDA graph.  Write uint8 crops into `.input` (or pass them to
    `run`), replay, read `.logits` (V,2) / `.frame_scores` (F,) — static tensors owned by the graph."""

    def __init__(self, scorer: FrameScorer, frames_per_video: Sequence[int], height: int, width: int):
        self.scorer = scorer
        dev = scorer.device
        lens = [int(t) for t in frames_per_video]
        self.offsets = make_offsets(lens, dev)
        self.input = torch.zeros((sum(lens), height, width, 3), dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside capture (lazy module load, attributes)
            scorer.score(self.input, self.offsets)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits, self.frame_scores = scorer.score(self.input, self.offsets)

    def run(self, crops_u8: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if crops_u8 is not None:
            self.input.copy_(crops_u8, non_blocking=True)
        self.graph.replay()
        return self.logits, self.frame_scores


def make_offsets(frames_per_video: Sequence[int], device) -> torch.Tensor:
    """Host-side validation + upload of ragged video lengths (each video needs 1..1024 frames)."""
    lens = [int(t) for t in frames_per_video]
    if any(t < 1 or t > 1024 for t in lens):
        raise ValueError("every video needs between 1 and 1024 frames")
    off = torch.zeros(len(lens) + 1, dtype=torch.int32)
    if lens:
        off[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
    return off.to(device)
