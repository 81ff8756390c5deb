"""ctypes binding of `libdfd_b200.so` (include/dfd_b200.h, include/dfd_b200_kernels.h).

There is no fallback: if the library has not been built, importing a symbol raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFD_LIB_PATH") or os.path.join(PKG, "libdfd_b200.so")   # override: kernel-variant experiments only

DTYPE_BF16, DTYPE_FP16 = 0, 1
IN_U8_HWC, IN_F32_NCHW, IN_H16_NCHW = 0, 1, 2
FEATURE_DIM = 1280

_vp, _i64, _int, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
_SIGNATURES = {
    # dfd_b200.h
    "dfd_abi_version": (_int, []),
    "dfd_last_error": (C.c_char_p, []),
    "dfd_last_launch_count": (_int, []),
    "dfd_pack_weights": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _int, C.POINTER(_vp)]),
    "dfd_free_weights": (None, [_vp]),
    "dfd_weights_dtype": (_int, [_vp]),
    "dfd_preprocess_u8hwc_to_nchw": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp]),
    "dfd_workspace_bytes": (_int, [_i64, _int, _int, C.POINTER(C.c_size_t)]),
    "dfd_effnet_b0_features": (_int, [_vp, _vp, _int, _i64, _int, _int, _vp, _vp, C.c_size_t, _vp]),
    "dfd_attn_pool_head": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "dfd_score_workspace_bytes": (_int, [_i64, _int, _int, C.POINTER(C.c_size_t)]),
    "dfd_score_videos": (_int, [_vp, _vp, _int, _vp, _i64, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dfd_rnn_last_error": (C.c_char_p, []),
    "dfd_rnn_pack_weights": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _int, _int, _int, _int, C.POINTER(_vp)]),
    "dfd_rnn_free_weights": (None, [_vp]),
    "dfd_rnn_workspace_bytes": (_int, [_vp, _i64, _int, C.POINTER(C.c_size_t)]),
    "dfd_rnn_forward": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp, C.c_size_t, _vp]),
    "dfd_vit_last_error": (C.c_char_p, []),
    "dfd_vit_pack_weights": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _int, C.POINTER(_vp)]),
    "dfd_vit_free_weights": (None, [_vp]),
    "dfd_vit_workspace_bytes": (_int, [_i64, C.POINTER(C.c_size_t)]),
    "dfd_vit_features": (_int, [_vp, _vp, _i64, _vp, _vp, C.c_size_t, _vp]),
    "dfd_resnet_last_error": (C.c_char_p, []),
    "dfd_resnet50_pack_weights": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _int, C.POINTER(_vp)]),
    "dfd_resnet50_free_weights": (None, [_vp]),
    "dfd_resnet50_workspace_bytes": (_int, [_i64, C.POINTER(C.c_size_t)]),
    "dfd_resnet50_score_videos": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dfd_gcn_pack_weights": (_int, [_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _int, C.POINTER(_vp)]),
    "dfd_gcn_free_weights": (None, [_vp]),
    "dfd_gcn_head": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "dfd_resize_last_error": (C.c_char_p, []),
    "dfd_crop_resize_workspace_bytes": (_int, [_vp, _i64, _int, C.POINTER(C.c_size_t)]),
    "dfd_crop_resize_u8": (_int, [_vp, _vp, _i64, _int, _vp, _vp, C.c_size_t, _vp]),
    "dfd_profile_enable": (_int, [_int]),
    "dfd_profile_collect": (_int, [_vp, _int, C.POINTER(_int)]),
    # dfd_b200_kernels.h
    "dfd_k_stem": (_int, [_vp, _int, _vp, _vp, _vp, _i64, _int, _int, _int, _vp]),
    "dfd_k_stem_tc": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp]),
    "dfd_k_set_dw_channel_block": (None, [_int]),
    "dfd_k_dw_num_partials": (_int, [_int, _int, _int, _int, _int]),
    "dfd_k_dwconv": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp]),
    "dfd_k_se": (_int, [_vp, _int, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp]),
    "dfd_k_gemm": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp]),
    "dfd_k_vit_attention": (_int, [_vp, _vp, _i64, _int, _vp]),
    "dfd_k_conv1x1_conv3x3": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp, C.c_size_t, _vp]),
    "dfd_k_mbconv_fused_supported": (_int, [_int, _int, _int, _int, _int, _int]),
    "dfd_k_mbconv_fused": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _int, _vp]),
    "dfd_k_conv3x3_maps": (_i64, [_int, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "dfd_k_pack_stem_row": (_int, [_vp, _vp, _vp, _vp]),
    "dfd_k_resize_coeffs": (_int, [_int, _int, _vp, _vp, C.POINTER(_int)]),
    "dfd_k_gemm_pool": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _int, _int, _int, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class CropBox(C.Structure):                      # dfd_crop_box
    _fields_ = [("frame_offset", C.c_int64), ("frame_w", C.c_int32), ("frame_h", C.c_int32),
                ("x1", C.c_int32), ("y1", C.c_int32), ("x2", C.c_int32), ("y2", C.c_int32)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int), ("ms", C.c_double), ("bytes", C.c_double), ("flops", C.c_double)]

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """dlopen the library once and attach the prototypes.  Raises RuntimeError when it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
                    "`python -m deepfake_video_detection_b200._build` (needs nvcc). There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError here = header/library mismatch
                fn.restype, fn.argtypes = res, args
            if lib.dfd_abi_version() != 1:
                raise RuntimeError("libdfd_b200.so: ABI version mismatch")
            _lib = lib
        return _lib


def check(rc: int, what: str = "") -> None:
    """Map a negative status to RuntimeError (the reference's callers catch plain exceptions, app.py:2320)."""
    if rc != 0:
        msg = load().dfd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"dfd_b200 {what} failed ({rc}): {msg}")
