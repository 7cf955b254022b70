"""ctypes binding of libdkt_stereo_b200.so (the C ABI declared in include/dkt_stereo_b200.h).

PyTorch is only used for device memory and streams: every call passes raw device pointers
and the current CUDA stream.  There is no CPU fallback: if the library cannot be loaded the
import of the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import build as _build

c_f32p = C.c_void_p
c_u16p = C.c_void_p


class DktTensor(C.Structure):
    _fields_ = [("f32", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p),
                ("C", C.c_int32), ("c_begin", C.c_int32), ("c_count", C.c_int32)]


class DktEpilogue(C.Structure):
    _fields_ = [("kind", C.c_int32), ("act", C.c_int32), ("scale", C.c_float),
                ("bias", C.c_void_p), ("ctx", C.c_void_p), ("ctx_C", C.c_int32), ("ctx_c0", C.c_int32),
                ("out", DktTensor), ("z", DktTensor), ("h", DktTensor),
                ("tail", C.c_void_p), ("tail_C", C.c_int32),
                ("res", C.c_void_p), ("res_C", C.c_int32), ("res_c0", C.c_int32),
                ("proj", C.c_void_p), ("res_hi", C.c_void_p), ("res_lo", C.c_void_p),
                ("stats_partial", C.c_void_p)]


ABI_VERSION = 3      # include/dkt_stereo_b200.h: DKT_ABI_VERSION
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_LEAKY = 0, 1, 2, 3, 4
EPI_LINEAR, EPI_GRU_ZR, EPI_GRU_Q, EPI_PROJ = 0, 1, 2, 3
PROJ_LD = 12
MAX_LEVELS = 4

_I, _I64, _F, _P = C.c_int, C.c_int64, C.c_float, C.c_void_p
_TP = C.POINTER(DktTensor)
_EP = C.POINTER(DktEpilogue)

# name -> argtypes; mirrors include/dkt_stereo_b200.h one to one (tests check the list)
SIGNATURES = {
    "dkt_abi_version": [],
    "dkt_split_format": [],
    "dkt_error_string": [_I],
    "dkt_device_supported": [_I],
    "dkt_corr1d_build_f32": [_P, _P, _I64, _I64, _I64, _I64, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "dkt_corr1d_build_tc": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "dkt_corr1d_lookup": [_P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _I64, _I64, _I64, _I, _I, _I, _I, _P],
    "dkt_corr1d_lookup_enc_tc": [_P, _I, _I, _P, _P, _I, _P, _P, _P, _TP, _I, _I, _I, _I, _I, _P],
    "dkt_geo_pool_dc": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dkt_geo_lookup_enc_tc": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _TP, _I, _I, _I, _I, _P],
    "dkt_corr1d_lookup_backward": [_P, _P, _I, _P, _I, _I, _I, _I, _P],
    "dkt_geo_pool": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dkt_gwc_volume": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dkt_conv3d_c8": [_P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _P],
    "dkt_conv3d_k3": [_P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "dkt_deconv3d_k4s2": [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _P],
    "dkt_conv3d_k1": [_P, _I, _P, _I, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _P],
    "dkt_softargmin": [_P, _P, _I, _I, _I, _I, _P],
    "dkt_corr1d_lookup_enc": [_P, _I, _I, _P, _P, _I, _P, _P, _P, _TP, _I, _I, _I, _I, _P],
    "dkt_geo_lookup": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I64, _I64, _I64, _I, _I, _I, _P],
    "dkt_geo_lookup_enc": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _TP, _I, _I, _I, _P],
    "dkt_conv2d_simt": [_TP, _I, _P, _I, _I, _EP, _I, _I, _I, _P],
    "dkt_conv2d_tc": [_TP, _I, _P, _P, _I, _I, _EP, _I, _I, _I, _P],
    "dkt_conv2d_tc_ex": [_TP, _I, _P, _P, _I, _I, _I, _I, _EP, _I, _I, _I, _I, _I, _P],
    "dkt_pool2x": [_TP, _TP, _I, _I, _I, _I, _I, _P],
    "dkt_interp": [_TP, _TP, _I, _I, _I, _I, _I, _P],
    "dkt_convex_upsample": [_P, _I, _P, _P, _I, _I, _I, _I, _P],
    "dkt_context_upsample": [_P, _P, _P, _F, _F, _I, _I, _I, _P],
    "dkt_pixel_shuffle2": [_P, _I, _I, _TP, _I, _I, _I, _P],
    "dkt_context_upsample_logits": [_P, _P, _I, _P, _F, _F, _I, _I, _I, _P],
    "dkt_nchw_to_nhwc": [_P, _P, _TP, _I, _I, _I, _I, _P],
    "dkt_nhwc_to_nchw": [_TP, _P, _I, _I, _I, _I, _P],
    "dkt_dwconv3x3": [_TP, _P, _P, _F, _F, _F, _TP, _I, _I, _I, _I, _P],
    "dkt_ncdhw_to_ndhwc_pad": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dkt_ndhwc_pad_to_ncdhw": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dkt_stem_rows_bf16x2": [_P, _I64, _I64, _I64, _I64, _F, _F, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dkt_tapsum3x3": [_P, _I, _F, _P, _I, _I, _I, _I, _P],
    "dkt_instnorm_workspace_floats": [_I, _I],
    "dkt_instnorm_stats": [_TP, _P, _P, _F, _I, _I, _I, _P],
    "dkt_instnorm_apply": [_TP, _P, _TP, _TP, _I, _I, _I, _I, _P],
    "dkt_instnorm_tiles_workspace_floats": [_I, _I],
    "dkt_instnorm_finalize_tiles": [_P, _P, _P, _F, _I, _I, _I, _I, _P],
    "dkt_split_nchw_to_nhwc_bf16x2": [_P, _I64, _I64, _I64, _I64, _P, _P, _I, _I, _I, _I, _P],
}

_lib: Optional[C.CDLL] = None


class DktError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building in-tree first if the .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    override = os.environ.get("DKT_STEREO_LIB")      # A/B of two builds of the same ABI (tools/, never the default)
    if override:
        path = override
    elif os.path.exists(nvcc) and os.environ.get("DKT_NO_AUTOBUILD", "0") != "1":
        path = _build.build()
    if not os.path.exists(path):
        raise DktError(f"{path} is missing and cannot be built (no nvcc): the B200 engine has no CPU or "
                       "PyTorch fallback; run `python -m dkt_stereo_b200.build` on a machine with CUDA 12.9")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI drift; let it surface
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "dkt_error_string" else C.c_int
    if lib.dkt_abi_version() != ABI_VERSION:
        raise DktError("libdkt_stereo_b200.so ABI version mismatch")
    _lib = lib
    return lib


LAUNCHES = 0   # kernels launched through the C ABI by this process (every ok check() is one launch)


_split_dtype = None


def split_dtype() -> torch.dtype:
    """torch dtype of the library's 16-bit (hi, lo) planes: float16 (default build) or bfloat16 (DKT_SPLIT_FP16=0)."""
    global _split_dtype
    if _split_dtype is None:
        _split_dtype = torch.float16 if load().dkt_split_format() == 1 else torch.bfloat16
    return _split_dtype


def check(rc: int, what: str = "") -> None:
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = load().dkt_error_string(rc).decode()
        raise DktError(f"{what or 'dkt call'} failed with code {rc}: {msg}")


def require_device(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise DktError("the B200 engine only runs on CUDA tensors (no CPU fallback)")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def tensor_slice(f32: Optional[torch.Tensor] = None, hi: Optional[torch.Tensor] = None,
                 lo: Optional[torch.Tensor] = None, c_begin: int = 0, c_count: Optional[int] = None) -> DktTensor:
    """Describe an NHWC (B,H,W,C) buffer (any subset of precisions) as a dkt_tensor slice."""
    ref = f32 if f32 is not None else hi
    assert ref is not None and ref.is_contiguous()
    Cc = ref.shape[-1]
    sd = split_dtype() if (hi is not None or lo is not None) else None
    for t, dt in ((f32, torch.float32), (hi, sd), (lo, sd)):
        if t is not None:
            assert t.dtype == dt and t.is_contiguous() and t.shape[-1] == Cc, (t.dtype, t.shape)
    return DktTensor(ptr(f32), ptr(hi), ptr(lo), Cc, c_begin, Cc - c_begin if c_count is None else c_count)


def null_tensor() -> DktTensor:
    return DktTensor(None, None, None, 0, 0, 0)


def pointer_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * MAX_LEVELS)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
