"""ctypes binding of libfasn.so (C ABI declared in include/fasn.h).

The library is loaded lazily, on the first kernel call, so `import flash_attention_softmax_n` works on a
machine without a GPU or without the built library (the reference's own import crashes there,
flash_attn_triton.py:238).  There is NO fallback: if the library cannot be loaded, calling a fused
kernel raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

FASN_ABI_VERSION = 3
FASN_FP16, FASN_BF16 = 0, 1

_LIB_ENV = "FASN_LIBRARY"
_HERE = os.path.dirname(os.path.abspath(__file__))


class FasnTensor(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_b", ctypes.c_int64), ("stride_h", ctypes.c_int64),
                ("stride_s", ctypes.c_int64)]


class FasnAux(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_b", ctypes.c_int64), ("stride_h", ctypes.c_int64),
                ("stride_q", ctypes.c_int64)]


class FasnParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("dtype", ctypes.c_uint32),
        ("batch", ctypes.c_int32), ("heads", ctypes.c_int32), ("heads_kv", ctypes.c_int32),
        ("seqlen_q", ctypes.c_int32), ("seqlen_kv", ctypes.c_int32), ("head_dim", ctypes.c_int32),
        ("q", FasnTensor), ("k", FasnTensor), ("v", FasnTensor), ("o", FasnTensor),
        ("lse", ctypes.c_void_p),
        ("dout", FasnTensor), ("dq", FasnTensor), ("dk", FasnTensor), ("dv", FasnTensor),
        ("delta", ctypes.c_void_p), ("dq_accum", ctypes.c_void_p),
        ("softmax_n", ctypes.c_float), ("scale", ctypes.c_float), ("is_causal", ctypes.c_int32),
        ("dropout_p", ctypes.c_float),
        ("philox_seed", ctypes.c_uint64), ("philox_offset", ctypes.c_uint64), ("bh_offset", ctypes.c_int64),
        ("mask", FasnAux), ("bias", FasnAux),
        ("stream", ctypes.c_void_p),
        ("alibi_slopes", ctypes.c_void_p),
        ("dbias", ctypes.c_void_p), ("dbias_stride_b", ctypes.c_int64), ("dbias_stride_h", ctypes.c_int64),
        ("dbias_stride_q", ctypes.c_int64),
        ("o_f32", ctypes.c_void_p),
        ("dk_accum", ctypes.c_void_p), ("dv_accum", ctypes.c_void_p),
    ]


EXPORTS = ("fasn_version", "fasn_last_error", "fasn_fwd", "fasn_bwd", "fasn_bwd_workspace",
           "fasn_dropout_mask", "fasn_probe", "fasn_attention_host", "fasn_profile", "fasn_profile_read",
           "fasn_softmax_n_fwd", "fasn_softmax_n_bwd", "fasn_copy_async")

_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


class FasnError(RuntimeError):
    pass


def library_path() -> str:
    return os.environ.get(_LIB_ENV) or os.path.join(_HERE, "libfasn.so")


def load() -> ctypes.CDLL:
    """Load libfasn.so (once) and declare the prototypes.  Raises FasnError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise FasnError(
                f"libfasn.so not found at {path}: build it with `python flash-attention-softmax-n_b200/build.py` "
                f"(or __graft_entry__.build()).  There is no CPU or PyTorch fallback for the fused kernels.")
        lib = ctypes.CDLL(path)
        lib.fasn_version.restype = ctypes.c_int
        lib.fasn_last_error.restype = ctypes.c_char_p
        lib.fasn_fwd.argtypes = [ctypes.POINTER(FasnParams)]
        lib.fasn_fwd.restype = ctypes.c_int
        lib.fasn_bwd.argtypes = [ctypes.POINTER(FasnParams)]
        lib.fasn_bwd.restype = ctypes.c_int
        lib.fasn_bwd_workspace.argtypes = [ctypes.POINTER(FasnParams), ctypes.POINTER(ctypes.c_uint64),
                                           ctypes.POINTER(ctypes.c_uint64)]
        lib.fasn_bwd_workspace.restype = ctypes.c_int
        lib.fasn_dropout_mask.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64,
                                          ctypes.c_int64, ctypes.c_void_p]
        lib.fasn_dropout_mask.restype = ctypes.c_int
        lib.fasn_probe.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]
        lib.fasn_probe.restype = ctypes.c_int
        lib.fasn_attention_host.argtypes = [
            ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64,
            ctypes.c_void_p]
        lib.fasn_attention_host.restype = ctypes.c_int
        lib.fasn_softmax_n_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
        lib.fasn_softmax_n_fwd.restype = ctypes.c_int
        lib.fasn_softmax_n_bwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_void_p]
        lib.fasn_softmax_n_bwd.restype = ctypes.c_int
        lib.fasn_profile.argtypes = [ctypes.c_int]
        lib.fasn_profile.restype = ctypes.c_int
        lib.fasn_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32),
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)]
        lib.fasn_profile_read.restype = ctypes.c_int
        lib.fasn_copy_async.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        lib.fasn_copy_async.restype = ctypes.c_int
        v = lib.fasn_version()
        if v != FASN_ABI_VERSION:
            raise FasnError(f"libfasn.so ABI version {v} != binding version {FASN_ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fasn_last_error().decode("utf-8", "replace")
        raise FasnError(f"{what} failed (code {rc}): {msg}")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return FASN_FP16
    if dtype == torch.bfloat16:
        return FASN_BF16
    raise FasnError(f"fused softmax_n attention supports float16 and bfloat16 inputs, got {dtype}")


def tensor_view(t: torch.Tensor) -> FasnTensor:
    """(B,H,S,D) tensor -> FasnTensor (element strides; the last dimension must be contiguous)."""
    assert t.ndim == 4 and t.stride(-1) == 1
    return FasnTensor(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def aux_view(t: Optional[torch.Tensor]) -> FasnAux:
    """(B|1,H|1,L,S) mask or bias -> FasnAux; broadcast axes get stride 0."""
    if t is None:
        return FasnAux(None, 0, 0, 0)
    assert t.ndim == 4 and t.stride(-1) == 1
    sb = 0 if t.size(0) == 1 else t.stride(0)
    sh = 0 if t.size(1) == 1 else t.stride(1)
    sq = 0 if t.size(2) == 1 else t.stride(2)
    return FasnAux(t.data_ptr(), sb, sh, sq)


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
