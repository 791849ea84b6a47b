"""Fused softmax_n over one axis (CUDA): `softmax_n_fused(x, n, dim, dtype)` has the arguments and results of the
reference's `softmax_n` (flash_attention_softmax_n/core/functional.py:15-29) -- softmax_n(x)_i = exp(x_i) / (n + sum_j exp(x_j)),
gradient through x only (the stabilising shift carries none) -- but reads each row once and writes it once instead of
four elementwise passes.  It is what a model that cannot use the fused attention (relative-position scores, head masks;
the reference's surgery: surgery_functions/_bert.py:101, _xlnet.py:62) swaps in for `softmax_n`.

CUDA tensors in float16 / bfloat16 / float32 only; no CPU fallback (the eager `softmax_n` stays the definition)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from flash_attention_softmax_n import _native

_CODES = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


def _pair_ok(din: torch.dtype, dout: torch.dtype) -> bool:
    return din in _CODES and dout in _CODES and (din == dout or din == torch.float32 or dout == torch.float32)


def _rows(t: Tensor):
    """(…, C) tensor with unit last stride -> (2-D view, rows, row stride); copies only if the leading axes do not collapse."""
    c = t.shape[-1]
    t2 = t.reshape(-1, c) if t.is_contiguous() else None
    if t2 is None:
        try:
            t2 = t.view(-1, c)
        except RuntimeError:
            t2 = t.contiguous().view(-1, c)
    if t2.stride(-1) != 1 and t2.shape[-1] > 1:
        t2 = t2.contiguous()
    return t2, t2.shape[0], (t2.stride(0) if t2.shape[0] > 1 else c)


class _SoftmaxN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, n: float, out_dtype: torch.dtype) -> Tensor:
        lib = _native.load()
        x2, rows, sx = _rows(x)
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        y2 = y.view(-1, x.shape[-1])
        with torch.cuda.device(x.device):
            _native.check(lib.fasn_softmax_n_fwd(x2.data_ptr(), y2.data_ptr(), rows, x.shape[-1], sx, x.shape[-1], _CODES[x.dtype],
                                                 _CODES[out_dtype], float(n), _native.current_stream_ptr(x.device)), "fasn_softmax_n_fwd")
        ctx.save_for_backward(y)
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        (y,) = ctx.saved_tensors
        lib = _native.load()
        c = y.shape[-1]
        dy2, rows, sdy = _rows(dy.to(y.dtype))
        dx = torch.empty(y.shape, dtype=ctx.in_dtype, device=y.device)
        with torch.cuda.device(y.device):
            _native.check(lib.fasn_softmax_n_bwd(y.view(-1, c).data_ptr(), dy2.data_ptr(), dx.view(-1, c).data_ptr(), rows, c, c, sdy, c,
                                                 _CODES[ctx.in_dtype], _CODES[y.dtype], _native.current_stream_ptr(y.device)),
                          "fasn_softmax_n_bwd")
        return dx, None, None


def softmax_n_fused(x: Tensor, n: Optional[float] = None, dim: Optional[int] = None, dtype: Optional[torch.dtype] = None) -> Tensor:
    """Drop-in for `softmax_n(x, n, dim, dtype)` on CUDA tensors (same argument meaning and defaults, functional.py:15-29)."""
    if not x.is_cuda:
        raise NotImplementedError("softmax_n_fused runs on CUDA tensors only (use softmax_n for the eager definition)")
    n = 0.0 if n is None else float(n)
    if not n >= 0.0:
        raise ValueError("softmax_n parameter must be >= 0")
    dim = -1 if dim is None else dim
    out_dtype = x.dtype if dtype is None else dtype
    if not _pair_ok(x.dtype, out_dtype):
        raise NotImplementedError(f"softmax_n_fused: unsupported dtypes {x.dtype} -> {out_dtype} (float16, bfloat16, float32; "
                                  "mixed pairs must involve float32)")
    if x.ndim == 0:
        raise ValueError("softmax_n_fused needs at least one axis")
    if x.numel() == 0:
        return torch.empty(x.shape, dtype=out_dtype, device=x.device)
    dim = dim % x.ndim
    if dim != x.ndim - 1:
        return _SoftmaxN.apply(x.transpose(dim, -1).contiguous(), n, out_dtype).transpose(dim, -1)
    if x.stride(-1) != 1 and x.shape[-1] > 1:
        x = x.contiguous()
    return _SoftmaxN.apply(x, n, out_dtype)
