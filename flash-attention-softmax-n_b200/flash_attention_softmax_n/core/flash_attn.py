"""`flash_attention_n`: the reference's 9-argument entry point
(flash_attention_softmax_n/core/flash_attn.py:42-52) on hand-written sm_100a kernels.

Same names, order, defaults and meaning of the arguments; what happens underneath differs:

* the reference pads K/V with n zero rows and calls `scaled_dot_product_attention` (flash_attn.py:66-67,
  115-124), which restricts n to integers and, on a B200, selects the math / mem-efficient backends with a
  materialised (B,H,L,S+n) additive mask (flash_attn.py:28-33, 97-113).  Here the "+n" is the initial value
  of the online-softmax running sum inside one fused kernel, n is any real >= 0, and causality is an index
  comparison in registers;
* there is no backend chooser (`_flash_attn_config`, flash_attn.py:17-35) and no fallback: inputs the
  kernels do not cover raise `NotImplementedError` (use `slow_attention_n` for those).

Covered: CUDA tensors, float16 / bfloat16 (float32 tensors are computed with float16 operands and float32 accumulation
and returned in float32, see `flash_attention_n`), head dims up to 128 (64 and 128 run unpadded), 4-D query, 4-D or 3-D
key/value (3-D = shared by all heads, flash_attn.py:75-79), any L and S, boolean `attn_mask` (True = attend),
`attn_bias` of shape (H,L,S) or 4-D broadcastable, `is_causal` bottom-right aligned, dropout.
Gradients flow to query, key, value and `attn_bias`.
"""
from __future__ import annotations

import ctypes
from math import sqrt
from typing import Optional, Tuple

import torch
from torch import Tensor

from flash_attention_softmax_n import _native

_SUPPORTED_HEAD_DIMS = (64, 128)


def _next_philox(device: torch.device) -> Tuple[int, int]:
    """(seed, offset) from torch's CUDA generator, advancing it so every call gets a fresh stream."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = gen.initial_seed()
    offset = gen.get_offset()
    gen.set_offset(offset + 4)
    return seed & 0xFFFFFFFFFFFFFFFF, offset // 4


def _rowmajor(t: Tensor) -> Tensor:
    """Kernels need unit stride on the last axis, 16-byte aligned rows and base pointer."""
    if t.stride(-1) != 1 or any(s % 8 for s in t.stride()[:-1]) or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def _canon_aux(t: Tensor, S: int) -> Tensor:
    """Mask / bias layout the kernels read: unit stride along keys; batch, head and query axes may be
    broadcast (size 1 or stride 0), which the C ABI expresses as a zero stride."""
    if t.size(-1) != S:                       # size-1 key axis: give it its real extent
        t = t.expand(*t.shape[:-1], S)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _prepare_aux(attn_mask: Optional[Tensor], attn_bias: Optional[Tensor], q: Tensor, S: int
                 ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    B, H, L, _ = q.shape
    if attn_mask is not None:
        assert attn_mask.ndim == 4                                           # flash_attn.py:88
        if attn_mask.dtype != torch.bool:
            raise TypeError("attn_mask must be a boolean tensor (True = take part in attention)")
        attn_mask.expand(B, H, L, S)                                         # shape check only (flash_attn.py:89)
        attn_mask = _canon_aux(attn_mask, S)
    if attn_bias is not None:
        if attn_bias.ndim == 3:
            attn_bias = attn_bias.unsqueeze(0)                               # 'h i j -> 1 h i j' (flash_attn.py:101-102)
        assert attn_bias.ndim == 4
        attn_bias.expand(B, H, L, S)                                         # shape check (flash_attn.py:103)
        attn_bias = _canon_aux(attn_bias.to(q.dtype), S)
    return attn_mask, attn_bias


def _fill_common(p: _native.FasnParams, q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, heads_kv: int,
                 n: float, scale: float, causal: bool, dropout_p: float, seed: int, offset: int, bh_offset: int,
                 mask: Optional[Tensor], bias: Optional[Tensor], alibi: Optional[Tensor] = None) -> None:
    B, H, L, D = q.shape
    p.struct_size = ctypes.sizeof(_native.FasnParams)
    p.dtype = _native.dtype_code(q.dtype)
    p.batch, p.heads, p.heads_kv = B, H, heads_kv
    p.seqlen_q, p.seqlen_kv, p.head_dim = L, k.shape[2], D
    p.q, p.k, p.v, p.o = (_native.tensor_view(t) for t in (q, k, v, o))
    p.lse = lse.data_ptr()
    p.softmax_n, p.scale, p.is_causal, p.dropout_p = float(n), float(scale), int(bool(causal)), float(dropout_p)
    p.philox_seed, p.philox_offset, p.bh_offset = seed, offset, bh_offset
    p.mask, p.bias = _native.aux_view(mask), _native.aux_view(bias)
    p.alibi_slopes = alibi.data_ptr() if alibi is not None else None
    p.stream = _native.current_stream_ptr(q.device)


class _FusedAttentionN(torch.autograd.Function):
    """Counterpart of the reference's `_FlashAttentionN` (flash_attn_triton.py:241-336): saves
    (q, k, v, o, lse) and returns gradients for q, k, v -- and for a dense `attn_bias` that requires one, which the
    reference's SDPA route provides through aten autograd (flash_attn.py:100-124)."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, heads_kv: int, n: float, scale: float, causal: bool,
                dropout_p: float, mask: Optional[Tensor], bias: Optional[Tensor], seed: int, offset: int,
                bh_offset: int, alibi: Optional[Tensor] = None) -> Tensor:
        lib = _native.load()
        B, H, L, D = q.shape
        with torch.cuda.device(q.device):
            o = torch.empty((B, H, L, D), dtype=q.dtype, device=q.device)     # fresh outputs (flash_attn_triton.py:271)
            lse = torch.empty((B, H, L), dtype=torch.float32, device=q.device)
            p = _native.FasnParams()
            _fill_common(p, q, k, v, o, lse, heads_kv, n, scale, causal, dropout_p, seed, offset, bh_offset, mask, bias, alibi)
            _native.check(lib.fasn_fwd(ctypes.byref(p)), "fasn_fwd")
        ctx.save_for_backward(q, k, v, o, lse, mask, bias, alibi)
        ctx.cfg = (heads_kv, n, scale, causal, dropout_p, seed, offset, bh_offset)
        return o

    @staticmethod
    def backward(ctx, do: Tensor):
        q, k, v, o, lse, mask, bias, alibi = ctx.saved_tensors
        heads_kv, n, scale, causal, dropout_p, seed, offset, bh_offset = ctx.cfg
        lib = _native.load()
        B, H, L, D = q.shape
        S = k.shape[2]
        do = _rowmajor(do)
        Lp = (L + 127) // 128 * 128
        with torch.cuda.device(q.device):
            dq = torch.empty((B, H, L, D), dtype=q.dtype, device=q.device)
            # dK / dV are produced per query head; with shared K/V (heads_kv == 1) the kernel adds every head's contribution
            # into float32 (B,1,S,D) accumulators instead (FasnParams.dk_accum / dv_accum)
            head_sum = heads_kv == 1 and H > 1
            if head_sum:
                dk_acc = torch.zeros((B, 1, S, D), dtype=torch.float32, device=q.device)
                dv_acc = torch.zeros((B, 1, S, D), dtype=torch.float32, device=q.device)
            else:
                dk = torch.empty((B, H, S, D), dtype=q.dtype, device=q.device)
                dv = torch.empty((B, H, S, D), dtype=q.dtype, device=q.device)
            ws = torch.empty((2, B, H, Lp), dtype=torch.float32, device=q.device)
            dq_accum = torch.empty((B, H, Lp, D), dtype=torch.float32, device=q.device)
            p = _native.FasnParams()
            _fill_common(p, q, k, v, o, lse, heads_kv, n, scale, causal, dropout_p, seed, offset, bh_offset, mask, bias, alibi)
            p.dout, p.dq = _native.tensor_view(do), _native.tensor_view(dq)
            if head_sum:
                p.dk_accum, p.dv_accum = dk_acc.data_ptr(), dv_acc.data_ptr()
            else:
                p.dk, p.dv = _native.tensor_view(dk), _native.tensor_view(dv)
            p.delta, p.dq_accum = ws.data_ptr(), dq_accum.data_ptr()
            ds = None
            if bias is not None and ctx.needs_input_grad[9]:
                # dS for every (b, h, i, j); entries above the causal diagonal are never visited by a CTA and stay zero
                ds = torch.zeros((B, H, L, S), dtype=q.dtype, device=q.device)
                p.dbias, p.dbias_stride_b, p.dbias_stride_h, p.dbias_stride_q = ds.data_ptr(), H * L * S, L * S, S
            _native.check(lib.fasn_bwd(ctypes.byref(p)), "fasn_bwd")
        if head_sum:
            dk, dv = dk_acc.to(q.dtype), dv_acc.to(q.dtype)
        dbias = None
        if ds is not None:
            # reduce over the axes the bias broadcasts (size 1), accumulating in float32
            dims = [d for d in range(4) if bias.shape[d] == 1 and ds.shape[d] != 1]
            dbias = (ds.sum(dim=dims, keepdim=True, dtype=torch.float32) if dims else ds).to(bias.dtype)
        return dq, dk, dv, None, None, None, None, None, None, dbias, None, None, None, None


def flash_attention_n(
        query: Tensor,
        key: Tensor,
        value: Tensor,
        softmax_n_param: Optional[float] = None,
        scale: Optional[float] = None,
        dropout_p: float = 0.,
        attn_mask: Optional[Tensor] = None,
        attn_bias: Optional[Tensor] = None,
        is_causal: bool = False,
        *,
        _philox: Optional[Tuple[int, int]] = None,
        _bh_offset: int = 0,
        _alibi_slopes: Optional[Tensor] = None,
) -> Tensor:
    """
    Fused attention with softmax_n on B200.

    :param query: Query tensor; shape (N, H, L, E).
    :param key: Key tensor; shape (N, H, S, E) or (N, S, E) (shared by all heads).
    :param value: Value tensor; shape (N, H, S, Ev) or (N, S, Ev).  E, Ev <= 128 (64 and 128 run unpadded).
        query / key / value: float16, bfloat16, or float32 (computed with float16 operands, float32 accumulation).
    :param softmax_n_param: Regularization parameter n >= 0 of softmax_n (any real number; None = 0).
    :param scale: Scaling factor applied prior to softmax. If None, the default value is set to 1 / sqrt(E).
    :param dropout_p: Dropout probability; if greater than 0.0, dropout is applied.
    :param attn_mask: Boolean attention mask, 4-D, broadcastable to (N, H, L, S); True = attend.
    :param attn_bias: Additive (e.g. ALiBi) bias; shape (H, L, S) or 4-D broadcastable to (N, H, L, S).
    :param is_causal: If true, causal masking aligned to the bottom-right corner (row i sees j <= i + S - L).
    :return: Attention output; shape (N, H, L, Ev).

    `_philox` (seed, offset) pins the dropout stream and `_bh_offset` is the global index of the first
    (batch, head) unit of this call; both exist for tests and for batch x head sharding (parallel.py).
    `_alibi_slopes` (H,) generates the ALiBi bias  slopes[h] * (j - i - (S - L))  inside the kernels, equivalent to (and
    exclusive with) passing that (H, L, S) tensor as `attn_bias`, without its B*H*L*S elements of HBM traffic.
    """
    n = 0.0 if softmax_n_param is None else float(softmax_n_param)
    if n < 0:
        raise ValueError("softmax_n_param must be >= 0")
    if query.ndim != 4:
        raise ValueError(f"query must be 4-D (N, H, L, E), got {tuple(query.shape)}")        # flash_attn.py:85
    if not query.is_cuda:
        raise NotImplementedError("flash_attention_n runs on CUDA (B200) tensors only; there is no CPU path. "
                                  "Use slow_attention_n for eager evaluation.")
    if query.dtype == torch.float32 and key.dtype == torch.float32 and value.dtype == torch.float32:
        # float32 tensors (supported by the reference's SDPA route, README.md:42; its GPU test asks atol 1e-3,
        # tests/gpu/core/test_flash_attn.py:14): the tensor cores take 16-bit operands, so the inputs are rounded to float16
        # -- the 10-bit mantissa a `kind::tf32` MMA would keep as well -- products are accumulated in float32, and the result
        # and the gradients come back in float32 (autograd carries them through the casts).  Values beyond the float16
        # range (|x| > 65504) overflow; use bfloat16 tensors or slow_attention_n for such inputs.
        f16 = lambda t: None if t is None else (t.to(torch.float16) if t.is_floating_point() else t)
        out = flash_attention_n(f16(query), f16(key), f16(value), softmax_n_param, scale, dropout_p, attn_mask, f16(attn_bias),
                                is_causal, _philox=_philox, _bh_offset=_bh_offset, _alibi_slopes=_alibi_slopes)
        return out.to(torch.float32)
    if query.dtype not in (torch.float16, torch.bfloat16) or key.dtype != query.dtype or value.dtype != query.dtype:
        raise NotImplementedError(f"fused kernel supports float16 / bfloat16 / float32 with matching dtypes, got "
                                  f"{query.dtype}/{key.dtype}/{value.dtype}; use slow_attention_n")
    B, H, L, E = query.shape
    heads_kv = H
    if key.ndim == 3:                                                     # 'b ... -> b 1 ...' (flash_attn.py:75-76)
        key = key.unsqueeze(1)
    if value.ndim == 3:
        value = value.unsqueeze(1)
    if key.ndim != 4 or value.ndim != 4:
        raise ValueError("key and value must be 3-D or 4-D")
    if key.shape[1] != value.shape[1] or key.shape[1] not in (1, H):
        raise ValueError(f"key/value heads {key.shape[1]}/{value.shape[1]} must both be {H} or 1")
    if key.shape[1] == 1 and H > 1:
        heads_kv = 1
    S = key.shape[2]
    if key.shape[0] != B or value.shape[0] != B or value.shape[2] != S or key.shape[3] != E:
        raise ValueError("inconsistent query/key/value shapes")
    Ev = value.shape[3]
    if max(E, Ev) > max(_SUPPORTED_HEAD_DIMS):
        raise NotImplementedError(f"fused kernel supports head dims up to {max(_SUPPORTED_HEAD_DIMS)}, got E={E}, Ev={Ev}; "
                                  "use slow_attention_n")
    # The kernels are built for E == Ev in {64, 128}.  Other head dims (the Triton path's 16 / 32, flash_attn_triton.py:266;
    # Ev != E on the SDPA path, README.md:50) are zero-padded up to the next supported size: zero feature columns change
    # neither Q K^T nor the first Ev output columns, and autograd carries the gradients through the pad / slice.
    Dp = min(d for d in _SUPPORTED_HEAD_DIMS if d >= max(E, Ev))
    if not 0.0 <= dropout_p < 1.0:
        raise ValueError("dropout_p must be in [0, 1)")
    if dropout_p > 0.0 and round((1.0 - dropout_p) * 256.0) < 1:
        # the counter-based generator decides with 8-bit uniforms: P(keep) = round(256 (1-p)) / 256, kept entries are
        # scaled by its inverse; a keep probability that rounds to 0 has no unbiased estimator
        raise ValueError(f"dropout_p={dropout_p} is closer to 1 than the dropout generator resolves (1/256)")
    sm_scale = 1.0 / sqrt(E) if scale is None else float(scale)           # flash_attn.py:59, 81-83

    if E != Dp:
        query = torch.nn.functional.pad(query, (0, Dp - E))
        key = torch.nn.functional.pad(key, (0, Dp - E))
    if Ev != Dp:
        value = torch.nn.functional.pad(value, (0, Dp - Ev))
    query, key, value = _rowmajor(query), _rowmajor(key), _rowmajor(value)
    mask, bias = _prepare_aux(attn_mask, attn_bias, query, S)
    alibi = None
    if _alibi_slopes is not None:
        if bias is not None:
            raise ValueError("_alibi_slopes and attn_bias are mutually exclusive")
        if _alibi_slopes.shape != (H,):
            raise ValueError(f"_alibi_slopes must have shape ({H},), got {tuple(_alibi_slopes.shape)}")
        alibi = _alibi_slopes.detach().to(device=query.device, dtype=torch.float32).contiguous()
    seed, offset = (0, 0)
    if dropout_p > 0.0:
        if _philox is None and torch.cuda.is_current_stream_capturing():
            # (seed, offset) are host integers in the kernel arguments: a captured graph would replay ONE dropout mask for ever
            raise RuntimeError("flash_attention_n with dropout_p > 0 cannot be captured into a CUDA graph: the dropout stream "
                               "(seed, offset) is drawn on the host per call; capture with dropout_p = 0 or run this call eagerly")
        seed, offset = _philox if _philox is not None else _next_philox(query.device)
    out = _FusedAttentionN.apply(query, key, value, heads_kv, n, sm_scale, bool(is_causal), float(dropout_p),
                                 mask, bias, int(seed), int(offset), int(_bh_offset), alibi)
    return out if Ev == Dp else out[..., :Ev]
