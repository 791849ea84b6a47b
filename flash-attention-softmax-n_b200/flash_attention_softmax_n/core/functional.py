"""Eager softmax_n and attention: same names, arguments and results as the reference's
`flash_attention_softmax_n/core/functional.py` (`softmax_n` :15-29, `slow_attention_n` :32-93).

These are the un-fused definitions of the operator.  They are part of the package's import surface
(`from flash_attention_softmax_n import softmax_n, slow_attention_n`), run on whatever device their
inputs live on, and are never used by the fused path (`flash_attention_n`) as a fallback.
"""
from __future__ import annotations

import math
from typing import Optional, TYPE_CHECKING

import torch
from torch import Tensor

if TYPE_CHECKING:
    from torch.types import _dtype as DType
else:
    DType = int   # the reference exports this alias (functional.py:8-13); tests import it


def softmax_n(x: Tensor, n: Optional[float] = None, dim: Optional[int] = None, dtype: Optional[DType] = None) -> Tensor:
    """softmax_n(x)_i = exp(x_i) / (n + sum_j exp(x_j)).

    Not shift invariant for n != 0: subtracting the row maximum m for stability turns the constant
    into n * exp(-m).  The maximum carries no gradient."""
    n = 0.0 if n is None else n
    dim = -1 if dim is None else dim
    m = torch.amax(x, dim=dim, keepdim=True).detach()
    e = torch.exp(x - m)
    y = e / (e.sum(dim=dim, keepdim=True) + n * torch.exp(-m))
    return y if dtype is None else y.type(dtype=dtype)


def slow_attention_n(query: Tensor, key: Tensor, value: Tensor, attn_mask: Optional[Tensor] = None,
                     dropout_p: float = 0.0, is_causal: bool = False, scale: Optional[float] = None,
                     softmax_n_param: Optional[float] = None, softmax_dtype: Optional[DType] = None,
                     train: bool = True) -> Tensor:
    """Attention that materialises the (L, S) score matrix.

    query (N,...,L,E), key (N,...,S,E), value (N,...,S,Ev) -> (N,...,L,Ev).
    `attn_mask`: boolean (True = attend) broadcastable to the scores, or a float tensor added to them.
    `is_causal`: bottom-right aligned, row i sees keys j <= i + (S - L); exclusive with `attn_mask`.
    Scores are formed in the input dtype; dropout acts on the normalised weights.

    Unlike the reference (functional.py:85-86), a boolean mask is applied (and not modified)."""
    n = 0.0 if softmax_n_param is None else softmax_n_param
    out_dtype = query.dtype if softmax_dtype is None else softmax_dtype
    L, S = query.size(-2), key.size(-2)
    sm_scale = scale if scale is not None else 1.0 / math.sqrt(query.size(-1))

    scores = torch.matmul(query, key.transpose(-2, -1)) * sm_scale
    if is_causal:
        assert attn_mask is None, "is_causal and attn_mask are mutually exclusive"
        rows = torch.arange(L, device=query.device).unsqueeze(-1)
        cols = torch.arange(S, device=query.device).unsqueeze(0)
        scores = scores.masked_fill(cols > rows + (S - L), float("-inf"))
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            scores = scores.masked_fill(~attn_mask, float("-inf"))
        else:
            scores = scores + attn_mask
    weights = softmax_n(scores, n=n, dim=-1, dtype=out_dtype)
    weights = torch.dropout(weights, dropout_p, train=train)
    return torch.matmul(weights, value)
