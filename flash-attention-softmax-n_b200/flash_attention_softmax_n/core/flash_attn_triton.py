"""`flash_attention_n_triton`: same signature as the reference's Triton entry point
(flash_attention_softmax_n/core/flash_attn_triton.py:339-357), served by the same sm_100a kernels as
`flash_attention_n`.  No Triton is involved; the name is kept so that callers do not change.

Differences from the reference kernel, all in the direction of the operator's definition:
real-valued n is exact in forward AND backward (the reference's epilogue mixes exp bases, :83 vs :114,
and its backward recomputes P without n, :116,:210-211); sequence lengths need not be multiples of
128/64; bf16 is accepted; dQ is returned in the input dtype.
"""
from __future__ import annotations

from typing import Optional

from torch import Tensor

from flash_attention_softmax_n.core.flash_attn import flash_attention_n


def flash_attention_n_triton(query: Tensor, key: Tensor, value: Tensor, is_causal: bool = False,
                             scale: Optional[float] = None, softmax_n_param: Optional[float] = None) -> Tensor:
    assert query.shape[-1] == key.shape[-1] == value.shape[-1]       # flash_attn_triton.py:264-265
    return flash_attention_n(query, key, value, softmax_n_param=softmax_n_param, scale=scale, is_causal=is_causal)
