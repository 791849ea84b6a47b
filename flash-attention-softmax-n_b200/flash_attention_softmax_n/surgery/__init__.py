"""Callers of the fused path: route Hugging Face attention modules to `flash_attention_n`.
Same entry-point name as the reference's `flash_attention_softmax_n.surgery` (surgery/__init__.py:1-5)."""
from flash_attention_softmax_n.surgery.attention_softmax_n import (
    FUSED, EAGER, AttentionSoftmaxN, apply_attention_softmax_n, attention_softmax_n_forward, eager_attention_softmax_n_forward,
    register_attention_softmax_n,
)

__all__ = ["FUSED", "EAGER", "AttentionSoftmaxN", "apply_attention_softmax_n", "attention_softmax_n_forward",
           "eager_attention_softmax_n_forward", "register_attention_softmax_n"]
