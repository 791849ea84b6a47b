"""B200-native drop-in for the hot path of `flash_attention_softmax_n` (reference __init__.py:1-12):
the same names at the same module paths, with `flash_attention_n` / `flash_attention_n_triton` served by
hand-written sm_100a kernels behind a C ABI (libfasn.so)."""
from flash_attention_softmax_n.core.flash_attn import flash_attention_n
from flash_attention_softmax_n.core.functional import softmax_n, slow_attention_n
from flash_attention_softmax_n.core.flash_attn_triton import flash_attention_n_triton
from flash_attention_softmax_n.core.softmax import softmax_n_fused      # not in the reference: fused CUDA softmax_n

# The reference sets this when `import triton` succeeds (__init__.py:5-9); here it means
# "flash_attention_n_triton is importable", which is always true.
TRITON_INSTALLED = True

__all__ = ["flash_attention_n", "softmax_n", "slow_attention_n", "flash_attention_n_triton", "TRITON_INSTALLED",
           "softmax_n_fused"]
