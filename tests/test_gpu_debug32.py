"""The fp32-accumulate debug build of the forward (BASELINE.md section 4 / north star "<= 1e-5 fp32-accum"): with the
probabilities fed to the tensor core as P_hi + P_lo (about 22 mantissa bits), every exponential on the MUFU and the output
taken in float32, the kernel agrees with the float64 oracle to 1e-5 relative L2 -- i.e. the error of the product build
(3e-3 bf16 / 5e-4 fp16) is the 16-bit rounding of P and O, not the algorithm (online softmax with the +n term, lazy rescaling,
masking, dropout)."""
import ctypes

import pytest
import torch

from oracle import attention_oracle as orc
from tests._util import make_qkv

pytestmark = pytest.mark.gpu


def _fwd_f32(lib, q, k, v, n, causal, p=0.0, philox=(0, 0)):
    from flash_attention_softmax_n import _native
    from flash_attention_softmax_n.core.flash_attn import _fill_common
    B, H, L, D = q.shape
    o = torch.empty_like(q)
    o32 = torch.full((B, H, L, D), float("nan"), dtype=torch.float32, device=q.device)
    lse = torch.empty(B, H, L, dtype=torch.float32, device=q.device)
    prm = _native.FasnParams()
    _fill_common(prm, q, k, v, o, lse, H, n, D ** -0.5, causal, p, philox[0], philox[1], 0, None, None)
    prm.o_f32 = o32.data_ptr()
    rc = lib.fasn_fwd(ctypes.byref(prm))
    assert rc == 0, lib.fasn_last_error()
    torch.cuda.synchronize()
    return o, o32, lse


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("D,L,S,causal,n,p", [(128, 384, 384, True, 0.5, 0.0), (64, 300, 520, False, 1.0, 0.0), (128, 257, 640, True, 2.0, 0.25),
                                              (64, 512, 512, True, 0.0, 0.0)])
def test_debug_build_meets_1e5(fasn_debug32_lib, dtype, D, L, S, causal, n, p):
    B, H = 2, 2
    q, k, v, _ = make_qkv(B, H, L, S, D, dtype, seed=D + L)
    o16, o32, lse = _fwd_f32(fasn_debug32_lib, q, k, v, n, causal, p, (11, 3))
    keep = orc.dropout_keep_mask(11, 3, B, H, L, S, p) if p > 0 else None
    want, want_lse = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), softmax_n_param=n, is_causal=causal,
                                          keep_mask=keep, dropout_p=p, return_lse=True)
    rel = orc.rel_l2(o32, want)
    assert rel <= 1e-5, f"float32 output of the debug build: rel-L2 {rel:.2e}"
    finite = torch.isfinite(want_lse)
    assert (lse.double().cpu()[finite] - want_lse[finite]).abs().max().item() <= 2e-5 * max(1.0, want_lse[finite].abs().max().item())
    assert orc.rel_l2(o16, want) <= (3e-3 if dtype == torch.bfloat16 else 5e-4)      # and its 16-bit output is the usual one


def test_product_library_rejects_o_f32(fasn_lib):
    from flash_attention_softmax_n import _native
    q, k, v, _ = make_qkv(1, 1, 128, 128, 64, torch.float16, seed=1)
    with pytest.raises(AssertionError, match="debug library"):
        _fwd_f32(fasn_lib, q, k, v, 1.0, False)
