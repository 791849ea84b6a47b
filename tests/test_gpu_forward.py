"""Forward parity of the fused sm_100a kernel (through flash_attention_n -> ctypes -> fasn_fwd) against the oracle."""
import math

import pytest
import torch

from oracle import attention_oracle as orc
from tests._util import make_qkv, oracle_all, native_lowp_all, check_close, REL_L2

pytestmark = pytest.mark.gpu

CASES = [
    # B  H  L     S     D    n     causal scale
    (1, 1, 128, 128, 64, 1.0, False, None),      # BASELINE.json configs[0] shape
    (2, 3, 256, 256, 128, 0.5, True, None),
    (2, 2, 384, 384, 64, 0.0, False, 0.1),
    (1, 2, 200, 333, 128, 4.0, False, 0.5),      # ragged L and S
    (1, 2, 333, 200, 64, 1.0, True, None),       # causal with S < L: leading rows see no key
    (2, 1, 96, 160, 64, 0.5, True, None),        # causal with S > L (bottom-right aligned)
    (1, 1, 1, 77, 128, 1.0, False, None),        # single query row
    (1, 2, 1024, 1152, 64, 1e-3, True, 0.3),     # the reference's GPU analytic-test shape
    (1, 1, 640, 640, 128, 1e-6, True, None),
    (5, 8, 512, 384, 128, 0.5, True, None),      # 40 units: both halves of the launch order (unit-major head, tile-major tail)
    (7, 6, 300, 300, 64, 1.0, True, None),       # 42 units, ragged
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,S,D,n,causal,scale", CASES)
def test_forward_matches_oracle(fasn_lib, B, H, L, S, D, n, causal, scale, dtype):
    from flash_attention_softmax_n import flash_attention_n
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=L + S)
    kw = dict(softmax_n_param=n, scale=scale, is_causal=causal)
    out = flash_attention_n(q, k, v, **kw)
    torch.cuda.synchronize()
    assert out.shape == (B, H, L, D) and out.dtype == dtype
    want = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), **kw)
    native = orc.slow_attention_n(q, k, v, **kw)
    check_close("O", out, want, native, dtype)


@pytest.mark.parametrize("name", ["c1", "causal", "n0c"])
def test_forward_golden_vectors(fasn_lib, golden, name):
    """Inputs and outputs produced by the reference's own slow_attention_n (tests/golden/make_golden.py);
    the inputs are exactly representable in bf16, so the kernel sees the very same numbers."""
    from flash_attention_softmax_n import flash_attention_n
    n, scale, causal = golden[f"slow_{name}_meta"]
    for dtype in (torch.bfloat16, torch.float16):
        q, k, v = (torch.from_numpy(golden[f"slow_{name}_{x}"]).to(dtype).cuda() for x in "qkv")
        out = flash_attention_n(q, k, v, softmax_n_param=float(n), scale=None if scale < 0 else float(scale),
                                is_causal=bool(causal))
        want = torch.from_numpy(golden[f"slow_{name}_o"]).double()
        assert orc.rel_l2(out, want) <= REL_L2[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [0, 1, 4])
@pytest.mark.parametrize("weight", [10, 3, 0.5, 0.04, 0, -0.02, -0.5, -3, -10])
def test_forward_analytic(fasn_lib, n, weight, dtype):
    """The reference's closed-form GPU test (tests/gpu/core/test_flash_attn.py:51-91): constant inputs,
    L=1024, S=1152, E=Ev=64, scale 0.3; same tolerances (atol 1e-3; causal rtol 2e-3, bf16 2e-2)."""
    from flash_attention_softmax_n import flash_attention_n
    N, L, S, E, scale = 6, 1024, 1152, 64, 0.3
    q = torch.full((N, 1, L, E), weight, dtype=dtype, device="cuda")
    k = torch.full((N, 1, S, E), weight, dtype=dtype, device="cuda")
    v = torch.full((N, 1, S, E), weight, dtype=dtype, device="cuda")
    w = q[0, 0, 0, 0].item()                         # the value after rounding to the 16-bit type
    out = flash_attention_n(q, k, v, scale=scale, softmax_n_param=n)
    expect = orc.analytic_answer(N, L, S, E, E, scale, w, n)
    torch.testing.assert_close(out[:, 0].double().cpu(), expect, atol=1e-3 * max(1.0, abs(w)), rtol=4e-3 if dtype == torch.bfloat16 else 5e-4)
    outc = flash_attention_n(q, k, v, scale=scale, is_causal=True, softmax_n_param=n)
    expect_c = orc.analytic_causal_answer(N, L, S, E, E, scale, w, n)
    torch.testing.assert_close(outc.double().sum(dim=0).sum(dim=-1)[0].cpu(), expect_c, atol=1e-6,
                               rtol=2e-2 if dtype == torch.bfloat16 else 2e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_forward_mask_and_bias(fasn_lib, dtype):
    """attn_mask (bool, broadcast over heads) AND causal, plus a 3-D (H,L,S) bias: semantics of flash_attn.py:87-113."""
    from flash_attention_softmax_n import flash_attention_n
    B, H, L, S, D = 2, 3, 200, 264, 64
    q, k, v, _ = make_qkv(B, H, L, S, D, dtype, seed=5)
    g = torch.Generator().manual_seed(11)
    mask = (torch.rand(B, 1, L, S, generator=g) > 0.3)
    mask[..., 0] = True
    mask[0, 0, 7, :] = False                       # one fully masked row: result must be exactly 0
    bias = torch.randn(H, L, S, generator=g).to(dtype)
    kw = dict(softmax_n_param=2, scale=0.2, is_causal=True)
    out = flash_attention_n(q, k, v, attn_mask=mask.cuda(), attn_bias=bias.cuda(), **kw)
    want = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), attn_mask=mask,
                                attn_bias=bias.double(), **kw)
    native = orc.slow_attention_n(q, k, v, attn_mask=mask.cuda(), attn_bias=bias.cuda(), **kw)
    check_close("O(mask,bias)", out, want, native, dtype)
    assert out[0, :, 7].abs().max().item() == 0.0
    # key-padding style mask broadcast over heads and queries, no causal, n = 0
    pad = torch.ones(B, 1, 1, S, dtype=torch.bool)
    pad[0, ..., 200:] = False
    out2 = flash_attention_n(q, k, v, attn_mask=pad.cuda())
    want2 = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), attn_mask=pad)
    check_close("O(padding mask)", out2, want2, orc.slow_attention_n(q, k, v, attn_mask=pad.cuda()), dtype)


def test_forward_shared_kv_heads(fasn_lib):
    """3-D key/value = shared by all heads (flash_attn.py:75-79, intended meaning)."""
    from flash_attention_softmax_n import flash_attention_n
    dtype = torch.bfloat16
    q, k, v, _ = make_qkv(2, 4, 130, 190, 64, dtype, seed=9, heads_kv=1)
    out = flash_attention_n(q, k[:, 0], v[:, 0], softmax_n_param=1)
    want = orc.slow_attention_n(q.double().cpu(), k[:, 0].double().cpu(), v[:, 0].double().cpu(), softmax_n_param=1)
    check_close("O(shared kv)", out, want, orc.slow_attention_n(q, k[:, 0], v[:, 0], softmax_n_param=1), dtype)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("causal", [False, True])
def test_forward_dropout_matches_oracle_with_same_mask(fasn_lib, dtype, causal):
    from flash_attention_softmax_n import flash_attention_n
    B, H, L, S, D, p = 2, 2, 300, 260, 128, 0.1
    q, k, v, _ = make_qkv(B, H, L, S, D, dtype, seed=21)
    seed, offset = 0x5EED, 3
    out = flash_attention_n(q, k, v, softmax_n_param=0.5, dropout_p=p, is_causal=causal, _philox=(seed, offset))
    keep = orc.dropout_keep_mask(seed, offset, B, H, L, S, p)
    want = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), softmax_n_param=0.5,
                                is_causal=causal, keep_mask=keep, dropout_p=p)
    native = orc.slow_attention_n(q, k, v, softmax_n_param=0.5, is_causal=causal, keep_mask=keep.cuda(), dropout_p=p)
    check_close("O(dropout)", out, want, native, dtype, rel_scale=1.5)
    out_b = flash_attention_n(q, k, v, softmax_n_param=0.5, dropout_p=p, is_causal=causal, _philox=(seed, offset + 1))
    assert not torch.equal(out, out_b)


def test_forward_strided_inputs_and_errors(fasn_lib):
    from flash_attention_softmax_n import flash_attention_n
    dtype = torch.float16
    B, H, L, D = 2, 3, 160, 64
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(B, L, 3, H, D, generator=g) * 0.5).to(dtype).cuda()      # packed (B,L,3,H,D) projection output
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))              # (B,H,L,D) views, row stride 3*H*D
    out = flash_attention_n(q, k, v, softmax_n_param=1, is_causal=True)
    want = orc.slow_attention_n(q.double().cpu(), k.double().cpu(), v.double().cpu(), softmax_n_param=1, is_causal=True)
    check_close("O(strided)", out, want, orc.slow_attention_n(q, k, v, softmax_n_param=1, is_causal=True), dtype)
    with pytest.raises(NotImplementedError):
        flash_attention_n(q.double(), k.double(), v.double())                    # float64: use slow_attention_n
    with pytest.raises(NotImplementedError):
        flash_attention_n(q.float(), k, v)                                       # mixed dtypes
    with pytest.raises(NotImplementedError):
        big = torch.zeros(1, 1, 8, 160, dtype=dtype, device="cuda")
        flash_attention_n(big, big, big)                                         # head dims above 128


@pytest.mark.parametrize("E,Ev", [(32, 32), (16, 16), (64, 128), (128, 64), (80, 48), (8, 8)])
def test_other_head_dims_are_zero_padded(fasn_lib, E, Ev):
    """Head dims other than 64 / 128 (the Triton path's 16 / 32, flash_attn_triton.py:266) and Ev != E (README.md:50) run
    zero-padded to the next supported size; results and gradients equal the oracle's on the unpadded tensors."""
    from tests._util import run_fused, oracle_all, native_lowp_all
    dtype = torch.bfloat16
    B, H, L, S = 2, 2, 150, 210
    g = torch.Generator().manual_seed(E * 131 + Ev)
    q = (torch.randn(B, H, L, E, generator=g) * 0.5).to(dtype).cuda()
    k = (torch.randn(B, H, S, E, generator=g) * 0.5).to(dtype).cuda()
    v = (torch.randn(B, H, S, Ev, generator=g) * 0.5).to(dtype).cuda()
    do = torch.randn(B, H, L, Ev, generator=g).to(dtype).cuda()
    kw = dict(softmax_n_param=1.0, is_causal=True)
    got = run_fused(q, k, v, do, **kw)
    want = oracle_all(q, k, v, do, **kw)
    native = native_lowp_all(q, k, v, do, **kw)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        check_close(f"{name}(E={E},Ev={Ev})", a, b, nat, dtype, rel_scale=1.5)


@pytest.mark.parametrize("scale", [-0.2, 0.0])
def test_forward_backward_non_positive_scale(fasn_lib, scale):
    """A non-positive logit scale takes the generic (pre-scaling) path: the running max must follow the scaled scores."""
    from tests._util import run_fused, oracle_all, native_lowp_all
    dtype = torch.float16
    q, k, v, do = make_qkv(1, 2, 200, 264, 64, dtype, seed=41)
    kw = dict(softmax_n_param=1.0, scale=scale, is_causal=True)
    got = run_fused(q, k, v, do, **kw)
    want = oracle_all(q, k, v, do, **kw)
    native = native_lowp_all(q, k, v, do, **kw)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        check_close(f"{name}(scale={scale})", a, b, nat, dtype, rel_scale=1.5)


def test_forward_long_sequence_rows(fasn_lib):
    """BASELINE.json configs[4] shape for one head (S = 65536, D = 64, bf16, n = 1, causal): the first and the last 128
    query rows against the float64 oracle evaluated on just those rows (bottom-right alignment makes a row block of the
    full problem equal to a smaller causal problem with S_kv = last visible key + 1)."""
    from flash_attention_softmax_n import flash_attention_n
    dtype, S, D = torch.bfloat16, 65536, 64
    g = torch.Generator().manual_seed(65)
    q, k, v = ((torch.randn(1, 1, S, D, generator=g) * 0.5).to(dtype).cuda() for _ in range(3))
    out = flash_attention_n(q, k, v, softmax_n_param=1.0, is_causal=True)
    torch.cuda.synchronize()
    for lo, hi in ((0, 128), (S - 128, S), (30000, 30100)):
        want = orc.slow_attention_n(q[:, :, lo:hi].double().cpu(), k[:, :, :hi].double().cpu(), v[:, :, :hi].double().cpu(),
                                    softmax_n_param=1.0, is_causal=True)
        native = orc.slow_attention_n(q[:, :, lo:hi], k[:, :, :hi], v[:, :, :hi], softmax_n_param=1.0, is_causal=True)
        check_close(f"O[{lo}:{hi}]", out[:, :, lo:hi], want, native, dtype)


def test_cuda_graph_capture_and_replay(fasn_lib):
    """The kernels are capturable (no synchronisation, no allocation inside the C ABI once the per-device work counters exist): a
    captured forward + backward replays on new input values and matches the eager call bit for bit -- which also exercises the
    persistent kernels' work counters being handed back at zero by every launch (a replay reuses its counter pair).  With
    dropout the call refuses capture: (seed, offset) are host integers, a graph would replay one mask for ever."""
    from flash_attention_softmax_n import flash_attention_n
    dtype = torch.float16
    q, k, v, do = make_qkv(2, 3, 300, 300, 128, dtype, seed=123)
    sq, sk, sv = (t.clone().requires_grad_() for t in (q, k, v))
    flash_attention_n(sq, sk, sv, softmax_n_param=1.0, is_causal=True).backward(do)        # warm-up outside the capture
    torch.cuda.synchronize()
    sq.grad = sk.grad = sv.grad = None
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            so = flash_attention_n(sq, sk, sv, softmax_n_param=1.0, is_causal=True)
            so.backward(do)
    torch.cuda.current_stream().wait_stream(side)
    for rep in range(3):                      # new values in the static inputs, replay, compare with the eager call
        q2, k2, v2, _ = make_qkv(2, 3, 300, 300, 128, dtype, seed=200 + rep)
        with torch.no_grad():
            sq.copy_(q2); sk.copy_(k2); sv.copy_(v2)
        g.replay()
        torch.cuda.synchronize()
        eq, ek, ev = (t.clone().requires_grad_() for t in (q2, k2, v2))
        eo = flash_attention_n(eq, ek, ev, softmax_n_param=1.0, is_causal=True)
        eo.backward(do)
        torch.cuda.synchronize()
        assert torch.equal(so, eo) and torch.equal(sk.grad, ek.grad) and torch.equal(sv.grad, ev.grad)
        assert orc.rel_l2(sq.grad, eq.grad) <= 1e-3       # dQ is reduced with L2 atomics: the order of the fp32 adds is not fixed
    with pytest.raises(RuntimeError, match="CUDA graph"):
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g2, stream=side):
                flash_attention_n(q, k, v, softmax_n_param=1.0, dropout_p=0.1)
