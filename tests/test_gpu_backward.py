"""Backward parity (dQ, dK, dV through autograd -> fasn_bwd) against the float64 oracle."""
import pytest
import torch

from oracle import attention_oracle as orc
from tests._util import make_qkv, oracle_all, native_lowp_all, check_close, run_fused, REL_L2

pytestmark = pytest.mark.gpu

CASES = [
    # B  H  L     S     D    n     causal scale
    (1, 1, 128, 128, 64, 1.0, False, None),
    (2, 2, 256, 256, 128, 0.5, True, None),
    (1, 2, 384, 256, 64, 0.0, False, 0.1),
    (1, 2, 200, 333, 128, 4.0, False, 0.5),      # ragged
    (1, 2, 333, 200, 64, 1.0, True, None),       # S < L causal: some K/V tiles see few queries, some rows see no key
    (2, 1, 96, 160, 64, 0.5, True, None),        # S > L causal
    (1, 1, 640, 640, 128, 1e-3, True, None),
    (6, 1, 1024, 1024, 64, 1.0, True, 0.5),      # the reference's GPU test shape (tests/gpu/core/test_flash_attn.py:16-18)
    (5, 8, 256, 384, 128, 0.5, True, None),      # 40 units: both halves of the launch order (unit-major head, tile-major tail of 32)
    (7, 6, 200, 200, 64, 1.0, True, None),       # 42 units, ragged
    (1, 2, 256, 333, 128, 0.5, False, None),     # three K/V tiles, ragged tail
    (1, 2, 333, 200, 128, 1.0, True, None),      # S < L causal at head dim 128
    (2, 1, 96, 160, 128, 0.5, True, None),       # S > L causal
    (1, 2, 1024, 1024, 128, 0.5, True, None),    # 8 x 8 tiles, causal trip counts
    (1, 1, 512, 768, 128, 0.0, False, 0.2),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,S,D,n,causal,scale", CASES)
def test_backward_matches_oracle(fasn_lib, B, H, L, S, D, n, causal, scale, dtype):
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=3 * L + S)
    kw = dict(softmax_n_param=n, scale=scale, is_causal=causal)
    got = run_fused(q, k, v, do, **kw)
    want = oracle_all(q, k, v, do, **kw)
    native = native_lowp_all(q, k, v, do, **kw)
    for name, g, w, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        assert g.shape == w.shape and g.dtype == dtype
        check_close(name, g, w, nat, dtype, rel_scale=1.5)


@pytest.mark.parametrize("name", ["c1", "causal", "n0c"])
def test_backward_golden_vectors(fasn_lib, golden, name):
    n, scale, causal = golden[f"slow_{name}_meta"]
    dtype = torch.bfloat16
    q, k, v, do = (torch.from_numpy(golden[f"slow_{name}_{x}"]).to(dtype).cuda() for x in ("q", "k", "v", "do"))
    got = run_fused(q, k, v, do, softmax_n_param=float(n), scale=None if scale < 0 else float(scale), is_causal=bool(causal))
    for key, g in zip(("o", "dq", "dk", "dv"), got):
        want = torch.from_numpy(golden[f"slow_{name}_{key}"]).double()
        assert orc.rel_l2(g, want) <= 1.5 * REL_L2[dtype], key


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_dropout_same_mask(fasn_lib, dtype, causal):
    """The backward kernel regenerates the forward's keep mask from (seed, offset): gradients must equal the
    oracle's with that exact mask."""
    B, H, L, S, D, p = 1, 2, 300, 260, 128, 0.1
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=77)
    seed, offset = 0x5EED, 0
    keep = orc.dropout_keep_mask(seed, offset, B, H, L, S, p)
    kw = dict(softmax_n_param=0.5, is_causal=causal)
    got = run_fused(q, k, v, do, dropout_p=p, _philox=(seed, offset), **kw)
    want = oracle_all(q, k, v, do, keep_mask=keep, dropout_p=p, **kw)
    native = native_lowp_all(q, k, v, do, keep_mask=keep, dropout_p=p, **kw)
    for name, g, w, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        check_close(name + "(dropout)", g, w, nat, dtype, rel_scale=2.0)


def test_backward_mask_bias_and_shared_kv(fasn_lib):
    dtype = torch.bfloat16
    B, H, L, S, D = 2, 3, 136, 200, 64
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=13)
    g = torch.Generator().manual_seed(4)
    mask = torch.rand(B, 1, L, S, generator=g) > 0.3
    mask[..., 0] = True
    bias = torch.randn(H, L, S, generator=g).to(dtype)
    kw = dict(softmax_n_param=2, scale=0.2, is_causal=True)
    got = run_fused(q, k, v, do, attn_mask=mask.cuda(), attn_bias=bias.cuda(), **kw)
    want = oracle_all(q, k, v, do, attn_mask=mask, attn_bias=bias.double(), **kw)
    native = native_lowp_all(q, k, v, do, attn_mask=mask, attn_bias=bias, **kw)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        check_close(name + "(mask,bias)", a, b, nat, dtype, rel_scale=1.5)
    # shared K/V: gradients are summed over the heads
    q, k, v, do = make_qkv(2, 4, 130, 190, 64, dtype, seed=9, heads_kv=1)
    got = run_fused(q, k[:, 0], v[:, 0], do, softmax_n_param=1)
    want = oracle_all(q, k[:, 0], v[:, 0], do, softmax_n_param=1)
    native = native_lowp_all(q, k[:, 0], v[:, 0], do, softmax_n_param=1)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        assert a.shape == b.shape
        check_close(name + "(shared kv)", a, b, nat, dtype, rel_scale=2.0)


def test_triton_alias_forward_backward(fasn_lib):
    """flash_attention_n_triton(query, key, value, is_causal, scale, softmax_n_param): the reference's Triton test
    shape family (tests/gpu/core/test_flash_attn_triton.py:13-48), real-valued n, gradients included."""
    from flash_attention_softmax_n import flash_attention_n_triton
    dtype = torch.float16
    q, k, v, do = make_qkv(2, 4, 512, 512, 64, dtype, seed=31)
    for n, causal in [(1e-3, True), (3.0, False), (0.5, True)]:
        qq, kk, vv = (t.detach().clone().requires_grad_() for t in (q, k, v))
        o = flash_attention_n_triton(qq, kk, vv, causal, 0.2, n)
        o.backward(do)
        want = oracle_all(q, k, v, do, softmax_n_param=n, scale=0.2, is_causal=causal)
        native = native_lowp_all(q, k, v, do, softmax_n_param=n, scale=0.2, is_causal=causal)
        for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), (o, qq.grad, kk.grad, vv.grad), want, native):
            check_close(name + "(triton alias)", a, b, nat, dtype, rel_scale=1.5)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [0.0, 2.0])
def test_large_logits_exercise_rescaling(fasn_lib, dtype, n):
    """Inputs with std 2.5 give logits of std ~ 7 (D=128): the running maximum jumps by more than the lazy-rescale
    threshold many times per row, so the O-accumulator rescaling path in TMEM runs constantly; softmax is nearly
    one-hot, which also stresses the +n term (rows whose best logit is negative are dominated by n)."""
    B, H, L, S, D = 1, 2, 384, 512, 128
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=99, std=2.5)
    kw = dict(softmax_n_param=n, is_causal=True)
    got = run_fused(q, k, v, do, **kw)
    want = oracle_all(q, k, v, do, **kw)
    native = native_lowp_all(q, k, v, do, **kw)
    for name, g, w, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        assert torch.isfinite(g).all()
        rel = orc.rel_l2(g, w)
        nat_rel = orc.rel_l2(nat, w)
        # near-one-hot softmax amplifies the rounding of the 16-bit logits themselves: judge against the reference's own
        # low-precision path rather than a fixed envelope
        assert rel <= max(2.0 * nat_rel, 3.0 * REL_L2[dtype]), f"{name}: rel-L2 {rel:.3e} vs native {nat_rel:.3e}"


@pytest.mark.parametrize("D,causal,with_bias", [(128, True, False), (64, False, False), (128, False, True)])
def test_backward_key_padding_mask(fasn_lib, D, causal, with_bias):
    """(B,1,1,S) key-padding masks are broadcast over the query axis (row stride 0): the backward treats a padded key like
    a key beyond S (fast path) unless a dense bias forces the generic path."""
    dtype = torch.float16
    B, H, L, S = 3, 2, 200, 300
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=21)
    lens = torch.tensor([300, 170, 64])
    mask = (torch.arange(S)[None, :] < lens[:, None]).view(B, 1, 1, S)
    kw = dict(softmax_n_param=1.0, is_causal=causal)
    bias = None
    if with_bias:
        bias = torch.randn(H, L, S, generator=torch.Generator().manual_seed(5)).to(dtype)
    got = run_fused(q, k, v, do, attn_mask=mask.cuda(), attn_bias=None if bias is None else bias.cuda(), **kw)
    want = oracle_all(q, k, v, do, attn_mask=mask, attn_bias=None if bias is None else bias.double(), **kw)
    native = native_lowp_all(q, k, v, do, attn_mask=mask, attn_bias=bias, **kw)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        check_close(name + "(key padding)", a, b, nat, dtype, rel_scale=1.5)
    # gradients of padded keys are exactly zero
    for bi, n in enumerate(lens.tolist()):
        assert float(got[2][bi, :, n:].abs().max() if n < S else 0.0) == 0.0
        assert float(got[3][bi, :, n:].abs().max() if n < S else 0.0) == 0.0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,S,D,causal", [(2, 4, 256, 256, 128, True), (1, 3, 200, 333, 64, False), (1, 2, 333, 200, 128, True)])
def test_alibi_slopes_equal_the_dense_bias(fasn_lib, B, H, L, S, D, causal, dtype):
    """`_alibi_slopes` generates slopes[h] * (j - i - (S - L)) in the kernels: same results as the dense (H, L, S) attn_bias
    (flash_attn.py:62 names ALiBi as the use of attn_bias) in forward and backward."""
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=31)
    slopes = 2.0 ** (-8.0 * torch.arange(1, H + 1) / H)
    dist = (torch.arange(S)[None, :] - torch.arange(L)[:, None] - (S - L)).double()
    bias = slopes.double()[:, None, None] * dist[None]
    kw = dict(softmax_n_param=1.0, is_causal=causal)
    got = run_fused(q, k, v, do, _alibi_slopes=slopes.cuda(), **kw)
    want = oracle_all(q, k, v, do, attn_bias=bias, **kw)
    native = native_lowp_all(q, k, v, do, attn_bias=bias.to(dtype), **kw)     # the dense bias rounded to the I/O dtype, as a caller would pass it
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        check_close(name + "(alibi)", a, b, nat, dtype, rel_scale=1.5)
    with pytest.raises(ValueError):
        run_fused(q, k, v, do, _alibi_slopes=slopes.cuda(), attn_bias=bias.to(dtype).cuda(), **kw)
