"""The oracle (oracle/attention_oracle.py) against everything that pins the reference's results for this path:
its known-answer rows, the closed-form constant-input answers, and golden vectors produced by running the
reference's own files (tests/golden/make_golden.py).  Also the package's eager functions against the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

T = torch.from_numpy


# ---- softmax_n known answers: tests/cpu/core/test_functional.py:10-37 of the reference
NUMERATORS = [[1, 3, 6], [3, 1, 4], [1 / 6, 1 / 3, 1 / 2], [0.5, 1.5, 3], [100, 200, 300],
              [1 / 600, 1 / 300, 1 / 200], [2 / 7, 4 / 7, 8 / 7]]


@pytest.mark.parametrize("n", [0.0, 1.0, 1e-3, 1e-6, 4.0])
def test_softmax_n_known_answers(n):
    num = torch.tensor(NUMERATORS, dtype=torch.float64)
    out = orc.softmax_n(torch.log(num), n)
    expect = num / (n + num.sum(-1, keepdim=True))
    torch.testing.assert_close(out, expect, rtol=1e-12, atol=0)
    big = orc.softmax_n(torch.tensor([12.0, 89.0, 710.0], dtype=torch.float64), n)      # overflow guard row
    assert big.sum().item() == pytest.approx(1.0)


def test_softmax_n_golden(golden):
    rows = T(golden["sm_rows"])
    for i in range(6):
        n = float(golden[f"sm_n{i}"])
        torch.testing.assert_close(orc.softmax_n(rows, n), T(golden[f"sm_out{i}"]), rtol=1e-13, atol=0)
    torch.testing.assert_close(orc.softmax_n(T(golden["sm_big"]), 1.0), T(golden["sm_big_out"]), rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", ["c1", "causal", "scale", "n0c"])
def test_slow_attention_golden(golden, name):
    g = {k: T(golden[f"slow_{name}_{k}"]).double() for k in ("q", "k", "v", "do", "o", "dq", "dk", "dv")}
    n, scale, causal = golden[f"slow_{name}_meta"]
    kw = dict(softmax_n_param=float(n), scale=None if scale < 0 else float(scale), is_causal=bool(causal))
    o, dq, dk, dv = orc.attention_fwd_bwd(g["q"], g["k"], g["v"], g["do"], **kw)
    # the reference results are stored as float32: agreement is limited by that rounding only
    for got, want in ((o, g["o"]), (dq, g["dq"]), (dk, g["dk"]), (dv, g["dv"])):
        torch.testing.assert_close(got, want, rtol=3e-7, atol=3e-7)


def test_slow_attention_float_mask_golden(golden):
    q, k, v = (T(golden[f"slowmask_{x}"]).double() for x in "qkv")
    mask = T(golden["slowmask_mask"]).double()
    out = orc.slow_attention_n(q, k, v, softmax_n_param=2.0, attn_bias=mask)
    torch.testing.assert_close(out, T(golden["slowmask_o"]).double(), rtol=3e-7, atol=3e-7)


def test_flash_attention_reference_route_golden(golden):
    """The reference's own `flash_attention_n` (zero-pad + SDPA on CPU, fp32) with bool mask AND causal + bias."""
    g = {k: T(golden[f"flash_{k}"]) for k in ("q", "k", "v", "do", "mask", "bias", "o", "dq", "dk", "dv")}
    o, dq, dk, dv = orc.attention_fwd_bwd(g["q"], g["k"], g["v"], g["do"], softmax_n_param=2.0, scale=0.2,
                                          attn_mask=g["mask"], attn_bias=g["bias"], is_causal=True)
    for got, want in ((o, g["o"]), (dq, g["dq"]), (dk, g["dk"]), (dv, g["dv"])):
        torch.testing.assert_close(got.float(), want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("n", [0.0, 1.0, 1e-3, 4.0, 0.5])
@pytest.mark.parametrize("w", [10.0, 1.0, 0.1, -0.1, -1.0, -10.0])
def test_analytic_answers(n, w):
    """tests/common.py:29-44 / tests/cpu/core/test_functional.py:126-149 of the reference."""
    N, L, S, E, Ev, scale = 2, 3, 4, 8, 7, 0.3
    q = torch.full((N, L, E), w, dtype=torch.float64)
    k = torch.full((N, S, E), w, dtype=torch.float64)
    v = torch.full((N, S, Ev), w, dtype=torch.float64)
    out = orc.slow_attention_n(q, k, v, scale=scale, softmax_n_param=n)
    torch.testing.assert_close(out, orc.analytic_answer(N, L, S, E, Ev, scale, w, n), rtol=1e-10, atol=1e-12)
    outc = orc.slow_attention_n(q, k, v, scale=scale, softmax_n_param=n, is_causal=True)
    torch.testing.assert_close(outc.sum(0).sum(-1), orc.analytic_causal_answer(N, L, S, E, Ev, scale, w, n),
                               rtol=1e-10, atol=1e-12)


def test_lse_and_fully_masked_rows():
    torch.manual_seed(0)
    q, k, v = torch.randn(1, 2, 5, 8, dtype=torch.float64), torch.randn(1, 2, 7, 8, dtype=torch.float64), torch.randn(1, 2, 7, 8, dtype=torch.float64)
    mask = torch.ones(1, 1, 5, 7, dtype=torch.bool)
    mask[..., 2, :] = False
    for n in (0.0, 1.5):
        out, lse = orc.slow_attention_n(q, k, v, softmax_n_param=n, attn_mask=mask, return_lse=True)
        assert torch.isfinite(out).all() and out[..., 2, :].abs().max() == 0
        s = orc.attention_scores(q, k, attn_mask=mask)
        want = torch.log(n + torch.exp(s).sum(-1))
        torch.testing.assert_close(lse[..., [0, 1, 3, 4]], want[..., [0, 1, 3, 4]])


def test_package_eager_functions_match_oracle():
    from flash_attention_softmax_n import softmax_n, slow_attention_n
    torch.manual_seed(1)
    x = torch.randn(4, 9, dtype=torch.float64)
    for n in (None, 0.0, 1.0, 0.5):
        torch.testing.assert_close(softmax_n(x, n), orc.softmax_n(x, n), rtol=1e-13, atol=0)
    q, k, v = torch.randn(2, 3, 6, 8, dtype=torch.float64), torch.randn(2, 3, 9, 8, dtype=torch.float64), torch.randn(2, 3, 9, 5, dtype=torch.float64)
    for causal in (False, True):
        a = slow_attention_n(q, k, v, is_causal=causal, softmax_n_param=1.0, scale=0.4)
        b = orc.slow_attention_n(q, k, v, is_causal=causal, softmax_n_param=1.0, scale=0.4)
        torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-14)
    m = torch.rand(2, 1, 6, 9) > 0.3
    m[..., 0] = True
    m0 = m.clone()
    a = slow_attention_n(q, k, v, attn_mask=m, softmax_n_param=2.0)
    assert torch.equal(m, m0)                      # the reference mutates a bool mask (functional.py:86); we must not
    torch.testing.assert_close(a, orc.slow_attention_n(q, k, v, attn_mask=m, softmax_n_param=2.0), rtol=1e-12, atol=1e-14)


# ---- dropout generator
def test_philox_known_answers():
    """Random123 known-answer vectors for Philox-4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(*[np.uint32(c) for c in ctr], *key)
        assert tuple(int(x) for x in got) == want


def test_dropout_mask_properties():
    p = 0.1
    m = orc.dropout_keep_mask(0x5EED, 3, 2, 3, 64, 200, p)
    assert m.shape == (2, 3, 64, 200) and m.dtype == torch.bool
    T8 = orc.keep_threshold(p)
    assert T8 == 230
    rate, n = m.float().mean().item(), m.numel()
    assert abs(rate - T8 / 256) < 4 * math.sqrt(0.1 * 0.9 / n)
    # pure function of the global unit index: a shard starting at unit 4 reproduces the tail of the full mask
    tail = orc.dropout_keep_mask(0x5EED, 3, 1, 2, 64, 200, p, bh_offset=4)
    assert torch.equal(tail.reshape(2, 64, 200), m.reshape(6, 64, 200)[4:])
    assert not torch.equal(m, orc.dropout_keep_mask(0x5EED, 4, 2, 3, 64, 200, p))
    assert orc.dropout_keep_mask(1, 0, 1, 1, 8, 40, 0.0).all()


def test_dropout_generator_statistics():
    """The kernels' generator (Philox-4x32-7 planes + bitwise threshold compare): keep rate T/256 and no visible correlation
    between neighbouring keys, neighbouring rows, neighbouring units or the two Philox calls of a word."""
    B, H, L, S, p = 2, 2, 256, 1024, 0.1
    keep = orc.dropout_keep_mask(0x5EED, 3, B, H, L, S, p).numpy().astype(np.float64)
    T = orc.keep_threshold(p)
    rate, n = keep.mean(), keep.size
    assert abs(rate - T / 256.0) < 5 * np.sqrt(rate * (1 - rate) / n)
    z = keep - rate
    var = z.var()
    for a, b in [(z[..., :-1], z[..., 1:]), (z[..., :-1, :], z[..., 1:, :]), (z[:, :-1], z[:, 1:]), (z[..., :-32], z[..., 32:]),
                 (z[..., :-16], z[..., 16:])]:
        corr = (a * b).mean() / var
        assert abs(corr) < 6.0 / np.sqrt(a.size), corr
    # different (seed, offset) streams are unrelated
    other = orc.dropout_keep_mask(0x5EED, 4, B, H, L, S, p).numpy().astype(np.float64) - rate
    assert abs((z * other).mean() / var) < 6.0 / np.sqrt(n)


@pytest.mark.parametrize("dtype, w2", [(torch.float16, 256.0), (torch.bfloat16, 1.0)])
def test_lse_fold_split_is_fp32_accurate(dtype, w2):
    """Arithmetic of the experimental LSE2 fold (csrc/fasn_bwd.cu, FOLD): -LSE2 / c enters the S^T MMA as
    2w * (hi + lo + lo2) with three 16-bit terms.  The reconstruction must leave the exponent c*S - LSE2 accurate to ~1e-5
    (a 1e-5 relative error in P), far inside the 16-bit rounding of P itself."""
    g = torch.Generator().manual_seed(5)
    c = (1.0 / 128 ** 0.5) * 1.4426950408889634
    lse2 = torch.empty(200000).uniform_(-8.0, 60.0, generator=g)
    x = lse2 * (-1.0 / (c * w2))
    h0 = x.to(dtype)
    r1 = x - h0.float()
    h1 = r1.to(dtype)
    h2 = (r1 - h1.float()).to(dtype)
    recon = w2 * (h0.float() + h1.float() + h2.float())            # what the tensor core adds to S (fp32 accumulation)
    assert (recon.double() * c + lse2.double()).abs().max().item() < 1e-5
