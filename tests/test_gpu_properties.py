"""Full-size checks (BASELINE.json configs[2]: B=4 H=32 S=4096 D=128 fp16, n=0.5, causal, dropout 0.1) through
size-independent properties, plus spot checks of single (batch, head) units against the oracle at full length."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc
from tests._util import check_close

pytestmark = pytest.mark.gpu

B, H, S, D, N_PARAM, P_DROP = 4, 32, 4096, 128, 0.5, 0.1
SEED, OFFSET = 0x5EED, 7


@pytest.fixture(scope="module")
def c3(fasn_lib):
    from flash_attention_softmax_n import flash_attention_n
    torch.manual_seed(1234)
    q, k, v = (torch.empty(B, H, S, D, device="cuda", dtype=torch.float16).normal_(0, 0.5).requires_grad_() for _ in range(3))
    do = torch.randn(B, H, S, D, device="cuda", dtype=torch.float16)
    o = flash_attention_n(q, k, v, softmax_n_param=N_PARAM, dropout_p=P_DROP, is_causal=True, _philox=(SEED, OFFSET))
    o.backward(do)
    torch.cuda.synchronize()
    return dict(q=q.detach(), k=k.detach(), v=v.detach(), do=do, o=o.detach(), dq=q.grad, dk=k.grad, dv=v.grad)


def test_full_size_finite_and_adjoint_identities(c3):
    """<O, dO> = <V, dV>  (O = A V, dV = A^T dO, dropout included) and <Q, dQ> = <K, dK> (both equal scale <dS, S>)."""
    for t in c3.values():
        assert torch.isfinite(t).all()
    f = lambda a, b: (a.double() * b.double()).sum().item()
    o_do, v_dv = f(c3["o"], c3["do"]), f(c3["v"], c3["dv"])
    q_dq, k_dk = f(c3["q"], c3["dq"]), f(c3["k"], c3["dk"])
    norm = lambda a, b: a.double().norm().item() * b.double().norm().item()
    assert abs(o_do - v_dv) <= 2e-3 * norm(c3["o"], c3["do"])
    assert abs(q_dq - k_dk) <= 2e-3 * norm(c3["q"], c3["dq"])


@pytest.mark.parametrize("b,h", [(0, 0), (2, 17), (3, 31)])
def test_full_size_units_match_oracle(c3, b, h):
    """One (batch, head) unit at S=4096 against the float64 oracle with the same keep mask (global unit index b*H+h)."""
    sl = lambda t: t[b:b + 1, h:h + 1].cpu()
    keep = orc.dropout_keep_mask(SEED, OFFSET, 1, 1, S, S, P_DROP, bh_offset=b * H + h)
    want = orc.attention_fwd_bwd(sl(c3["q"]), sl(c3["k"]), sl(c3["v"]), sl(c3["do"]), softmax_n_param=N_PARAM,
                                 is_causal=True, keep_mask=keep, dropout_p=P_DROP)
    dev = lambda t: t[b:b + 1, h:h + 1]
    native = orc.attention_fwd_bwd(dev(c3["q"]), dev(c3["k"]), dev(c3["v"]), dev(c3["do"]), dtype=torch.float16, softmax_n_param=N_PARAM,
                                   is_causal=True, keep_mask=keep.cuda(), dropout_p=P_DROP)
    for name, got, ref, nat in zip(("O", "dQ", "dK", "dV"), (c3["o"], c3["dq"], c3["dk"], c3["dv"]), want, native):
        check_close(f"{name}[{b},{h}]", sl(got), ref, nat, torch.float16, rel_scale=2.0)


def test_forward_is_deterministic_and_shard_invariant(c3):
    """Same inputs -> identical bits; a slab of units computed on its own with _bh_offset reproduces the full result
    (what batch x head sharding over GPUs relies on, dropout included)."""
    from flash_attention_softmax_n import flash_attention_n
    kw = dict(softmax_n_param=N_PARAM, dropout_p=P_DROP, is_causal=True, _philox=(SEED, OFFSET))
    again = flash_attention_n(c3["q"], c3["k"], c3["v"], **kw)
    assert torch.equal(again, c3["o"])
    lo, hi = 40, 72                                   # units 40..71 of the flattened (B*H) axis
    flat = lambda t: t.reshape(1, B * H, S, D)[:, lo:hi]
    part = flash_attention_n(flat(c3["q"]), flat(c3["k"]), flat(c3["v"]), _bh_offset=lo, **kw)
    assert torch.equal(part, flat(c3["o"]))


def test_linearity_in_value_and_dout(fasn_lib):
    from flash_attention_softmax_n import flash_attention_n
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    mk = lambda s: (torch.randn(2, 4, 1024, 128, generator=g) * s).to(dtype).cuda()
    q, k, v1, v2 = mk(0.5), mk(0.5), mk(0.5), mk(0.5)
    kw = dict(softmax_n_param=1.0, dropout_p=0.2, is_causal=True, _philox=(3, 9))
    o1, o2 = flash_attention_n(q, k, v1, **kw), flash_attention_n(q, k, v2, **kw)
    o12 = flash_attention_n(q, k, (2 * v1.float() - 0.5 * v2.float()).to(dtype), **kw)
    want = 2 * o1.float() - 0.5 * o2.float()
    assert orc.rel_l2(o12, want) < 8e-3               # three bf16 roundings


def test_host_buffer_entry_point(fasn_lib):
    """fasn_attention_host: HOST pointers in, HOST pointers out (the end-to-end C-ABI call)."""
    from flash_attention_softmax_n import flash_attention_n
    b, h, s, d = 2, 3, 384, 64
    g = torch.Generator().manual_seed(8)
    hq, hk, hv = ((torch.randn(b, h, s, d, generator=g) * 0.5).to(torch.float16) for _ in range(3))
    hdo = torch.randn(b, h, s, d, generator=g).to(torch.float16)
    outs = [torch.empty(b, h, s, d, dtype=torch.float16) for _ in range(4)]
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = fasn_lib.fasn_attention_host(0, b, h, s, s, d, ptr(hq), ptr(hk), ptr(hv), ptr(outs[0]), ptr(hdo), ptr(outs[1]),
                                      ptr(outs[2]), ptr(outs[3]), ctypes.c_float(1.0), ctypes.c_float(d ** -0.5), 1,
                                      ctypes.c_float(0.1), 11, 2, None)
    assert rc == 0, fasn_lib.fasn_last_error()
    q, k, v = (t.cuda().requires_grad_() for t in (hq, hk, hv))
    o = flash_attention_n(q, k, v, softmax_n_param=1.0, dropout_p=0.1, is_causal=True, _philox=(11, 2))
    o.backward(hdo.cuda())
    assert torch.equal(outs[0], o.detach().cpu())
    assert torch.equal(outs[2], k.grad.cpu()) and torch.equal(outs[3], v.grad.cpu())
    assert orc.rel_l2(outs[1], q.grad) < 1e-3         # dQ is reduced with fp32 adds whose order is not fixed


def test_attention_host_pipeline_matches_device_call(fasn_lib):
    """Host tensors in / out, chunks of units pipelined over streams: same bits as one device call (dQ: fp32 add order)."""
    from flash_attention_softmax_n import flash_attention_n
    from flash_attention_softmax_n.host import attention_host
    g = torch.Generator().manual_seed(2)
    b, h, l, s, d = 2, 5, 256, 384, 128
    hq = (torch.randn(b, h, l, d, generator=g) * 0.5).to(torch.bfloat16)
    hk, hv = ((torch.randn(b, h, s, d, generator=g) * 0.5).to(torch.bfloat16) for _ in range(2))
    hdo = torch.randn(b, h, l, d, generator=g).to(torch.bfloat16)
    kw = dict(softmax_n_param=0.5, is_causal=True, dropout_p=0.1, _philox=(21, 6))
    o, dq, dk, dv = attention_host(hq, hk, hv, hdo, chunks=4, **kw)
    assert not o.is_cuda and o.shape == (b, h, l, d)
    q, k, v = (t.cuda().requires_grad_() for t in (hq, hk, hv))
    ref = flash_attention_n(q, k, v, **kw)
    ref.backward(hdo.cuda())
    assert torch.equal(o, ref.detach().cpu()) and torch.equal(dk, k.grad.cpu()) and torch.equal(dv, v.grad.cpu())
    assert orc.rel_l2(dq, q.grad) < 2e-3
    o2 = attention_host(hq, hk, hv, chunks=3, **kw)          # forward only
    assert torch.equal(o2, o)
