"""Seeded random sweep over shapes and options (ragged L and S, both head dims, real n, custom scales, causal with S != L,
dropout, key-padding masks, shared K/V), forward + backward against the float64 oracle.
Complements the hand-picked cases of test_gpu_forward.py / test_gpu_backward.py."""
import random

import pytest
import torch

from oracle import attention_oracle as orc
from tests._util import make_qkv, oracle_all, native_lowp_all, check_close, run_fused

pytestmark = pytest.mark.gpu


def _cases(count, seed):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        D = rng.choice([64, 128])
        B, H = rng.randint(1, 3), rng.randint(1, 4)
        L, S = rng.randint(1, 520), rng.randint(1, 520)
        causal = rng.random() < 0.5
        n = rng.choice([0.0, 0.5, 1.0, 3.0, 1e-3])
        scale = rng.choice([None, None, 0.05, 0.3])
        p = rng.choice([0.0, 0.0, 0.1, 0.35])
        pad = rng.random() < 0.3
        shared = rng.random() < 0.2
        dtype = rng.choice([torch.float16, torch.bfloat16])
        out.append((i, B, H, L, S, D, causal, n, scale, p, pad, shared, dtype))
    return out


@pytest.mark.parametrize("case", _cases(28, 2026), ids=lambda c: "-".join(str(x) for x in c[:6]))
def test_random_case(fasn_lib, case):
    i, B, H, L, S, D, causal, n, scale, p, pad, shared, dtype = case
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=1000 + i, heads_kv=1 if shared else None)
    kw = dict(softmax_n_param=n, scale=scale, is_causal=causal)
    okw = dict(kw)
    if pad:
        lens = torch.randint(1, S + 1, (B,), generator=torch.Generator().manual_seed(i))
        mask = (torch.arange(S)[None, :] < lens[:, None]).view(B, 1, 1, S)
        kw["attn_mask"], okw["attn_mask"] = mask.cuda(), mask
    if p > 0:
        seed, offset = 77 + i, 3 * i
        kw.update(dropout_p=p, _philox=(seed, offset))
        okw.update(dropout_p=p, keep_mask=orc.dropout_keep_mask(seed, offset, B, H, L, S, p))
    kk, vv = (k[:, 0], v[:, 0]) if shared else (k, v)
    got = run_fused(q, kk, vv, do, **kw)
    want = oracle_all(q, kk, vv, do, **okw)
    native = native_lowp_all(q, kk, vv, do, **okw)
    for name, a, b, nat in zip(("O", "dQ", "dK", "dV"), got, want, native):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        check_close(f"{name}[case {i}]", a, b, nat, dtype, rel_scale=2.0)
