"""Workload for compute-sanitizer (scripts/gpu_sanitize.sh): the seeded random sweep of tests/test_gpu_random_shapes.py (both head
dims, ragged L / S, causal, dropout, key-padding masks, shared K/V, fp16 / bf16) plus dense mask / bias (with its gradient), ALiBi
slopes, float32 inputs, a launch with more work items than SMs (so persistent CTAs take several) and the fused softmax_n rows --
forward + backward through the public API, each checked for finite results.  No pytest: the sanitizer attaches to this process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
import torch
from flash_attention_softmax_n import flash_attention_n, softmax_n_fused
from tests.test_gpu_random_shapes import _cases
from tests._util import make_qkv

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 28


def run(q, k, v, do, **kw):
    q, k, v = (t.detach().clone().requires_grad_() for t in (q, k, v))
    o = flash_attention_n(q, k, v, **kw)
    o.backward(do)
    torch.cuda.synchronize()
    for t in (o, q.grad, k.grad, v.grad):
        assert torch.isfinite(t).all()


done = 0
for case in _cases(n_cases, 2026):
    i, B, H, L, S, D, causal, n, scale, p, pad, shared, dtype = case
    q, k, v, do = make_qkv(B, H, L, S, D, dtype, seed=1000 + i, heads_kv=1 if shared else None)
    kw = dict(softmax_n_param=n, scale=scale, is_causal=causal)
    if pad:
        lens = torch.randint(1, S + 1, (B,), generator=torch.Generator().manual_seed(i))
        kw["attn_mask"] = (torch.arange(S)[None, :] < lens[:, None]).view(B, 1, 1, S).cuda()
    if p > 0:
        kw.update(dropout_p=p, _philox=(77 + i, 3 * i))
    run(q, k[:, 0] if shared else k, v[:, 0] if shared else v, do, **kw)
    done += 1
# dense mask AND causal + dense bias with gradient, dropout (the AUX backward instantiation writes dS)
q, k, v, do = make_qkv(2, 3, 136, 200, 128, torch.float16, seed=7)
g = torch.Generator().manual_seed(3)
mask = (torch.rand(2, 1, 136, 200, generator=g) > 0.2).cuda()
bias = (torch.randn(3, 136, 200, generator=g) * 0.3).half().cuda().requires_grad_()
run(q, k, v, do, softmax_n_param=1.0, is_causal=True, attn_mask=mask, attn_bias=bias, dropout_p=0.25, _philox=(5, 1))
assert torch.isfinite(bias.grad).all()
# ALiBi slopes, float32 inputs, many work items per persistent CTA
q, k, v, do = make_qkv(1, 4, 300, 300, 64, torch.bfloat16, seed=8)
run(q, k, v, do, softmax_n_param=0.5, is_causal=True, _alibi_slopes=torch.tensor([0.5, 0.25, 0.125, 0.0625]))
q, k, v, do = make_qkv(2, 2, 257, 257, 64, torch.float32, seed=9)
run(q, k, v, do, softmax_n_param=1.0, is_causal=True)
q, k, v, do = make_qkv(8, 40, 384, 384, 128, torch.float16, seed=10)     # 960 backward items, 640 forward items on 148 SMs
run(q, k, v, do, softmax_n_param=0.5, is_causal=True, dropout_p=0.1, _philox=(1, 2))
# fused softmax_n rows
for shape, dt in (((7, 5, 333), torch.float16), ((64, 4096), torch.bfloat16), ((3, 70000), torch.float32)):
    x = torch.randn(*shape, device="cuda", dtype=dt).requires_grad_()
    y = softmax_n_fused(x, 1.0)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
print(f"sanitize_cases OK: {done} random cases + 4 feature cases + 3 softmax_n cases")
