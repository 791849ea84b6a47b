"""Shared helpers for the GPU parity tests: seeded inputs, the oracle comparison and the stated tolerances.

Tolerance: compare with the oracle evaluated in float64 on the up-cast inputs;
pass if rel-L2 <= 3e-3 (bf16) / 5e-4 (fp16) AND max-abs <= 2 x the error of the reference's own eager
`slow_attention_n` run natively in the low-precision dtype on the same inputs (+ a small absolute floor).
The reference's own envelopes (atol 1e-2 fp16 / 5e-2 bf16, tests/gpu/core/test_flash_attn.py:14) are looser
and are asserted as well."""
import torch

from oracle import attention_oracle as orc

REL_L2 = {torch.float16: 5e-4, torch.bfloat16: 3e-3}
REF_ATOL = {torch.float16: 1e-2, torch.bfloat16: 5e-2}
ABS_FLOOR = {torch.float16: 2e-4, torch.bfloat16: 2e-3}


def make_qkv(B, H, L, S, D, dtype, seed=0, std=0.5, device="cuda", heads_kv=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    Hk = H if heads_kv is None else heads_kv
    q = (torch.randn(B, H, L, D, generator=g) * std).to(dtype).to(device)
    k = (torch.randn(B, Hk, S, D, generator=g) * std).to(dtype).to(device)
    v = (torch.randn(B, Hk, S, D, generator=g) * std).to(dtype).to(device)
    do = torch.randn(B, H, L, D, generator=g).to(dtype).to(device)
    return q, k, v, do


def oracle_all(q, k, v, do, **kw):
    """float64 oracle (CPU): O, dQ, dK, dV."""
    cpu = lambda t: None if t is None else t.detach().cpu()
    kw = {a: (cpu(b) if torch.is_tensor(b) else b) for a, b in kw.items()}
    return orc.attention_fwd_bwd(cpu(q), cpu(k), cpu(v), cpu(do), dtype=torch.float64, **kw)


def native_lowp_all(q, k, v, do, **kw):
    """The same oracle evaluated natively in the inputs' low-precision dtype on the GPU: the error level the
    reference's own eager path has (used to scale the max-abs criterion)."""
    kw = {a: (b.to(q.device) if torch.is_tensor(b) else b) for a, b in kw.items()}
    return orc.attention_fwd_bwd(q, k, v, do, dtype=q.dtype, **kw)


def check_close(name, got, want64, native, dtype, rel_scale=1.0):
    got64 = got.detach().double().cpu()
    want64 = want64.detach().double().cpu()
    assert torch.isfinite(got64).all(), f"{name}: non-finite values"
    rel = orc.rel_l2(got64, want64)
    err = (got64 - want64).abs().max().item()
    nat = (native.detach().double().cpu() - want64).abs().max().item() if native is not None else float("inf")
    scale = max(1.0, want64.abs().max().item())
    assert rel <= REL_L2[dtype] * rel_scale, f"{name}: rel-L2 {rel:.3e} > {REL_L2[dtype] * rel_scale:.1e}"
    assert err <= 2.0 * nat + ABS_FLOOR[dtype] * scale, f"{name}: max-abs {err:.3e} vs native low-precision {nat:.3e}"
    assert err <= REF_ATOL[dtype] * scale, f"{name}: max-abs {err:.3e} exceeds the reference's own envelope"
    return rel, err


def run_fused(q, k, v, do, **kw):
    from flash_attention_softmax_n import flash_attention_n
    q = q.detach().clone().requires_grad_()
    k = k.detach().clone().requires_grad_()
    v = v.detach().clone().requires_grad_()
    o = flash_attention_n(q, k, v, **kw)
    o.backward(do)
    torch.cuda.synchronize()
    return o.detach(), q.grad, k.grad, v.grad
