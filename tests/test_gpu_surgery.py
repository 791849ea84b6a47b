"""Hugging Face models routed to the fused kernels (`flash_attention_softmax_n.surgery`):
the fused route against the eager softmax_n route (the operator's definition, what the reference's patched forwards
compute, surgery_functions/_bert.py:73-111) on the same randomly initialised model, forward and parameter gradients.

Tolerance: with the fp32 eager route as the truth, the fp16 fused model must be no further from it than 2 x the
fp16 eager model is (+ 2e-3 floor on O(1) layer-normalised activations): both carry the same fp16 weight rounding, so
what is compared is the attention arithmetic."""
import copy

import pytest
import torch

transformers = pytest.importorskip("transformers")
from transformers import BertConfig, BertModel, LlamaConfig, LlamaModel, XLNetConfig, XLNetModel  # noqa: E402

from flash_attention_softmax_n.surgery import EAGER, FUSED, apply_attention_softmax_n  # noqa: E402
from flash_attention_softmax_n.surgery import attention_softmax_n as S  # noqa: E402

pytestmark = pytest.mark.gpu


def _bert():
    return BertModel(BertConfig(hidden_size=256, num_attention_heads=4, num_hidden_layers=2, intermediate_size=512,
                                vocab_size=1000, max_position_embeddings=512, hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.0), add_pooling_layer=False)


def _llama():
    return LlamaModel(LlamaConfig(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2,
                                  intermediate_size=512, vocab_size=1000, max_position_embeddings=512))


def _run(model, ids, am, w):
    model.zero_grad(set_to_none=True)
    out = model(input_ids=ids, attention_mask=am).last_hidden_state
    valid = am.bool()
    (out.float() * w)[valid].sum().backward()
    grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None and ("query" in n or "q_proj" in n)}
    return out.detach().float()[valid], grads


@pytest.mark.parametrize("make, n, padded", [(_bert, 1.0, True), (_bert, 0.0, False), (_llama, 0.5, True), (_llama, 1.0, False)])
def test_fused_route_matches_eager_softmax_n_route(make, n, padded):
    torch.manual_seed(0)
    truth = make().cuda().train()
    B, L = 3, 200
    ids = torch.randint(0, 1000, (B, L), device="cuda")
    am = torch.ones(B, L, dtype=torch.long, device="cuda")
    if padded:
        am[1, 150:] = 0
        am[2, 77:] = 0
    w = torch.randn(B, L, truth.config.hidden_size, device="cuda")
    eager16 = copy.deepcopy(truth).half()
    fused16 = copy.deepcopy(truth).half()
    apply_attention_softmax_n(truth, n, implementation=EAGER)
    apply_attention_softmax_n(eager16, n, implementation=EAGER)
    assert apply_attention_softmax_n(fused16, n) == 2 and fused16.config._attn_implementation == FUSED

    seen = []
    orig = S.flash_attention_n

    def spy(q, k, v, **kw):
        m = kw.get("attn_mask")
        seen.append((None if m is None else tuple(m.shape), kw["is_causal"], kw["softmax_n_param"]))
        return orig(q, k, v, **kw)

    S.flash_attention_n = spy
    try:
        got, got_g = _run(fused16, ids, am, w)
    finally:
        S.flash_attention_n = orig
    want, want_g = _run(truth, ids, am, w)
    base, base_g = _run(eager16, ids, am, w)

    # every layer reached the kernels with n, and with the O(B*S) mask description instead of the dense (B,1,L,L) mask
    assert len(seen) == 2 and all(s[2] == n for s in seen)
    causal_model = make is _llama
    for shape, causal, _ in seen:
        assert causal == causal_model
        assert shape == ((B, 1, 1, L) if padded else None)

    assert torch.isfinite(got).all()
    err, ref = (got - want).abs().max().item(), (base - want).abs().max().item()
    assert err <= 2 * ref + 2e-3, (err, ref)
    assert got_g.keys() == want_g.keys() and len(got_g) >= 2
    for name in want_g:
        scale = want_g[name].abs().max().item()
        e, r = (got_g[name] - want_g[name]).abs().max().item(), (base_g[name] - want_g[name]).abs().max().item()
        assert e <= 2 * r + 2e-3 * scale, (name, e, r, scale)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_xlnet_core_on_the_fused_row_kernel(dtype):
    """XLNet keeps its relative-position score arithmetic; its softmax runs on `softmax_n_fused`.  Against the eager
    softmax_n route on the same weights: fp32 to 1e-4 (relative to the largest value), fp16 within 2 x the eager fp16 model's
    own distance from the fp32 truth."""
    torch.manual_seed(0)
    truth = XLNetModel(XLNetConfig(d_model=256, n_head=4, n_layer=2, d_inner=512, vocab_size=1000, dropout=0.0)).cuda().train()
    B, L = 3, 160
    ids = torch.randint(0, 1000, (B, L), device="cuda")
    am = torch.ones(B, L, device="cuda")
    am[1, 100:] = 0
    w = torch.randn(B, L, 256, device="cuda")
    fused = copy.deepcopy(truth).to(dtype)
    eager = copy.deepcopy(truth).to(dtype)
    n = 1.0
    apply_attention_softmax_n(truth, n, implementation=EAGER)
    apply_attention_softmax_n(eager, n, implementation=EAGER)
    assert apply_attention_softmax_n(fused, n) == 2

    def run(model):
        model.zero_grad(set_to_none=True)
        out = model(input_ids=ids, attention_mask=am.to(next(model.parameters()).dtype)).last_hidden_state
        (out.float() * w).sum().backward()
        return out.detach().float(), model.layer[0].rel_attn.q.grad.detach().float()

    got, got_g = run(fused)
    want, want_g = run(truth)
    base, base_g = run(eager)
    assert torch.isfinite(got).all() and torch.isfinite(got_g).all()
    for a, b, c in ((got, want, base), (got_g, want_g, base_g)):
        scale = b.abs().max().item()
        if dtype == torch.float32:
            assert (a - b).abs().max().item() <= 1e-4 * scale + 1e-6
        else:
            assert (a - b).abs().max().item() <= 2 * (c - b).abs().max().item() + 2e-3 * scale
