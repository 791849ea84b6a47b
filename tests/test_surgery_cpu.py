"""Host logic of the Hugging Face route (`flash_attention_softmax_n.surgery`): registration,
the eager softmax_n route against the library's own eager attention, mask decomposition.  No kernels run here."""
import pytest
import torch

transformers = pytest.importorskip("transformers")
from transformers import BertConfig, BertModel, LlamaConfig, LlamaModel, XLNetConfig, XLNetModel  # noqa: E402

from flash_attention_softmax_n.surgery import EAGER, FUSED, apply_attention_softmax_n  # noqa: E402
from flash_attention_softmax_n.surgery import attention_softmax_n as S  # noqa: E402


def _bert():
    torch.manual_seed(0)
    return BertModel(BertConfig(hidden_size=128, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256,
                                vocab_size=100, max_position_embeddings=64), add_pooling_layer=False).eval()


def _llama():
    torch.manual_seed(0)
    return LlamaModel(LlamaConfig(hidden_size=128, num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2,
                                  intermediate_size=256, vocab_size=100)).eval()


def _inputs():
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 100, (2, 16), generator=g)
    am = torch.ones(2, 16, dtype=torch.long)
    am[1, 10:] = 0
    return ids, am


@pytest.mark.parametrize("make", [_bert, _llama])
def test_eager_route_with_n0_is_the_models_own_attention(make):
    model = make()
    ids, am = _inputs()
    model.config._attn_implementation = "eager"
    with torch.no_grad():
        want = model(input_ids=ids, attention_mask=am).last_hidden_state
        count = apply_attention_softmax_n(model, 0.0, implementation=EAGER)
        assert count == 2 and model.config._attn_implementation == EAGER
        got = model(input_ids=ids, attention_mask=am).last_hidden_state
        valid = am.bool()
        assert torch.allclose(got[valid], want[valid], atol=5e-6)
        apply_attention_softmax_n(model, 1.0, implementation=EAGER)
        other = model(input_ids=ids, attention_mask=am).last_hidden_state
        assert (other[valid] - want[valid]).abs().max() > 1e-3          # n changes the result


def test_n_is_stamped_on_every_attention_module_and_validated():
    model = _bert()
    assert apply_attention_softmax_n(model, 0.5) == 2                  # default: the fused route
    assert model.config._attn_implementation == FUSED
    stamped = [m.softmax_n_param for m in model.modules() if hasattr(m, "softmax_n_param")]
    assert stamped == [0.5, 0.5]
    with pytest.raises(ValueError):
        apply_attention_softmax_n(model, -1.0)
    with pytest.raises(ValueError):
        apply_attention_softmax_n(model, 1.0, implementation="sdpa")
    assert apply_attention_softmax_n(torch.nn.Linear(4, 4), 1.0) == 0   # nothing to switch: warns, returns 0


def test_fused_route_has_no_cpu_fallback():
    model = _bert()
    apply_attention_softmax_n(model, 1.0)
    ids, am = _inputs()
    with pytest.raises((NotImplementedError, RuntimeError, ValueError, AssertionError, TypeError)):
        model(input_ids=ids, attention_mask=am)


def test_unstamped_module_is_an_error():
    S.register_attention_softmax_n()
    model = _bert()
    model.config._attn_implementation = EAGER
    ids, am = _inputs()
    with pytest.raises(RuntimeError, match="apply_attention_softmax_n"):
        model(input_ids=ids, attention_mask=am)


def test_mask_decomposition():
    B, L, Sk = 2, 8, 12
    keypad = torch.ones(B, 1, 1, Sk, dtype=torch.bool)
    keypad[1, ..., 9:] = False
    dense = keypad.expand(B, 1, L, Sk).contiguous()
    m, causal = S._decompose_mask(dense, L, Sk)
    assert m.shape == (B, 1, 1, Sk) and not causal and torch.equal(m, keypad)
    tri = torch.arange(Sk).view(1, Sk) <= torch.arange(L).view(L, 1) + (Sk - L)
    m, causal = S._decompose_mask((keypad & tri).contiguous(), L, Sk)
    assert causal and torch.equal(m & tri, keypad & tri)
    odd = dense.clone()
    odd[0, 0, 3, 2] = False                                              # neither pattern: stays dense
    m, causal = S._decompose_mask(odd, L, Sk)
    assert m is odd and not causal
    again, _ = S._decompose_mask(odd, L, Sk)                             # cached
    assert again is odd


def test_xlnet_core_keeps_its_scores_and_gets_softmax_n():
    """XLNet's relative-attention core (reference: surgery_functions/_xlnet.py:25-75) runs its own method body with the
    softmax inside it replaced; n = 0 reproduces the unmodified model."""
    torch.manual_seed(0)
    model = XLNetModel(XLNetConfig(d_model=128, n_head=2, n_layer=2, d_inner=256, vocab_size=100)).eval()
    ids, am = _inputs()
    with torch.no_grad():
        want = model(input_ids=ids, attention_mask=am.float()).last_hidden_state
        assert apply_attention_softmax_n(model, 0.0, implementation=EAGER) == 2
        got = model(input_ids=ids, attention_mask=am.float()).last_hidden_state
        assert torch.allclose(got, want, atol=5e-6)
        apply_attention_softmax_n(model, 1.0, implementation=EAGER)          # re-applying replaces, does not nest
        other = model(input_ids=ids, attention_mask=am.float()).last_hidden_state
        assert (other - want).abs().max() > 1e-3
        # the eager definition, applied by hand to one layer's scores, is what the patched core computes
        layer = model.layer[0].rel_attn
        assert layer.softmax_n_param == 1.0 and layer.rel_attn_core.__wrapped__ is type(layer).rel_attn_core
        apply_attention_softmax_n(model, 1.0)                                # fused route: CUDA only, no fallback
        with pytest.raises(NotImplementedError):
            model(input_ids=ids, attention_mask=am.float())


def test_softmax_mode_maps_every_softmax_spelling():
    from flash_attention_softmax_n import softmax_n
    x = torch.randn(3, 5, 7)
    with S._SoftmaxNMode(2.0, softmax_n):
        a = torch.nn.functional.softmax(x, dim=1)
        b = x.softmax(dim=-1)
        c = torch.softmax(x, 2)
        d = torch.nn.functional.softmax(x, dim=-1, dtype=torch.float64)
        e = x.exp()                                                         # other functions pass through
    assert torch.allclose(a, softmax_n(x, 2.0, dim=1)) and torch.allclose(b, softmax_n(x, 2.0, dim=-1))
    assert torch.allclose(c, b) and d.dtype == torch.float64 and torch.allclose(d.float(), b, atol=1e-6)
    assert torch.equal(e, torch.exp(x))
    assert torch.allclose(torch.softmax(x, 2), torch.nn.functional.softmax(x, dim=2))     # and nothing leaks out of the mode


def test_trainer_plugin_keeps_the_reference_shape():
    """`AttentionSoftmaxN(softmax_n_param)` (reference: surgery/attention_softmax_n.py:66-108): applies once, on the model it is
    handed; the Composer event plumbing needs composer itself."""
    from types import SimpleNamespace
    from flash_attention_softmax_n.surgery import AttentionSoftmaxN
    algo = AttentionSoftmaxN(softmax_n_param=1.0, implementation=EAGER)
    assert repr(algo) == "AttentionSoftmaxN()" and AttentionSoftmaxN.required_on_load() and not algo._applied
    model = _bert()
    algo.apply(None, SimpleNamespace(model=model, optimizers=None), None)
    assert algo._applied and model.config._attn_implementation == EAGER
    assert [m.softmax_n_param for m in model.modules() if hasattr(m, "softmax_n_param")] == [1.0, 1.0]
    if S._Event is None:
        with pytest.raises(RuntimeError, match="Composer"):
            algo.match(None, None)


def test_decompose_mask_under_inference_mode_and_without_version_counter():
    """Inference tensors carry no version counter (`mask._version` raises): the cache is then keyed by identity alone."""
    L = S_ = 6
    with torch.inference_mode():
        pad = torch.tensor([[1, 1, 1, 1, 0, 0]], dtype=torch.bool).view(1, 1, 1, S_).expand(1, 1, L, S_).contiguous()
        m, causal = S._decompose_mask(pad, L, S_)
        assert m.shape == (1, 1, 1, S_) and causal is False
        m2, causal2 = S._decompose_mask(pad, L, S_)                    # cache hit
        assert m2 is m and causal2 is False
        tri = torch.tril(torch.ones(L, S_, dtype=torch.bool)).view(1, 1, L, S_) & pad
        m3, causal3 = S._decompose_mask(tri, L, S_)
        assert m3.shape == (1, 1, 1, S_) and causal3 is True


@pytest.mark.parametrize("kw", [dict(softcap=30.0), dict(s_aux=torch.zeros(2)), dict(head_mask=torch.ones(2)), dict(output_attentions=True)])
def test_semantics_changing_kwargs_are_refused(kw):
    mod = torch.nn.Module()
    mod.softmax_n_param = 1.0
    q = torch.randn(1, 2, 4, 8)
    with pytest.raises(NotImplementedError, match="do not implement"):
        S.eager_attention_softmax_n_forward(mod, q, q, q, None, **kw)
    with pytest.raises(NotImplementedError, match="do not implement"):
        S.attention_softmax_n_forward(mod, q, q, q, None, **kw)
    # absent / None / False values pass through
    out, _ = S.eager_attention_softmax_n_forward(mod, q, q, q, None, softcap=None, output_attentions=False)
    assert out.shape == (1, 4, 2, 8)
