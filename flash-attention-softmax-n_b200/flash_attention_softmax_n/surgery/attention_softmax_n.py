"""`apply_attention_softmax_n(model, softmax_n_param)`: make a Hugging Face model attend with softmax_n.

The reference does this with MosaicML Composer module surgery: it swaps the `forward` of `BertSelfAttention` /
`RobertaSelfAttention` / XLNet's `rel_attn_core` for copies in which `softmax` is replaced by the eager `softmax_n`
(surgery/attention_softmax_n.py:19-63, surgery_functions/_bert.py:14-121, _xlnet.py:11-75) -- the (B,H,L,S) score
matrix is still materialised and the fused attention path is never reached (README.md:225-235 sketches a policy that
would call `flash_attention_n`).

Here the model is routed to the fused kernels instead.  Current `transformers` attention modules all call one function
looked up by name (`ALL_ATTENTION_FUNCTIONS[config._attn_implementation]`), so no per-architecture forward copies are
needed: two attention functions are registered with that interface,

* ``"softmax_n_fused"``  -> `flash_attention_n` (sm_100a kernels; CUDA, fp16 / bf16; no fallback), and
* ``"softmax_n_eager"``  -> `slow_attention_n` (the operator's definition, any device / dtype; the checker),

both fed boolean masks (True = attend) by the library's `sdpa_mask` builder, and `apply_attention_softmax_n` stamps n on
every attention module of the model and switches its config to one of them.  Keeps the reference's signature
(`model, softmax_n_param, optimizers`); nothing is re-allocated, so `optimizers` has nothing to update.

Attention cores that cannot be expressed as fused attention -- XLNet's relative-position scores are a sum of three
query-dependent terms (`XLNetRelativeAttention.rel_attn_core`, the reference's surgery_functions/_xlnet.py:25-75) -- keep
their own score arithmetic and get the softmax inside them replaced: their core method runs under a torch-function mode
that turns `torch.nn.functional.softmax` / `Tensor.softmax` into the fused row kernel `softmax_n_fused` (one read and one
write per score, forward and backward; fp16 / bf16 / fp32 on CUDA) or, for ``"softmax_n_eager"``, the eager `softmax_n`.
No forward copies: the library's own method body is what runs.
"""
from __future__ import annotations

import logging
import weakref
from typing import Optional

import torch
from torch import Tensor
from torch.nn import Module

from torch.overrides import TorchFunctionMode

from flash_attention_softmax_n.core.flash_attn import flash_attention_n
from flash_attention_softmax_n.core.functional import slow_attention_n, softmax_n
from flash_attention_softmax_n.core.softmax import softmax_n_fused

log = logging.getLogger(__name__)

FUSED = "softmax_n_fused"
EAGER = "softmax_n_eager"
_ATTR = "softmax_n_param"

__all__ = ["FUSED", "EAGER", "AttentionSoftmaxN", "apply_attention_softmax_n", "attention_softmax_n_forward",
           "eager_attention_softmax_n_forward", "register_attention_softmax_n"]

try:                                             # MosaicML Composer is optional: only the trainer plug-in below needs it
    from composer.core import Algorithm as _AlgorithmBase, Event as _Event
except ImportError:                              # pragma: no cover - composer is not part of this image
    _AlgorithmBase, _Event = object, None


def _repeat_kv(x: Tensor, groups: int) -> Tensor:
    """(B, Hkv, S, D) -> (B, Hkv * groups, S, D) for grouped-query models (each K/V head serves `groups` query heads)."""
    if groups == 1:
        return x
    B, Hkv, S, D = x.shape
    return x[:, :, None].expand(B, Hkv, groups, S, D).reshape(B, Hkv * groups, S, D)


_MASK_CACHE: dict = {}      # one entry: every layer of a forward pass receives the same mask tensor object


def _decompose_mask(mask: Tensor, L: int, S: int):
    """Boolean (B, 1|H, L, S) mask -> (key-only mask (B, 1|H, 1, S) or the dense mask, causal flag).

    The library builds dense masks, but the two it builds in practice are "key padding" (all query rows equal) and
    "causal AND key padding".  Both have O(B*S) descriptions that keep the kernels on their fast paths (a dense mask
    costs B*H*L*S bytes of reads per pass): a key-only mask is passed with query stride 0, causality as the flag.
    The comparison reads the mask once and synchronises (once per forward pass: the result is cached for the other layers,
    keyed by the identity of the live tensor object -- a weak reference, so a recycled address can never hit -- and its
    version; inference tensors carry no version counter and are keyed by identity alone).  While a CUDA graph is being
    captured a synchronisation is illegal: the dense mask is then passed through unchanged (generic kernels)."""
    ver = None if mask.is_inference() else mask._version
    hit = _MASK_CACHE.get("k")
    if hit is not None and hit[0]() is mask and hit[1] == (ver, L, S):
        return hit[2], hit[3]
    if mask.is_cuda and torch.cuda.is_current_stream_capturing():
        return mask, False
    out = (mask, False)
    if L > 1 and mask.shape[2] == L:
        last = mask[:, :, -1:, :]
        if bool(torch.equal(mask, last.expand_as(mask))):
            out = (last, False)
        else:
            rows = torch.arange(L, device=mask.device).view(L, 1)
            cols = torch.arange(S, device=mask.device).view(1, S)
            tri = cols <= rows + (S - L)                       # bottom-right aligned, as is_causal (flash_attn.py:38-39)
            if bool(torch.equal(mask, last & tri)):
                out = (last, True)
    _MASK_CACHE["k"] = (weakref.ref(mask), (ver, L, S), out[0], out[1])
    return out


# Keyword arguments of the attention interface that change the attention arithmetic and that neither route implements.
# They are refused when present (not None / not False) instead of being swallowed: a model that needs them (Gemma-2's
# logit soft-capping, GPT-OSS attention sinks, head masks, returned attention weights) would otherwise run with silently
# different attention.
_UNSUPPORTED_KWARGS = ("softcap", "s_aux", "sinks", "head_mask", "sliding_window", "output_attentions")


def _check_kwargs(kwargs: dict) -> None:
    bad = [k for k in _UNSUPPORTED_KWARGS if kwargs.get(k) is not None and kwargs.get(k) is not False]
    if bad:
        raise NotImplementedError(f"softmax_n attention routes do not implement {bad}; this architecture needs an attention "
                                  "function of its own")


def _common(module: Module, query: Tensor, key: Tensor, value: Tensor, attention_mask: Optional[Tensor],
            is_causal: Optional[bool], decompose: bool):
    n = getattr(module, _ATTR, None)
    if n is None:
        raise RuntimeError(f"{type(module).__name__} has no `{_ATTR}`: call apply_attention_softmax_n(model, n) first")
    groups = getattr(module, "num_key_value_groups", 1)
    key, value = _repeat_kv(key, groups), _repeat_kv(value, groups)
    mask = bias = None
    if attention_mask is not None:
        if attention_mask.ndim != 4:
            raise ValueError(f"expected a 4-D attention mask, got {tuple(attention_mask.shape)}")
        if attention_mask.shape[-1] != key.shape[-2]:
            attention_mask = attention_mask[..., : key.shape[-2]]
        if attention_mask.dtype == torch.bool:
            mask = attention_mask
        else:                                   # additive float mask (the `eager` mask builder): added to the scaled scores
            bias = attention_mask
    causal = bool(getattr(module, "is_causal", False)) if is_causal is None else bool(is_causal)
    # as the library's own SDPA route: the flag only stands when no explicit mask is given and there is more than one query
    causal = causal and attention_mask is None and query.shape[2] > 1
    if mask is not None and decompose:
        mask, causal = _decompose_mask(mask, query.shape[2], key.shape[2])
    return float(n), key, value, mask, bias, causal


def attention_softmax_n_forward(module: Module, query: Tensor, key: Tensor, value: Tensor,
                                attention_mask: Optional[Tensor], dropout: float = 0.0,
                                scaling: Optional[float] = None, is_causal: Optional[bool] = None, **kwargs):
    """`transformers` attention-interface function on the fused kernels.  query/key/value are (B, H, L|S, D);
    returns ((B, L, H, D) output, None): attention weights are never materialised."""
    _check_kwargs(kwargs)
    n, key, value, mask, bias, causal = _common(module, query, key, value, attention_mask, is_causal, True)
    if not query.is_cuda or query.dtype not in (torch.float16, torch.bfloat16):
        raise NotImplementedError(
            f"the '{FUSED}' route runs on CUDA tensors in float16 / bfloat16 (got {query.device.type}, {query.dtype}): move the "
            f"model with .cuda().half() / .bfloat16() or run it under torch.autocast, or pass implementation='{EAGER}' to "
            "apply_attention_softmax_n for the eager definition -- there is no silent fallback")
    out = flash_attention_n(query, key, value, softmax_n_param=n, scale=scaling, dropout_p=float(dropout),
                            attn_mask=mask, attn_bias=bias, is_causal=causal)
    return out.transpose(1, 2).contiguous(), None


def eager_attention_softmax_n_forward(module: Module, query: Tensor, key: Tensor, value: Tensor,
                                      attention_mask: Optional[Tensor], dropout: float = 0.0,
                                      scaling: Optional[float] = None, is_causal: Optional[bool] = None, **kwargs):
    """The same interface on the operator's eager definition (`slow_attention_n`): what the reference's patched
    forwards compute (_bert.py:73-111), kept as the checker for the fused route and for CPU / fp32 models."""
    _check_kwargs(kwargs)
    n, key, value, mask, bias, causal = _common(module, query, key, value, attention_mask, is_causal, False)
    out = slow_attention_n(query, key, value, attn_mask=mask if mask is not None else bias, dropout_p=float(dropout),
                           is_causal=causal, scale=scaling, softmax_n_param=n, train=module.training)
    return out.transpose(1, 2).contiguous(), None


def register_attention_softmax_n() -> None:
    """Register both attention functions (and their boolean mask builder) with `transformers`.  Idempotent."""
    from transformers.masking_utils import ALL_MASK_ATTENTION_FUNCTIONS, sdpa_mask
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    for name, fn in ((FUSED, attention_softmax_n_forward), (EAGER, eager_attention_softmax_n_forward)):
        ALL_ATTENTION_FUNCTIONS.register(name, fn)
        ALL_MASK_ATTENTION_FUNCTIONS.register(name, sdpa_mask)


# method name -> classes (by name) whose attention core keeps its own score arithmetic and only has its softmax replaced
_SOFTMAX_CORES = {"XLNetRelativeAttention": "rel_attn_core"}
_SOFTMAX_FUNCS = (torch.nn.functional.softmax, torch.softmax, Tensor.softmax)


class _SoftmaxNMode(TorchFunctionMode):
    """Inside the mode every softmax call is softmax_n with the given n (other functions pass through untouched)."""

    def __init__(self, n: float, impl):
        super().__init__()
        self.n, self.impl = n, impl

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _SOFTMAX_FUNCS:
            x = args[0] if args else kwargs["input"]
            dim = args[1] if len(args) > 1 else kwargs.get("dim", -1)
            dtype = kwargs.get("dtype", args[3] if len(args) > 3 and func is torch.nn.functional.softmax else None)
            return self.impl(x, n=self.n, dim=dim, dtype=dtype)
        return func(*args, **kwargs)


def _wrap_softmax_core(module: Module, method: str, implementation: str) -> None:
    """Instance-level patch: `module.<method>` runs the class's own method under `_SoftmaxNMode`."""
    original = getattr(type(module), method)
    impl = softmax_n_fused if implementation == FUSED else softmax_n

    def core(*args, **kwargs):
        with _SoftmaxNMode(getattr(module, _ATTR), impl):
            return original(module, *args, **kwargs)

    core.__wrapped__ = original
    setattr(module, method, core)


def _is_attention_module(m: Module) -> bool:
    cfg = getattr(m, "config", None)
    return cfg is not None and hasattr(cfg, "_attn_implementation") and (
        hasattr(m, "is_causal") or hasattr(m, "num_key_value_groups") or hasattr(m, "scaling"))


def apply_attention_softmax_n(model: Module, softmax_n_param: float, optimizers=None, implementation: str = FUSED) -> int:
    """Make every attention module of a `transformers` model compute softmax_n attention (reference signature:
    surgery/attention_softmax_n.py:19-23).  Returns the number of attention modules switched; warns (as the reference
    does, :57-63) when there are none.

    :param model: a `transformers.PreTrainedModel` (or any module tree holding its attention modules).
    :param softmax_n_param: the value of n (any real >= 0).
    :param optimizers: accepted for signature compatibility; parameters are untouched, so there is nothing to update.
    :param implementation: ``"softmax_n_fused"`` (default) or ``"softmax_n_eager"``.
    """
    if implementation not in (FUSED, EAGER):
        raise ValueError(f"implementation must be {FUSED!r} or {EAGER!r}, got {implementation!r}")
    if softmax_n_param is None or float(softmax_n_param) < 0:
        raise ValueError("softmax_n_param must be a real number >= 0")
    register_attention_softmax_n()
    count = 0
    configs = {}
    for m in model.modules():
        core = _SOFTMAX_CORES.get(type(m).__name__)
        if core is not None and hasattr(m, core):
            setattr(m, _ATTR, float(softmax_n_param))
            _wrap_softmax_core(m, core, implementation)
            count += 1
        elif _is_attention_module(m) and not isinstance(m, type(model)):
            setattr(m, _ATTR, float(softmax_n_param))
            configs[id(m.config)] = m.config
            count += 1
    top = getattr(model, "config", None)
    if configs and top is not None and hasattr(top, "_attn_implementation"):
        configs[id(top)] = top
    if count == 0:
        log.warning("AttentionSoftmaxN had no effect on the model: no module that dispatches through the transformers "
                    "attention interface was found (supported: every architecture whose attention calls "
                    "ALL_ATTENTION_FUNCTIONS, e.g. BERT, RoBERTa, GPT-2, Llama).")
        return 0
    for cfg in configs.values():
        cfg._attn_implementation = implementation
    log.info("softmax_n attention (n = %s, %s) set on %d attention modules", softmax_n_param, implementation, count)
    return count


class AttentionSoftmaxN(_AlgorithmBase):
    """Composer trainer plug-in with the reference's name, constructor and behaviour (surgery/attention_softmax_n.py:66-108):
    at `Event.INIT` it calls `apply_attention_softmax_n(state.model, softmax_n_param, optimizers=state.optimizers)` once.
    `implementation` picks the fused (default) or the eager route.  Needs `mosaicml` (composer) to be used in a Trainer;
    without it the object can still be constructed and `apply_to(model)` used directly."""

    def __init__(self, softmax_n_param: float, implementation: str = FUSED) -> None:
        self.softmax_n_param = softmax_n_param
        self.implementation = implementation
        self._applied = False

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"

    @staticmethod
    def required_on_load() -> bool:
        return True

    def match(self, event, state) -> bool:
        del state
        if _Event is None:
            raise RuntimeError("AttentionSoftmaxN.match needs MosaicML Composer (pip install mosaicml); "
                               "call apply_to(model) or apply_attention_softmax_n(model, n) instead")
        return event == _Event.INIT and not self._applied

    def apply_to(self, model: Module, optimizers=None) -> int:
        count = apply_attention_softmax_n(model, self.softmax_n_param, optimizers=optimizers, implementation=self.implementation)
        self._applied = True
        return count

    def apply(self, event, state, logger) -> None:
        del event, logger
        self.apply_to(state.model, optimizers=state.optimizers)
