"""Host-side checks that need no GPU: the import surface mirrors the reference's, the C-ABI library loads and
exports every symbol include/fasn.h declares, argument errors are reported (not crashes), and the fused
entry point refuses inputs it does not cover instead of falling back."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_import_surface_matches_reference():
    import flash_attention_softmax_n as pkg
    from flash_attention_softmax_n.core.flash_attn import flash_attention_n
    from flash_attention_softmax_n.core.functional import softmax_n, slow_attention_n, DType  # noqa: F401
    from flash_attention_softmax_n.core.flash_attn_triton import flash_attention_n_triton
    assert pkg.flash_attention_n is flash_attention_n and pkg.flash_attention_n_triton is flash_attention_n_triton
    assert pkg.TRITON_INSTALLED is True
    # reference signature: flash_attn.py:42-52 (names, order, defaults)
    sig = inspect.signature(flash_attention_n)
    names = [p for p in sig.parameters if not p.startswith("_")]
    assert names == ["query", "key", "value", "softmax_n_param", "scale", "dropout_p", "attn_mask", "attn_bias", "is_causal"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["softmax_n_param"], d["scale"], d["dropout_p"], d["attn_mask"], d["attn_bias"], d["is_causal"]) == \
           (None, None, 0.0, None, None, False)
    # flash_attn_triton.py:339-345
    tsig = inspect.signature(flash_attention_n_triton)
    assert list(tsig.parameters) == ["query", "key", "value", "is_causal", "scale", "softmax_n_param"]
    ssig = inspect.signature(slow_attention_n)
    assert list(ssig.parameters) == ["query", "key", "value", "attn_mask", "dropout_p", "is_causal", "scale",
                                     "softmax_n_param", "softmax_dtype", "train"]


def test_library_exports_every_declared_symbol(fasn_lib):
    header = open(os.path.join(ROOT, "include", "fasn.h")).read()
    declared = set(re.findall(r"\b(fasn_[a-z_]+)\s*\(", header))
    assert {"fasn_fwd", "fasn_bwd", "fasn_version", "fasn_last_error"} <= declared
    for name in declared:
        assert hasattr(fasn_lib, name), f"{name} declared in include/fasn.h but not exported"
    from flash_attention_softmax_n import _native
    assert declared == set(_native.EXPORTS)
    assert fasn_lib.fasn_version() == _native.FASN_ABI_VERSION


def test_argument_errors_are_reported_without_touching_the_gpu(fasn_lib):
    from flash_attention_softmax_n import _native
    assert fasn_lib.fasn_fwd(None) == -1 and b"null" in fasn_lib.fasn_last_error()
    p = _native.FasnParams()
    p.struct_size = 7
    assert fasn_lib.fasn_fwd(ctypes.byref(p)) == -1 and b"struct_size" in fasn_lib.fasn_last_error()
    # a correct struct_size gets past the size check: proves ctypes' layout has the C struct's size
    p.struct_size = ctypes.sizeof(_native.FasnParams)
    p.dtype = 9
    assert fasn_lib.fasn_fwd(ctypes.byref(p)) == -2 and b"dtype" in fasn_lib.fasn_last_error()
    p.dtype, p.head_dim = 0, 96
    assert fasn_lib.fasn_bwd(ctypes.byref(p)) == -2 and b"head_dim" in fasn_lib.fasn_last_error()
    p.head_dim, p.batch, p.heads, p.heads_kv, p.seqlen_q, p.seqlen_kv = 64, 1, 2, 2, 8, 8
    p.softmax_n = -1.0
    assert fasn_lib.fasn_fwd(ctypes.byref(p)) == -1 and b"softmax_n" in fasn_lib.fasn_last_error()
    d, a = ctypes.c_uint64(), ctypes.c_uint64()
    p.seqlen_q, p.head_dim = 130, 128
    assert fasn_lib.fasn_bwd_workspace(ctypes.byref(p), ctypes.byref(d), ctypes.byref(a)) == 0
    assert d.value == 2 * 1 * 2 * 256 * 4 and a.value == 1 * 2 * 256 * 128 * 4


def test_no_cpu_fallback():
    from flash_attention_softmax_n import flash_attention_n
    q = torch.randn(1, 2, 8, 64)
    with pytest.raises(NotImplementedError):
        flash_attention_n(q, q, q)                       # CPU tensors: the reference would run SDPA; we refuse
    with pytest.raises(ValueError):
        flash_attention_n(q[0], q[0], q[0])              # query must be 4-D (flash_attn.py:85)
    with pytest.raises(ValueError):
        flash_attention_n(q, q, q, softmax_n_param=-1)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from flash_attention_softmax_n import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setenv("FASN_LIBRARY", str(tmp_path / "nope.so"))
    with pytest.raises(_native.FasnError, match="no CPU or PyTorch fallback"):
        _native.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flash-attention-softmax-n_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "attention_oracle" not in src.replace(
                    "oracle/attention_oracle.py", ""), f
