import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-softmax-n_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def fasn_lib():
    """libfasn.so, built on demand (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, PKG)
    import build as fasn_build
    fasn_build.build()
    from flash_attention_softmax_n import _native
    return _native.load()


@pytest.fixture(scope="session")
def fasn_debug32_lib(fasn_lib):
    """libfasn_debug32.so: the forward with P as two 16-bit terms and float32 output (-DFASN_DEBUG_FP32_P=1)."""
    import ctypes
    import build as fasn_build
    from flash_attention_softmax_n import _native
    path = os.path.join(PKG, "flash_attention_softmax_n", "libfasn_debug32.so")
    src_time = max(os.path.getmtime(os.path.join(PKG, "csrc", f)) for f in os.listdir(os.path.join(PKG, "csrc")))
    if not os.path.exists(path) or os.path.getmtime(path) < src_time:
        fasn_build.build_variant("debug32", ["-DFASN_DEBUG_FP32_P=1"])
    lib = ctypes.CDLL(path)
    lib.fasn_fwd.argtypes = [ctypes.POINTER(_native.FasnParams)]
    lib.fasn_fwd.restype = ctypes.c_int
    lib.fasn_last_error.restype = ctypes.c_char_p
    return lib
