"""Bring-up: one 128x128x128 tcgen05 MMA per operand form the attention kernels use (fasn_probe), and the
device dropout generator against its numpy restatement (bit-exact)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_probe_mma_forms(fasn_lib, mode, dtype):
    g = torch.Generator().manual_seed(100 + mode)
    x = torch.randn(128, 128, generator=g).to(dtype).cuda()
    y = torch.randn(128, 128, generator=g).to(dtype).cuda()
    c = torch.full((128, 128), float("nan"), dtype=torch.float32, device="cuda")
    rc = fasn_lib.fasn_probe(mode, 0 if dtype == torch.float16 else 1, x.data_ptr(), y.data_ptr(), c.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fasn_lib.fasn_last_error()
    torch.cuda.synchronize()
    xf, yf = x.double(), y.double()
    want = {0: xf @ yf.T, 1: xf @ yf, 2: xf.T @ yf, 3: xf @ yf}[mode]
    torch.testing.assert_close(c.double(), want, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", [4, 5, 6])
def test_probe_extension_k_step(fasn_lib, mode, dtype):
    """Q K^T plus one extra K-step whose no-swizzle operands carry a per-column constant (three-term 16-bit split), plain
    (4), with the second K-chunk aliasing the first (5), and with one core matrix serving every row group of A (6)."""
    g = torch.Generator().manual_seed(300 + mode)
    x = torch.randn(128, 128, generator=g).to(dtype).cuda()
    y = torch.randn(128, 128, generator=g).to(dtype).cuda()
    c = torch.full((128, 128), float("nan"), dtype=torch.float32, device="cuda")
    rc = fasn_lib.fasn_probe(mode, 0 if dtype == torch.float16 else 1, x.data_ptr(), y.data_ptr(), c.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fasn_lib.fasn_last_error()
    torch.cuda.synchronize()
    e = 3.25 * x[:, 0].double()
    want = x.double() @ y.double().T + (1.0 if mode == 4 else 2.0) * e[None, :]
    torch.testing.assert_close(c.double(), want, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", [10, 11, 12])
def test_probe_pair_mma_forms(fasn_lib, mode, dtype):
    """cta_group::2 operand forms of the paired backward kernel (cluster of two CTAs)."""
    g = torch.Generator().manual_seed(200 + mode)
    x = torch.randn(256, 128, generator=g).to(dtype).cuda()
    y = torch.randn(256 if mode == 11 else 128, 128, generator=g).to(dtype).cuda()
    shape = (128, 128) if mode == 11 else (256, 128)
    c = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
    rc = fasn_lib.fasn_probe(mode, 0 if dtype == torch.float16 else 1, x.data_ptr(), y.data_ptr(), c.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fasn_lib.fasn_last_error()
    torch.cuda.synchronize()
    xf, yf = x.double(), y.double()
    want = xf @ yf.T if mode == 10 else (xf.T @ yf if mode == 11 else xf @ yf)
    torch.testing.assert_close(c.double(), want, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("p,B,H,L,S,bh_offset", [(0.1, 2, 3, 70, 200, 0), (0.5, 1, 2, 129, 33, 5), (0.25, 1, 1, 16, 1024, 0)])
def test_dropout_mask_matches_numpy(fasn_lib, p, B, H, L, S, bh_offset):
    out = torch.zeros(B, H, L, S, dtype=torch.uint8, device="cuda")
    seed, offset = 0x1234_5678_9ABC_DEF0, 17
    rc = fasn_lib.fasn_dropout_mask(out.data_ptr(), B, H, L, S, ctypes.c_float(p), seed, offset, bh_offset,
                                    torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fasn_lib.fasn_last_error()
    torch.cuda.synchronize()
    want = orc.dropout_keep_mask(seed, offset, B, H, L, S, p, bh_offset=bh_offset)
    assert torch.equal(out.cpu().bool(), want)
