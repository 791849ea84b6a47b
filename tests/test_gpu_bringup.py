"""Bring-up: one 128x128x128 tcgen05 MMA per operand form the attention kernels use (fasn_probe), and the
device dropout generator against its numpy restatement (bit-exact)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
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
    want = {0: xf @ yf.T, 1: xf @ yf, 2: xf.T @ yf, 3: xf @ yf, 4: xf @ yf}[mode]   # 4: A staged in the backward's quad layout
    torch.testing.assert_close(c.double(), want, rtol=1e-4, atol=1e-3)


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
