"""Fused softmax_n (csrc/fasn_softmax.cu through fasn_softmax_n_fwd / fasn_softmax_n_bwd) against the float64 oracle
restatement of the reference's softmax_n (functional.py:15-29) and against the reference's known-answer rows
(tests/cpu/core/test_functional.py:15-37)."""
import pytest
import torch

from flash_attention_softmax_n import softmax_n_fused
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {torch.float32: (2e-6, 1e-6), torch.float16: (2e-3, 1e-3), torch.bfloat16: (1.6e-2, 8e-3)}   # rtol (rounding of the output), atol


def _ref(x, n, dim=-1):
    return orc.softmax_n(x.double().cpu(), n=n, dim=dim)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,n", [((7, 33), 1.0), ((3, 5, 128), 0.0), ((2, 4, 96, 200), 0.5), ((5, 1024), 4.0), ((9, 1032), 1.0),
                                     ((4, 2048), 1e-3), ((3, 4096), 2.0), ((2, 5000), 1.0), ((1, 40000), 0.25), ((300, 8), 1.0)])
def test_forward_matches_oracle(fasn_lib, dtype, shape, n):
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 3).to(dtype).cuda()
    y = softmax_n_fused(x, n)
    assert y.dtype == dtype and y.shape == x.shape
    rtol, atol = TOL[dtype]
    torch.testing.assert_close(y.double().cpu(), _ref(x, n), rtol=rtol, atol=atol)


def test_known_answer_rows(fasn_lib):
    """The reference's explicit rows (tests/cpu/core/test_functional.py:15-37): numerator / (n + sum), incl. the overflow row."""
    for row, n in [([1.0, 2.0, 3.0], 0.0), ([1.0, 2.0, 3.0], 1.0), ([-1.0, 0.0, 1.0], 1e-3), ([0.5, -0.5], 1e-6), ([1.0, 2.0, 3.0, 4.0], 4.0),
                   ([12.0, 89.0, 710.0], 1.0)]:
        x = torch.tensor([row], dtype=torch.float32, device="cuda")
        e = torch.exp(torch.tensor(row, dtype=torch.float64) - max(row))
        want = e / (n * torch.exp(torch.tensor(-max(row), dtype=torch.float64)) + e.sum())
        torch.testing.assert_close(softmax_n_fused(x, n)[0].double().cpu(), want, rtol=5e-6, atol=1e-7)


def test_very_negative_rows_and_masked_entries(fasn_lib):
    x = torch.full((2, 64), -200.0, device="cuda")            # the reference's exp(-max) overflows here (-> 0); exact value ~e^-200/n
    y = softmax_n_fused(x, 1.0)
    assert torch.isfinite(y).all() and float(y.abs().max()) < 1e-30
    x = torch.randn(4, 300, device="cuda")
    x[:, 100:] = float("-inf")
    x[3] = float("-inf")                                       # a row with no finite entry: 0 (n > 0: exact; n = 0: defined)
    for n in (0.0, 1.0):
        y = softmax_n_fused(x, n)
        assert torch.isfinite(y).all()
        assert float(y[:, 100:].abs().max()) == 0.0 and float(y[3].abs().max()) == 0.0
        torch.testing.assert_close(y[:3, :100].double().cpu(), _ref(x[:3, :100], n), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("dtype,out", [(torch.float16, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, None)])
def test_dtype_argument_and_other_axis(fasn_lib, dtype, out):
    x = (torch.randn(6, 130, 12, generator=torch.Generator().manual_seed(3)) * 2).to(dtype).cuda()
    y = softmax_n_fused(x, 1.0, dim=1, dtype=out)
    assert y.dtype == (dtype if out is None else out)
    rtol, atol = TOL[y.dtype]
    torch.testing.assert_close(y.double().cpu(), _ref(x, 1.0, dim=1), rtol=max(rtol, 2e-3 if dtype != torch.float32 else 0), atol=atol)
    xs = x[:, ::2, :]                                          # non-contiguous input
    rtol, atol = TOL[dtype]
    torch.testing.assert_close(softmax_n_fused(xs, 0.5).double().cpu(), _ref(xs, 0.5), rtol=rtol, atol=atol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(5, 40), (3, 7, 1000), (2, 3000), (2, 6000)])
def test_backward_matches_autograd_of_the_definition(fasn_lib, dtype, shape):
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(*shape, generator=g) * 2).to(dtype).cuda().requires_grad_()
    dy = torch.randn(*shape, generator=g).to(dtype).cuda()
    y = softmax_n_fused(x, 1.5)
    y.backward(dy)
    xr = x.detach().double().cpu().requires_grad_()
    yr = orc.softmax_n(xr, n=1.5, dim=-1)
    yr.backward(dy.double().cpu())
    assert x.grad.dtype == dtype
    rel = float((x.grad.double().cpu() - xr.grad).norm() / xr.grad.norm())
    assert rel < {torch.float32: 1e-5, torch.float16: 2e-3, torch.bfloat16: 1.5e-2}[dtype], rel


def test_argument_errors(fasn_lib):
    with pytest.raises(NotImplementedError):
        softmax_n_fused(torch.randn(3, 4), 1.0)                # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        softmax_n_fused(torch.randn(3, 4, device="cuda", dtype=torch.float64), 1.0)
    with pytest.raises(ValueError):
        softmax_n_fused(torch.randn(3, 4, device="cuda"), -1.0)
