"""Generate golden vectors by running the REFERENCE's own files (build container only).

    python tests/golden/make_golden.py          # writes tests/golden/golden_v1.npz

The reference package cannot be imported normally here (its Triton module allocates on CUDA at import,
flash_attn_triton.py:238), so the two hot-path files are loaded by path, as SURVEY.md section 8(c) describes.
The .npz travels with the repo; nothing at test time reads /root/reference.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("FASN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    warnings.simplefilter("ignore")
    ref_fn = _load("ref_functional", "flash_attention_softmax_n/core/functional.py")
    ref_fa = _load("ref_flash_attn", "flash_attention_softmax_n/core/flash_attn.py")
    out = {}
    g = torch.Generator().manual_seed(20231017)

    def randn(*shape, std=0.5, dtype=torch.float64):
        # values are rounded to the bf16 grid so the same fixture feeds fp16/bf16 kernels exactly
        x = torch.randn(*shape, generator=g, dtype=torch.float32) * std
        return x.to(torch.bfloat16).to(dtype)

    # (1) softmax_n rows -- the reference's KAT inputs (tests/cpu/core/test_functional.py:15-23) + overflow row
    rows = torch.log(torch.tensor([[1, 3, 6], [3, 1, 4], [1 / 6, 1 / 3, 1 / 2], [0.5, 1.5, 3],
                                   [100, 200, 300], [1 / 600, 1 / 300, 1 / 200], [2 / 7, 4 / 7, 8 / 7]],
                                  dtype=torch.float64))
    out["sm_rows"] = rows.numpy()
    for i, n in enumerate([0.0, 1.0, 1e-3, 1e-6, 4.0, 0.5]):
        out[f"sm_n{i}"] = np.float64(n)
        out[f"sm_out{i}"] = ref_fn.softmax_n(rows, n=n, dim=-1).numpy()
    big = torch.tensor([12.0, 89.0, 710.0], dtype=torch.float64)
    out["sm_big"] = big.numpy()
    out["sm_big_out"] = ref_fn.softmax_n(big, 1.0, dim=-1).numpy()

    # (2) slow_attention_n, forward + gradients, float64
    cases = [
        # name,      B  H  L    S    E   Ev  n     scale  causal
        ("c1",       1, 1, 128, 128, 64, 64, 1.0,  None,  False),   # BASELINE.json configs[0]
        ("causal",   1, 2, 96,  160, 64, 64, 0.5,  None,  True),    # Sq != Skv, bottom-right aligned
        ("scale",    2, 1, 80,  72,  32, 48, 4.0,  0.3,   False),   # Ev != E, custom scale
        ("n0c",      1, 1, 130, 130, 128, 128, 0.0, None, True),
    ]
    for name, B, H, L, S, E, Ev, n, scale, causal in cases:
        q = randn(B, H, L, E).requires_grad_()
        k = randn(B, H, S, E).requires_grad_()
        v = randn(B, H, S, Ev).requires_grad_()
        do = randn(B, H, L, Ev, std=1.0)
        o = ref_fn.slow_attention_n(q, k, v, softmax_n_param=n, scale=scale, is_causal=causal)
        o.backward(do)
        for key, val in dict(q=q, k=k, v=v, do=do, o=o, dq=q.grad, dk=k.grad, dv=v.grad).items():
            out[f"slow_{name}_{key}"] = val.detach().numpy().astype(np.float32)
        out[f"slow_{name}_meta"] = np.array([n, -1.0 if scale is None else scale, float(causal)])

    # 2-D float mask through the slow path (functional.py:87-88)
    q, k, v = randn(2, 2, 40, 32), randn(2, 2, 56, 32), randn(2, 2, 56, 32)
    fmask = randn(40, 56, std=1.0)
    fmask[torch.rand(40, 56, generator=g) < 0.2] = float("-inf")
    fmask[:, 0] = 0.0
    f32 = lambda t: t.numpy().astype(np.float32)
    out["slowmask_q"], out["slowmask_k"], out["slowmask_v"] = f32(q), f32(k), f32(v)
    out["slowmask_mask"] = f32(fmask)
    out["slowmask_o"] = f32(ref_fn.slow_attention_n(q, k, v, attn_mask=fmask.clone(), softmax_n_param=2.0))

    # (3) flash_attention_n (reference SDPA route on CPU): bool mask AND causal + 3-D bias, integer n
    B, H, L, S, E = 2, 3, 48, 64, 32
    q = randn(B, H, L, E, dtype=torch.float32).requires_grad_()
    k = randn(B, H, S, E, dtype=torch.float32).requires_grad_()
    v = randn(B, H, S, E, dtype=torch.float32).requires_grad_()
    do = randn(B, H, L, E, std=1.0, dtype=torch.float32)
    mask = torch.rand(B, 1, L, S, generator=g) > 0.25
    mask[..., 0] = True
    bias = randn(H, L, S, std=1.0, dtype=torch.float32)
    o = ref_fa.flash_attention_n(q, k, v, softmax_n_param=2, scale=0.2, attn_mask=mask, attn_bias=bias,
                                 is_causal=True)
    o.backward(do)
    for key, val in dict(q=q, k=k, v=v, do=do, mask=mask, bias=bias, o=o, dq=q.grad, dk=k.grad,
                         dv=v.grad).items():
        out[f"flash_{key}"] = val.detach().numpy()

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
