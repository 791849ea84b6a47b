"""First GPU check of the experimental LSE2-fold backward (build with -DFASN_BWD_FOLD_LSE=1):
parity of a few small cases against the float64 oracle, then the C3 step time.  NOT yet run on a GPU (round 1 ran out
of GPU minutes after the bring-up probe of its operand forms passed); run it before anything else next round:

    python -c "import sys; sys.path.insert(0, 'flash-attention-softmax-n_b200'); import build; build.build_variant('fold', ['-DFASN_BWD_FOLD_LSE=1'])"
    gpurun -- 'timeout 120 python scripts/fold_check.py'
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
os.environ.setdefault("FASN_LIBRARY", os.path.join(ROOT, "flash-attention-softmax-n_b200", "flash_attention_softmax_n", "libfasn_fold.so"))
import torch
from flash_attention_softmax_n import flash_attention_n
from oracle import attention_oracle as orc

cases = [(torch.float16, 2, 2, 384, 384, 128, True, 0.25, 0.5), (torch.bfloat16, 1, 3, 200, 333, 128, False, 0.0, 1.0),
         (torch.float16, 1, 2, 130, 130, 64, True, 0.0, 0.0), (torch.bfloat16, 2, 1, 512, 256, 64, False, 0.25, 2.0)]
for dtype, B, H, L, S, D, causal, p, n in cases:
    g = torch.Generator().manual_seed(L + S)
    q, k, v = ((torch.randn(B, H, m, D, generator=g) * 0.5).to(dtype).cuda().requires_grad_() for m in (L, S, S))
    do = torch.randn(B, H, L, D, generator=g).to(dtype).cuda()
    o = flash_attention_n(q, k, v, softmax_n_param=n, dropout_p=p, is_causal=causal, _philox=(7, 3))
    o.backward(do)
    torch.cuda.synchronize()
    keep = orc.dropout_keep_mask(7, 3, B, H, L, S, p) if p > 0 else None
    want = orc.attention_fwd_bwd(q.detach().cpu(), k.detach().cpu(), v.detach().cpu(), do.cpu(), softmax_n_param=n, is_causal=causal,
                                 keep_mask=keep, dropout_p=p)
    rels = [orc.rel_l2(a, b) for a, b in zip((o, q.grad, k.grad, v.grad), want)]
    print(dtype, (B, H, L, S, D), "causal" if causal else "full", "p", p, "n", n, "rel-L2 O/dQ/dK/dV", ["%.1e" % r for r in rels], flush=True)
    assert max(rels) < (5e-3 if dtype == torch.bfloat16 else 1e-3)

B, H, S, D = 4, 32, 4096, 128
q, k, v = (torch.empty(B, H, S, D, device="cuda", dtype=torch.float16).normal_(0, 0.5).requires_grad_() for _ in range(3))
do = torch.randn(B, H, S, D, device="cuda", dtype=torch.float16)
def step(i):
    q.grad = k.grad = v.grad = None
    flash_attention_n(q, k, v, softmax_n_param=0.5, is_causal=True, dropout_p=0.1, _philox=(1, i)).backward(do)
for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    step(10 + i)
e1.record(); torch.cuda.synchronize()
print("C3 step %.3f ms (library %s)" % (e0.elapsed_time(e1) / 20, os.path.basename(os.environ["FASN_LIBRARY"])))
