"""Bandwidth of the fused softmax_n kernels against the HBM roofline (algorithmic bytes: one read + one write per element
forward; two reads + one write backward), with the eager definition (four elementwise passes) beside it.
    python scripts/bench_softmax.py > gpurun_out/softmax_bench.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
import torch
from flash_attention_softmax_n import softmax_n_fused, softmax_n

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = []
for name, shape, dtype in [("scores B8 H16 L2048 S2048 fp16 (C2 shape)", (8 * 16 * 2048, 2048), torch.float16),
                           ("scores B4 H32 L1024 S4096 bf16", (4 * 32 * 1024, 4096), torch.bfloat16),
                           ("BERT-base scores B64 H12 L512 S512 fp32", (64 * 12 * 512, 512), torch.float32),
                           ("long rows 4096 x 65536 bf16", (4096, 65536), torch.bfloat16)]:
    x = torch.randn(*shape, device="cuda", dtype=dtype).requires_grad_()
    dy = torch.randn(*shape, device="cuda", dtype=dtype)
    es = x.element_size()
    t_f = timeit(lambda: softmax_n_fused(x.detach(), 1.0))
    y = softmax_n_fused(x, 1.0)
    t_b = timeit(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
    with torch.no_grad():
        t_e = timeit(lambda: softmax_n(x, 1.0), reps=3)
    nbytes = x.numel() * es
    out.append({"case": name, "fwd_ms": t_f, "fwd_GBs": 2 * nbytes / t_f / 1e6, "fwd_frac_of_hbm_peak": 2 * nbytes / t_f / 1e6 / peak,
                "bwd_ms": t_b, "bwd_GBs": 3 * nbytes / t_b / 1e6, "bwd_frac_of_hbm_peak": 3 * nbytes / t_b / 1e6 / peak,
                "eager_fwd_ms": t_e, "speedup_vs_eager": t_e / t_f})
    del x, dy, y
print(json.dumps({"hbm_peak_GBs": peak, "results": out}, indent=1))
