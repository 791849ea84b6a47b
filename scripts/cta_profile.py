"""Per-CTA lifetimes of the forward and backward main kernels on the headline workload (needs libfasn_timeline.so: build.py --timeline).
Every CTA records globaltimer at entry / exit, its clock64 cycle count, its SM and its iteration count.  Prints a least-squares fit
lifetime = overhead + period * iterations, the per-SM busy share of the kernel's span and the idle tail; writes
gpurun_out/cta_profile_<tag>.json."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
os.environ["FASN_LIBRARY"] = os.path.join(ROOT, "flash-attention-softmax-n_b200", "flash_attention_softmax_n", "libfasn_timeline.so")
import numpy as np
import torch
from flash_attention_softmax_n import flash_attention_n, _native

tag = sys.argv[1] if len(sys.argv) > 1 else "cta"
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
lib = _native.load()
B, H, S, D = 4, 32, 4096, 128
n_fwd, n_bwd = B * H * (S // 256), B * H * (S // 128)
bf = torch.zeros(5 * 2048 + 4 * n_fwd, dtype=torch.int64, device="cuda")
bb = torch.zeros(5 * 2048 + 4 * n_bwd, dtype=torch.int64, device="cuda")
for f in (lib.fasn_set_timeline_fwd, lib.fasn_set_timeline):
    f.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
q, k, v = (torch.empty(B, H, S, D, device="cuda", dtype=torch.float16).normal_(0, 0.5).requires_grad_() for _ in range(3))
do = torch.randn(B, H, S, D, device="cuda", dtype=torch.float16)
for i in range(4):
    if i == 3:
        lib.fasn_set_timeline_fwd(bf.data_ptr(), 0, 70)
        lib.fasn_set_timeline(bb.data_ptr(), 4, 70)
    q.grad = k.grad = v.grad = None
    flash_attention_n(q, k, v, softmax_n_param=0.5, is_causal=True, dropout_p=drop, _philox=(1, i)).backward(do)
torch.cuda.synchronize()
out = {}
for name, buf, n in (("fwd", bf, n_fwd), ("bwd", bb, n_bwd)):
    rec = buf.cpu().numpy()[5 * 2048:].reshape(n, 4).astype(np.uint64)
    rec = rec[rec[:, 1] > 0]              # persistent kernels launch one CTA per SM: only those rows are written
    n = len(rec)
    g0, g1, cyc = rec[:, 0].astype(np.float64), rec[:, 1].astype(np.float64), rec[:, 2].astype(np.float64)
    sm, iters = (rec[:, 3] >> np.uint64(32)).astype(np.int64), (rec[:, 3] & np.uint64(0xFFFFFFFF)).astype(np.float64)
    A = np.stack([np.ones_like(iters), iters], 1)
    (ov, per), *_ = np.linalg.lstsq(A, cyc, rcond=None)
    span = g1.max() - g0.min()
    busy = np.array([(g1[sm == s] - g0[sm == s]).sum() for s in np.unique(sm)])
    last = np.array([g1[sm == s].max() for s in np.unique(sm)]) - g0.min()
    first = np.array([g0[sm == s].min() for s in np.unique(sm)]) - g0.min()
    gaps = []
    for s in np.unique(sm):
        o = np.argsort(g0[sm == s]); a0, a1 = g0[sm == s][o], g1[sm == s][o]
        gaps += list(a0[1:] - a1[:-1])
    res = dict(ctas=int(n), sms=int(len(busy)), span_us=span / 1e3, cycles_total=float(cyc.sum()), iters_total=float(iters.sum()),
               fit_overhead_cycles=float(ov), fit_period_cycles=float(per), mean_cycles_per_iter=float(cyc.sum() / iters.sum()),
               clock_ghz_est=float(cyc.sum() / (g1 - g0).sum()), busy_share_mean=float((busy / span).mean()), busy_share_min=float((busy / span).min()),
               last_end_us_min=float(last.min() / 1e3), last_end_us_median=float(np.median(last) / 1e3), first_start_us_max=float(first.max() / 1e3),
               gap_ns_median=float(np.median(gaps)), gap_ns_mean=float(np.mean(gaps)), ctas_per_sm_min=int(np.bincount(sm).min()), ctas_per_sm_max=int(np.bincount(sm).max()))
    out[name] = res
    print(name, json.dumps(res))
    # per iteration-count buckets: mean lifetime
    for it in sorted(set(iters.tolist()))[:: max(1, len(set(iters.tolist())) // 8)]:
        m = iters == it
        print(f"   iters {int(it):3d}: {int(m.sum()):4d} CTAs, mean {cyc[m].mean():9.0f} cycles = {cyc[m].mean() / max(it, 1):7.0f} / iter")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"cta_profile_{tag}.json"), "w"), indent=1)
