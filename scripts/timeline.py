"""Phase timeline of one backward CTA (needs libfasn_timeline.so: build.py --timeline).  Writes gpurun_out/timeline_<tag>.txt.
Tags: MMA thread 1 S^T(0) issue, 2 p_full seen, 3 q_full(next) seen, 4 ds_full seen, 5 do_full(next) seen, 6 dq_empty seen;
compute 10 iteration start, 11 S^T ready, 12 P^T stored, 13 dP^T ready, 14 dS buffer free, 15 dS stored;
reducer 30 dQ ready, 31 dQ drained from TMEM; persistent kernel: the CTA given as x is recorded over all of its items --
compute 20 item fetched, 21 last dV MMA done, 22 every MMA of the item done, 23 dK staged; producer 40 first Q / dO tile of an
item issued, 41 sK free (K load issued), 42 sV free (V load issued)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
os.environ["FASN_LIBRARY"] = os.path.join(ROOT, "flash-attention-softmax-n_b200", "flash_attention_softmax_n", "libfasn_timeline.so")
import torch
from flash_attention_softmax_n import flash_attention_n, _native

tag = sys.argv[1] if len(sys.argv) > 1 else "tl"
x, y = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 70)
lib = _native.load()
buf = torch.zeros(5 * 2048 + 4 * 65536, dtype=torch.int64, device="cuda")   # role timelines + the per-CTA records behind them
lib.fasn_set_timeline.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
B, H, S, D = 4, 32, 4096, 128
q, k, v = (torch.empty(B, H, S, D, device="cuda", dtype=torch.float16).normal_(0, 0.5).requires_grad_() for _ in range(3))
do = torch.randn(B, H, S, D, device="cuda", dtype=torch.float16)
for i in range(3):
    if i == 2:
        lib.fasn_set_timeline(buf.data_ptr(), x, y)
    o = flash_attention_n(q, k, v, softmax_n_param=0.5, is_causal=True, dropout_p=0.1, _philox=(1, i))
    o.backward(do)
torch.cuda.synchronize()
ev = buf.cpu()[:5 * 2048].view(5, 2048)
rows = []
for role in range(5):
    for e in ev[role].tolist():
        if e:
            rows.append((e & 0xFFFFFFFFFFFF, role, (e >> 48) & 0xFFFF))
rows.sort()
t0 = rows[0][0]
names = {0: "mma", 1: "cmp0", 2: "cmp1", 3: "red", 4: "load"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.txt"), "w") as f:
    for t, role, tg in rows:
        f.write(f"{t - t0:9d} {names[role]:5s} {tg}\n")
print("events", len(rows), "span", rows[-1][0] - t0, "cycles")
