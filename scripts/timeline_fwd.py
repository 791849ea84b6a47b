"""Phase timeline of one forward CTA (libfasn_timeline.so).  Tags: MMA thread 1 first QK issue, 2 P0 seen, 3 P1 seen, 4 end of
iteration issue; softmax (role 1 = tile 0, 2 = tile 1) 10 iteration start, 11 S loads issued, 12 S in registers, 13 max done,
14 exp/pack done, 15 P stored + arrived."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
os.environ["FASN_LIBRARY"] = os.path.join(ROOT, "flash-attention-softmax-n_b200", "flash_attention_softmax_n", "libfasn_timeline.so")
import torch
from flash_attention_softmax_n import flash_attention_n, _native
tag = sys.argv[1] if len(sys.argv) > 1 else "tlf"
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
lib = _native.load()
buf = torch.zeros(5 * 2048, dtype=torch.int64, device="cuda")
lib.fasn_set_timeline_fwd.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
B, H, S, D = 4, 32, 4096, 128
q, k, v = (torch.empty(B, H, S, D, device="cuda", dtype=torch.float16).normal_(0, 0.5) for _ in range(3))
for i in range(3):
    if i == 2:
        lib.fasn_set_timeline_fwd(buf.data_ptr(), 0, 70)     # blockIdx.x = 0 is the heaviest causal block (last 256 rows)
    o = flash_attention_n(q, k, v, softmax_n_param=0.5, is_causal=True, dropout_p=drop, _philox=(1, i))
torch.cuda.synchronize()
ev = buf.cpu().view(5, 2048)
rows = sorted((e & 0xFFFFFFFFFFFF, role, (e >> 48) & 0xFFFF) for role in range(5) for e in ev[role].tolist() if e)
t0 = rows[0][0]
names = {0: "mma", 1: "sm0", 2: "sm1", 3: "gen"}
with open(os.path.join(ROOT, "gpurun_out", f"timeline_fwd_{tag}.txt"), "w") as f:
    for t, role, tg in rows:
        f.write(f"{t - t0:9d} {names.get(role, role)} {tg}\n")
print("events", len(rows), "span", rows[-1][0] - t0)
