"""BASELINE.json configs[3] end to end from root-held tensors: fwd bf16 B=64 H=40 S=8192 D=128 n=1 causal, Q/K/V on rank 0,
(batch, head) slabs scattered over the ranks with NCCL point-to-point sends, O gathered back (parallel.sharded_attention).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_sharded.py
Prints one JSON line per `chunks` setting: whole-job milliseconds (CUDA events on the root, barrier + synchronize on both
sides), end-to-end TFLOP/s, bytes the root sends / receives, and the kernel-only time of one rank's slab for comparison."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
import torch
import torch.distributed as dist
from flash_attention_softmax_n import flash_attention_n
from flash_attention_softmax_n.parallel import sharded_attention, partition_units

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, S, D = (int(x) for x in os.environ.get("SHAPE", "64,40,8192,128").split(","))
dtype = torch.bfloat16
kw = dict(softmax_n_param=1.0, is_causal=True)
flops = 4.0 * B * H * S * S * D * 0.5
q = k = v = None
if rank == 0:
    torch.manual_seed(1234)
    q, k, v = (torch.empty(B, H, S, D, device=dev, dtype=dtype).normal_(0, 0.5) for _ in range(3))

def run(chunks):
    return sharded_attention(q, k, v, shape=(B, H, S, S, D), dtype=dtype, device=dev, chunks=chunks, **kw)

def timed(fn, reps):
    fn(); fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), out

# kernel phase alone: this rank's slab, resident
lo, hi = partition_units(B * H, world)[rank]
ql, kl, vl = (torch.empty(1, hi - lo, S, D, device=dev, dtype=dtype).normal_(0, 0.5) for _ in range(3))
t_kernel, _ = timed(lambda: flash_attention_n(ql, kl, vl, _bh_offset=lo, **kw), 5)
del ql, kl, vl
ref = None
for chunks in [int(c) for c in os.environ.get("CHUNKS", "1,4,8,16").split(",")]:
    t, out = timed(lambda: run(chunks), 3)
    if rank == 0:
        if ref is None:
            ref = out
        same = bool(torch.equal(out, ref))
        unit = S * D * 2
        sent = 3 * (B * H - (hi - lo)) * unit
        print(json.dumps({"workload": f"fwd bf16 B={B} H={H} S={S} D={D} n=1 causal, root-held Q/K/V, {world} GPUs", "chunks": chunks,
                          "ms": t, "e2e_tflops": flops / (t * 1e-3) / 1e12, "kernel_phase_ms": t_kernel,
                          "kernel_phase_tflops_all_gpus": flops / (t_kernel * 1e-3) / 1e12,
                          "root_sends_GB": sent / 1e9, "root_receives_GB": (B * H - (hi - lo)) * unit / 1e9,
                          "root_egress_GBs": sent / (t * 1e-3) / 1e9, "identical_to_chunks_1": same}), flush=True)
dist.destroy_process_group()
