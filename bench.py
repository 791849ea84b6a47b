#!/usr/bin/env python
"""bench.py -- headline measurement for the fused softmax_n attention path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:
for the headline workload (BASELINE.json configs[2], "c3": fwd+bwd fp16 B=4 H=32 S=4096 D=128 n=0.5 causal
dropout 0.1) that is one forward and one backward of `flash_attention_n`, four kernel launches.

  value      whole-job algorithmic TFLOP/s with Q,K,V,dO resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the public API with HOST (pinned) buffers: H2D of q,k,v,dO and D2H of
             o,dq,dk,dv inside the timed region; the process is bound to the CPUs next to its GPU first, and the raw
             pinned-copy rates of every rank are reported beside it
  roofline   the dominant kernel (backward main kernel) against the measured bf16 tensor peak of the clock regime the timed
             region ran in: "burst" (< 1 s of load at >= 95 % of the maximum SM clock) or "sustained" (power-capped)
  sustained  a second, seconds-long leg of the same step: TFLOP/s and clocks under the 1000 W cap
  cpu_baseline  the oracle port of the reference's slow_attention_n on this box's host cores (bounded sample)

With N > 1 (torchrun), every rank runs the same per-GPU workload on its own resident slab of (batch, head)
units (weak scaling, no collective on the data path); the time is the max over ranks.  In addition the `sharded` object
times BASELINE.json configs[3] (fwd bf16 S=8192 D=128 causal, 320 units per GPU) with Q/K/V held by rank 0:
parallel.sharded_attention scatters the slabs over NVLink, every rank runs the kernel, O is gathered back, and the result
is compared bit for bit with rank 0 computing the same units alone.

--impl reference times the reference's own CPU implementation of the path (the oracle port: the reference is
pure Python, there is nothing to compile into oracle/_ref) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "flash-attention-softmax-n_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "attention fwd+bwd TFLOPS/GPU and % tensor-core peak at S=4096 D=128 bf16"

WORKLOADS = {
    # name: B, H, S, D, dtype, n, causal, dropout, backward
    "c3": dict(B=4, H=32, S=4096, D=128, dtype="f16", n=0.5, causal=True, dropout=0.1, bwd=True,
               desc="BASELINE.json configs[2]: fwd+bwd fp16 B=4 H=32 S=4096 D=128 n=0.5 causal dropout=0.1"),
    "c3bf16": dict(B=4, H=32, S=4096, D=128, dtype="bf16", n=0.5, causal=True, dropout=0.1, bwd=True,
                   desc="configs[2] in bf16"),
    "c3nd": dict(B=4, H=32, S=4096, D=128, dtype="f16", n=0.5, causal=True, dropout=0.0, bwd=True,
                 desc="configs[2] without dropout (diagnostic)"),
    "c3pad": dict(B=4, H=32, S=4096, D=128, dtype="f16", n=0.5, causal=True, dropout=0.0, bwd=True, aux="padmask",
                  desc="configs[2] shape, no dropout, dense key-padding attn_mask (B,1,1,S) AND causal (diagnostic)"),
    "c3alibi": dict(B=2, H=16, S=4096, D=128, dtype="f16", n=0.5, causal=True, dropout=0.0, bwd=True, aux="alibi",
                    desc="B=2 H=16 S=4096 D=128 with a dense ALiBi attn_bias (H,L,S) (diagnostic; 512 MiB of bias per pass)"),
    "c3alibis": dict(B=2, H=16, S=4096, D=128, dtype="f16", n=0.5, causal=True, dropout=0.0, bwd=True, aux="alibi_slopes",
                     desc="as c3alibi, but the ALiBi bias is generated in the kernels from 16 slopes (_alibi_slopes) (diagnostic)"),
    "c2": dict(B=8, H=16, S=2048, D=64, dtype="bf16", n=1.0, causal=False, dropout=0.0, bwd=False,
               desc="BASELINE.json configs[1]: fwd bf16 B=8 H=16 S=2048 D=64 n=1 non-causal"),
    "c4": dict(B=8, H=40, S=8192, D=128, dtype="bf16", n=1.0, causal=True, dropout=0.0, bwd=False,
               desc="BASELINE.json configs[3] per-GPU share: fwd bf16 320 (batch,head) units S=8192 D=128 n=1 causal"),
    "c5": dict(B=1, H=16, S=65536, D=64, dtype="bf16", n=1.0, causal=True, dropout=0.0, bwd=False,
               desc="BASELINE.json configs[4]: fwd bf16 B=1 H=16 S=65536 D=64 n=1 causal"),
}


def algorithmic_flops(w):
    """F_fwd = 4 B H Sq Skv D (x 1/2 causal); F_bwd = 2.5 F_fwd."""
    f = 4.0 * w["B"] * w["H"] * w["S"] * w["S"] * w["D"] * (0.5 if w["causal"] else 1.0)
    return f, (2.5 * f if w["bwd"] else 0.0)


def algorithmic_bytes(w):
    """HBM bytes a step must move: forward reads Q, K, V and writes O (+ fp32 LSE); the backward
    reads Q, K, V, O, dO (+ LSE, delta) and writes dQ, dK, dV."""
    rows, es = w["B"] * w["H"] * w["S"], 2
    fwd = 4 * rows * w["D"] * es + 4 * rows
    return fwd, ((8 * rows * w["D"] * es + 8 * rows) if w["bwd"] else 0)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.samples = []
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9 and f[1].isdigit():
                self.samples.append(f)

    def summary(self):
        if not getattr(self, "samples", None):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [int(f[1]) for f in self.samples]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(f[5 + i].lower().startswith("active") for f in self.samples)]
        pw = [float(f[3]) for f in self.samples if f[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": int(self.samples[0][2]), "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of slow_attention_n on the host cores
# --------------------------------------------------------------------------------------------------------------

def cpu_sample(w, budget_s, threads):
    """Time fwd(+bwd) of the oracle on single (batch, head) units of the workload, fp32, all host threads.
    Units are independent, so units/s extrapolates linearly; returns (TFLOP/s, units timed, seconds, description)."""
    from oracle import attention_oracle as orc
    torch.set_num_threads(threads)
    S, D = w["S"], w["D"]
    rows = S
    if w["S"] > 8192:                      # c5: the (S x S) score matrix does not fit; time one 4096-row query block
        rows = 4096
    g = torch.Generator().manual_seed(1234)
    q = (torch.randn(1, 1, rows, D, generator=g) * 0.5)
    k = (torch.randn(1, 1, S, D, generator=g) * 0.5)
    v = (torch.randn(1, 1, S, D, generator=g) * 0.5)
    do = torch.randn(1, 1, rows, D, generator=g)
    keep = None
    if w["dropout"] > 0:
        keep = orc.dropout_keep_mask(0x5EED, 0, 1, 1, rows, S, w["dropout"])
    kw = dict(softmax_n_param=w["n"], is_causal=w["causal"], keep_mask=keep, dropout_p=w["dropout"])

    def one():
        if w["bwd"]:
            orc.attention_fwd_bwd(q, k, v, do, dtype=torch.float32, **kw)
        else:
            with torch.no_grad():
                orc.slow_attention_n(q, k, v, **kw)

    one()                                  # warm-up (thread pool, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        one()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 64:
            break
    f_fwd, f_bwd = algorithmic_flops(dict(w, B=1, H=1))
    unit_flops = (f_fwd + f_bwd) * (rows / S if rows != S else 1.0)
    if rows != S and w["causal"]:          # bottom-right aligned block of `rows` queries sees all S keys minus a triangle
        unit_flops = 4.0 * D * (rows * S - rows * (rows - 1) / 2.0)
    tflops = unit_flops * n / el / 1e12
    desc = (f"{n} x one (batch,head) unit of the workload ({rows} query rows x {S} keys, D={D}, fp32, "
            f"{'fwd+bwd' if w['bwd'] else 'fwd'}), {el:.1f} s on {threads} threads; units are independent, "
            f"so whole-workload throughput is the same figure")
    return tflops, n, el, desc


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = []
    for _ in range(args.warmup):
        cpu_sample(w, 0.5, threads)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        tf, n, el, desc = cpu_sample(w, args.cpu_step_seconds, threads)
        per_step.append((tf, el))
    wall = time.perf_counter() - t_all
    tf = sum(t for t, _ in per_step) / len(per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup,
        # one whole step of the workload at the measured rate (units are independent: linear extrapolation of the sample);
        # the wall time actually spent per sampled step is `sample_ms_per_step`
        "ms_per_step": 1e3 * sum(algorithmic_flops(w)) / (tf * 1e12), "sample_ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": desc},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# N > 1: BASELINE.json configs[3] from root-held tensors (NVLink scatter / compute / gather)
# --------------------------------------------------------------------------------------------------------------

def run_sharded(args, dev, rank, world):
    """fwd bf16 S=8192 D=128 n=1 causal on `units_per_gpu` x world (batch, head) units held by rank 0 (B=64 H=40 = 2560 units
    on 8 GPUs).  Returns the `sharded` object on rank 0 (None elsewhere).  All ranks must call it."""
    import torch.distributed as dist
    from flash_attention_softmax_n import flash_attention_n
    from flash_attention_softmax_n.parallel import sharded_attention, partition_units, IpcSlabs

    S, D, dtype = 8192, 128, torch.bfloat16
    U = args.sharded_units_per_gpu * world
    shape = (1, U, S, S, D)
    kw = dict(softmax_n_param=1.0, is_causal=True)
    flops = 4.0 * U * S * S * D * 0.5
    q = k = v = None
    if rank == 0:
        torch.manual_seed(4321)
        q, k, v = (torch.empty(1, U, S, D, device=dev, dtype=dtype).normal_(0, 0.5) for _ in range(3))

    def timed(fn, reps):
        fn()
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

    lo, hi = partition_units(U, world)[rank]
    ql, kl, vl = (torch.empty(1, hi - lo, S, D, device=dev, dtype=dtype).normal_(0, 0.5) for _ in range(3))
    timed(lambda: flash_attention_n(ql, kl, vl, _bh_offset=lo, **kw), 3)          # clocks up after the host-bound e2e leg
    t_kernel, _ = timed(lambda: flash_attention_n(ql, kl, vl, _bh_offset=lo, **kw), 5)
    del ql, kl, vl
    res = {"workload": f"BASELINE.json configs[3] scaled to {world} GPU(s): fwd bf16 {U} (batch,head) units S={S} D={D} n=1 causal, Q/K/V held by rank 0",
           "chunks": args.sharded_chunks, "kernel_phase_ms": t_kernel, "kernel_phase_tflops_all_gpus": flops / (t_kernel * 1e-3) / 1e12}
    unit_bytes = S * D * 2
    sent = 3 * (U - (hi - lo)) * unit_bytes if rank == 0 else 0
    out = None
    try:
        slabs = IpcSlabs(shape, dtype, dev, args.sharded_chunks)
        t_ipc, out = timed(lambda: sharded_attention(q, k, v, shape=shape, dtype=dtype, device=dev, chunks=args.sharded_chunks,
                                                      transport="ipc", slabs=slabs, **kw), 3)
        res.update({"transport": "CUDA IPC peer copies on copy engines + 4-byte NCCL all-reduce per pipeline step", "e2e_ms": t_ipc})
        del slabs
    except Exception as e:      # reported, not hidden: the p2p numbers below then stand for the leg
        res["ipc_error"] = f"{type(e).__name__}: {e}"[:300]
    t_p2p, out_p2p = timed(lambda: sharded_attention(q, k, v, shape=shape, dtype=dtype, device=dev, chunks=min(4, args.sharded_chunks),
                                                      transport="p2p", **kw), 2)
    res["p2p_e2e_ms"] = t_p2p
    if "e2e_ms" not in res:
        res.update({"transport": "NCCL point-to-point sends", "e2e_ms": t_p2p})
        out = out_p2p
    if rank == 0:
        ref = torch.empty_like(out)
        for a in range(0, U, args.sharded_units_per_gpu):            # rank 0 alone, slab by slab
            b = min(U, a + args.sharded_units_per_gpu)
            ref[:, a:b] = flash_attention_n(q[:, a:b], k[:, a:b], v[:, a:b], _bh_offset=a, **kw)
        torch.cuda.synchronize()
        res["bit_identical"] = bool(torch.equal(out, ref)) and bool(torch.equal(out_p2p, ref))
        res.update({"root_sends_GB": sent / 1e9, "root_receives_GB": (U - (hi - lo)) * unit_bytes / 1e9,
                    "root_egress_GBs": sent / (res["e2e_ms"] * 1e-3) / 1e9, "e2e_tflops": flops / (res["e2e_ms"] * 1e-3) / 1e12,
                    "egress_floor_ms_at_770_GBs": sent / 770e9 * 1e3})
        return res
    return None


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------

def run_ours(args, w):
    import torch.distributed as dist
    from flash_attention_softmax_n import flash_attention_n, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fused kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = _native.load()
    dtype = torch.float16 if w["dtype"] == "f16" else torch.bfloat16
    B, H, S, D = w["B"], w["H"], w["S"], w["D"]
    units = B * H
    torch.manual_seed(1234 + rank)
    # synthetic inputs exactly as the reference's tests draw them: N(0, 0.5^2), dO ~ N(0,1) (tests/common.py:18-20)
    q = torch.empty(B, H, S, D, device=dev, dtype=dtype).normal_(0, 0.5).requires_grad_(w["bwd"])
    k = torch.empty(B, H, S, D, device=dev, dtype=dtype).normal_(0, 0.5).requires_grad_(w["bwd"])
    v = torch.empty(B, H, S, D, device=dev, dtype=dtype).normal_(0, 0.5).requires_grad_(w["bwd"])
    do = torch.randn(B, H, S, D, device=dev, dtype=dtype)
    kw = dict(softmax_n_param=w["n"], is_causal=w["causal"], dropout_p=w["dropout"], _bh_offset=rank * units)
    if w.get("aux") == "padmask":
        lens = torch.randint(S // 2, S + 1, (B,), device=dev)
        kw["attn_mask"] = (torch.arange(S, device=dev)[None, :] < lens[:, None]).view(B, 1, 1, S)
    elif w.get("aux") == "alibi":
        slopes = 2.0 ** (-8.0 * torch.arange(1, H + 1, device=dev) / H)
        dist_ = (torch.arange(S, device=dev)[None, :] - torch.arange(S, device=dev)[:, None]).float()
        kw["attn_bias"] = (slopes[:, None, None] * dist_[None]).to(dtype)
    elif w.get("aux") == "alibi_slopes":
        kw["_alibi_slopes"] = 2.0 ** (-8.0 * torch.arange(1, H + 1, device=dev) / H)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if args.flush_l2 else None

    def step(i):
        if flush is not None:
            flush.fill_(i & 0xFF)
        if w["dropout"] > 0:
            kw["_philox"] = (0x5EED, i)
        o = flash_attention_n(q, k, v, **kw)
        if w["bwd"]:
            q.grad = k.grad = v.grad = None
            o.backward(do)
        return o

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region: resident inputs -------------------------------------------------------------------
    lib.fasn_profile(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(100 + i)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    fwd_ms, fwd_n, bwd_ms, bwd_n = ctypes.c_double(), ctypes.c_int32(), ctypes.c_double(), ctypes.c_int32()
    _native.check(lib.fasn_profile_read(ctypes.byref(fwd_ms), ctypes.byref(fwd_n), ctypes.byref(bwd_ms), ctypes.byref(bwd_n)),
                  "fasn_profile_read")
    lib.fasn_profile(0)
    if flush is not None:                   # the flush kernel sits inside the event bracket: measure and subtract it
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
        f1.record()
        torch.cuda.synchronize()
        ms -= f0.elapsed_time(f1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    f_fwd, f_bwd = algorithmic_flops(w)
    flops_step = f_fwd + f_bwd
    value = world * flops_step * args.steps / (ms_max * 1e-3) / 1e12

    # ---- second leg: the same step for a few seconds (power-capped regime) ------------------------------------------
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(ms_max / args.steps, 1e-3)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as sclocks:
            barrier()
            s0.record()
            for i in range(n_sus):
                step(10000 + i)
            s1.record()
            barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if flush is not None:
            ts -= n_sus * (f0.elapsed_time(f1) / args.steps)
        if distributed:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"steps": n_sus, "seconds": ts.item() * 1e-3, "ms_per_step": ts.item() / n_sus,
                     "value": world * flops_step * n_sus / (ts.item() * 1e-3) / 1e12, "unit": "TFLOP/s", "clocks": sclocks.summary()}

    # ---- end to end: host buffers through the public API ----------------------------------------------------
    # flash_attention_softmax_n.host.attention_host: pinned host tensors in, pinned host tensors out; chunks of
    # (batch, head) units are pipelined over three streams so H2D, kernels and D2H overlap.
    e2e = None
    if not args.no_e2e and not w.get("aux"):
        from flash_attention_softmax_n.host import HostPipeline, bind_process_to_gpu, pinned_copy_rates
        binding = bind_process_to_gpu(local_rank)      # pinned buffers below are first touched on the GPU's NUMA node
        hq, hk, hv, hdo = (torch.empty(units, S, D, dtype=dtype).normal_(0, 0.5).pin_memory() for _ in range(4))
        ho = torch.empty(units, S, D, dtype=dtype).pin_memory()
        hg = tuple(torch.empty(units, S, D, dtype=dtype).pin_memory() for _ in range(3)) if w["bwd"] else None
        pipe = HostPipeline(units, S, S, D, dtype, dev, chunks=args.e2e_chunks, backward=w["bwd"])
        ekw = dict(softmax_n_param=w["n"], is_causal=w["causal"], dropout_p=w["dropout"])

        def e2e_step(i):
            if w["dropout"] > 0:
                ekw["_philox"] = (0x5EED, 1000 + i)
            pipe.run(hq, hk, hv, hdo if w["bwd"] else None, ho, hg, **ekw)      # returns with the results on the host

        for i in range(2):
            e2e_step(i)
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_e2e):
            e2e_step(i)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        nbytes = units * S * D * 2
        # raw pinned-copy rates of this rank while every rank copies at once: the ceiling the end-to-end leg runs against
        barrier()
        h2d_gbs, d2h_gbs, both_gbs = pinned_copy_rates(dev, hq, ho)
        rates = torch.tensor([h2d_gbs, d2h_gbs, both_gbs], device=dev, dtype=torch.float64)
        rmin, rsum = rates.clone(), rates.clone()
        if distributed:
            dist.all_reduce(rmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(rsum, op=dist.ReduceOp.SUM)
        step_bytes = nbytes * ((4 + 4) if w["bwd"] else (3 + 1))
        e2e = {"value": world * flops_step * n_e2e / (t2.item() * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": nbytes * (4 if w["bwd"] else 3), "d2h_bytes_per_step": nbytes * (4 if w["bwd"] else 1),
               "ms_per_step": t2.item() / n_e2e, "steps": n_e2e,
               "api": f"flash_attention_softmax_n.host.attention_host, {pipe.chunks} chunks of units pipelined over 3 streams",
               "host_binding": binding,
               "pinned_copy_GBs_per_gpu_min": {"h2d_alone": rmin[0].item(), "d2h_alone": rmin[1].item(), "both_directions_sum": rmin[2].item()},
               "pinned_copy_GBs_all_gpus": {"h2d_alone": rsum[0].item(), "d2h_alone": rsum[1].item(), "both_directions_sum": rsum[2].item()},
               "copy_bound_ms_per_step": step_bytes / max(rmin[2].item(), 1e-9) / 1e6,
               "limiter": "host<->device copies: the step moves %.2f GB per GPU; at the measured full-duplex pinned-copy rate of the "
                          "slowest rank (all ranks copying at once) that alone is the copy_bound_ms_per_step" % (step_bytes / 1e9)}

    # ---- N > 1: root-held configs[3] through the NVLink scatter / gather ----------------------------------------
    sharded = None
    if distributed and not args.no_sharded:
        del q, k, v, do
        torch.cuda.empty_cache()
        sharded = run_sharded(args, dev, rank, world)

    # ---- roofline of the dominant kernel --------------------------------------------------------------------
    peaks = measured_peaks()
    if w["bwd"] and bwd_n.value > 0:
        k_ms, k_flops, k_name = bwd_ms.value / bwd_n.value, f_bwd, "fasn_bwd_kernel (backward main)"
    else:
        k_ms, k_flops, k_name = fwd_ms.value / max(fwd_n.value, 1), f_fwd, "fasn_fwd_kernel"
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    # the denominator follows the clock record of the timed region itself: a sub-second region at (nearly) the maximum SM
    # clock is the burst regime of MEASURED_PEAKS.json, a long or clock-limited one the sustained regime
    csum = clocks.summary()
    near_max = csum.get("sm_mhz") is not None and csum.get("sm_max_mhz") and csum["sm_mhz"] >= 0.95 * csum["sm_max_mhz"]
    regime = "burst" if (ms_max < 1000.0 and (near_max or csum.get("sm_mhz") is None)) else "sustained"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks[regime], "unit": "TFLOP/s",
                "frac": achieved / peaks[regime], "traffic": None, "kernel": k_name, "kernel_ms": k_ms,
                "regime": regime, "regime_rule": "burst if the timed region is < 1 s and the median SM clock >= 95 % of max, else sustained",
                "peak_kind": regime + " bf16 dense, " + peaks["source"], "frac_of_burst_peak": achieved / peaks["burst"],
                "frac_of_sustained_peak": achieved / peaks["sustained"],
                "fwd_kernel_ms": fwd_ms.value / max(fwd_n.value, 1),
                "fwd_kernel_tflops": f_fwd / (fwd_ms.value / max(fwd_n.value, 1) * 1e-3) / 1e12 if fwd_n.value else None}
    # secondary: the step's algorithmic HBM traffic rate (every workload here is far above the 248 FLOP/B ridge, C5 included)
    b_fwd, b_bwd = algorithmic_bytes(w)
    roofline["hbm_algorithmic_bytes_per_step"] = b_fwd + b_bwd
    roofline["hbm_algorithmic_gbps"] = (b_fwd + b_bwd) * args.steps / (ms_max * 1e-3) / 1e9      # per GPU
    roofline["hbm_frac_of_copy_peak"] = roofline["hbm_algorithmic_gbps"] / peaks["hbm"]
    roofline["flop_per_byte"] = flops_step / (b_fwd + b_bwd)
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            pass

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            tf, n, el, desc = cpu_sample(w, args.cpu_seconds, os.cpu_count() or 1)
            cpu = {"value": tf, "unit": "TFLOP/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu_units": units, "parallelism": (f"value: {world} GPU(s), each on its resident slab of (batch, head) units, no data-path collective; "
                                       "sharded: root-held tensors scattered / gathered over NVLink (parallel.sharded_attention)" if world > 1 else
                                       "1 GPU, all units resident"),
                       "l2": "flush between steps (256 MiB write)" if args.flush_l2 else
                             f"working set {(8 if w['bwd'] else 4) * B * H * S * D * 2 / 2**20:.0f} MiB per step > 126 MB L2, no flush"},
            "per_gpu_tflops": value / world, "frac_of_peak": value / world / peaks[regime],
            "frac_of_burst_peak": value / world / peaks["burst"], "frac_of_sustained_peak": value / world / peaks["sustained"],
            "roofline": roofline, "sustained": sustained, "sharded": sharded, "cpu_baseline": cpu, "e2e": e2e, "clocks": csum,
            "gpu_launches": world * args.steps * ((1 + 3) if w["bwd"] else 1),
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--flush-l2", action="store_true", help="write 256 MiB between steps (default for small workloads)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget inside the default run")
    ap.add_argument("--cpu-step-seconds", type=float, default=6.0, help="--impl reference: CPU seconds per step")
    ap.add_argument("--sustained-seconds", type=float, default=1.5, help="length of the second, power-capped leg (0 disables)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the root-held scatter / gather leg")
    ap.add_argument("--sharded-chunks", type=int, default=16)
    ap.add_argument("--sharded-units-per-gpu", type=int, default=320)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.workload == "c2":
        args.flush_l2 = True               # 134 MB working set ~ L2 size
    if args.impl == "reference":
        args.cpu_step_seconds = min(args.cpu_step_seconds, 150.0 / max(args.steps, 1))   # whole run within a few minutes
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
