"""Attention on HOST-resident tensors, pipelined over (batch, head) units.

A caller whose Q/K/V (and dO) live in host memory pays PCIe for 4 inputs and 4 outputs per step, an order of
magnitude more time than the kernels.  Units (batch x head pairs) are independent, so the work is cut into chunks
of units and three CUDA streams overlap chunk i's kernels with chunk i+1's host->device copies and chunk i-1's
device->host copies; PCIe is full duplex, so a step costs about max(H2D, D2H) instead of H2D + kernels + D2H.
Dropout masks are keyed by the global unit index (`_bh_offset`), so the result does not depend on the chunking.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from flash_attention_softmax_n.core.flash_attn import flash_attention_n


def bind_process_to_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs that sit next to GPU `device_index` (NVML's CPU affinity of the device), so
    that pinned host buffers allocated afterwards are first touched -- and therefore placed -- on that GPU's NUMA node and
    the copy threads run there.  With several ranks on one host every rank otherwise inherits the same default CPU set and
    all host<->device traffic crosses one memory controller / socket link.  Returns what was done (for bench reports)."""
    import os
    info = {"gpu": device_index, "cpus_before": len(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0) or cpus
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            info["numa_node"] = None
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus), cpu_range=f"{min(cpus)}-{max(cpus)}")
        else:
            info.update(bound=False, reason="NVML reports no CPU affinity for this GPU")
    except Exception as e:      # no NVML, or a restricted container: report and carry on unbound
        info.update(bound=False, reason=f"{type(e).__name__}: {e}"[:200])
    return info


def pinned_copy_rates(device: torch.device, h_in: Tensor, h_out: Tensor, mbytes: int = 256):
    """(H2D alone, D2H alone, both directions at once: sum) in GB/s for `mbytes` of pinned memory per direction, timed with
    CUDA events on two streams.  `h_in` / `h_out` are pinned host tensors at least that large."""
    n = min(mbytes << 20, h_in.numel() * h_in.element_size(), h_out.numel() * h_out.element_size())
    src = h_in.view(-1).view(torch.uint8)[:n]
    dst = h_out.view(-1).view(torch.uint8)[:n]
    d0 = torch.empty(n, dtype=torch.uint8, device=device)
    d1 = torch.empty(n, dtype=torch.uint8, device=device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def timed(do_in: bool, do_out: bool) -> float:
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s_in.wait_event(a); s_out.wait_event(a)
        for _ in range(3):
            if do_in:
                with torch.cuda.stream(s_in):
                    d0.copy_(src, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    dst.copy_(d1, non_blocking=True)
        cur = torch.cuda.current_stream(device)
        cur.wait_stream(s_in); cur.wait_stream(s_out)
        b.record()
        torch.cuda.synchronize(device)
        return 3 * n * (int(do_in) + int(do_out)) / (a.elapsed_time(b) * 1e-3) / 1e9

    timed(True, True)
    return timed(True, False), timed(False, True), timed(True, True)


class HostPipeline:
    """Reusable buffers + streams for `attention_host`.  One instance per (shape, dtype, device)."""

    def __init__(self, units: int, L: int, S: int, D: int, dtype: torch.dtype, device: torch.device, chunks: int = 8,
                 backward: bool = True):
        self.units, self.L, self.S, self.D, self.dtype, self.device = units, L, S, D, dtype, device
        self.chunks = max(1, min(chunks, units))
        self.per = (units + self.chunks - 1) // self.chunks
        self.backward = backward
        mk = lambda n: [torch.empty(self.per, n, D, dtype=dtype, device=device) for _ in range(2)]   # double-buffered
        self.dq, self.dk, self.dv = mk(L), mk(S), mk(S)
        self.ddo = mk(L) if backward else None
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(device) for _ in range(3))
        self.in_done = [torch.cuda.Event() for _ in range(2)]
        self.comp_done = [torch.cuda.Event() for _ in range(2)]
        self.out_done = [torch.cuda.Event() for _ in range(2)]
        self.keep = [None, None]          # device results of the chunk in each slot, alive until copied out

    def run(self, q: Tensor, k: Tensor, v: Tensor, dout: Optional[Tensor], o: Tensor,
            grads: Optional[Tuple[Tensor, Tensor, Tensor]], **kw) -> None:
        """q,k,v,dout,o,grads: pinned host tensors shaped (units, rows, D).  Blocks until the outputs are on the host."""
        if kw.get("dropout_p", 0.0) > 0.0 and kw.get("_philox") is None:
            # one (seed, offset) per logical call: with the global unit index (`_bh_offset`) the dropout mask is then the
            # same whatever `chunks` is
            from flash_attention_softmax_n.core.flash_attn import _next_philox
            kw = dict(kw, _philox=_next_philox(self.device))
        cur = torch.cuda.current_stream(self.device)
        for st in (self.h2d, self.comp, self.d2h):
            st.wait_stream(cur)
        for c in range(self.chunks):
            lo, hi = c * self.per, min((c + 1) * self.per, self.units)
            if lo >= hi:
                break
            n, slot = hi - lo, c & 1
            with torch.cuda.stream(self.h2d):
                if c >= 2:
                    self.h2d.wait_event(self.comp_done[slot])      # kernels that read this slot's inputs are done
                self.dq[slot][:n].copy_(q[lo:hi], non_blocking=True)
                self.dk[slot][:n].copy_(k[lo:hi], non_blocking=True)
                self.dv[slot][:n].copy_(v[lo:hi], non_blocking=True)
                if self.backward:
                    self.ddo[slot][:n].copy_(dout[lo:hi], non_blocking=True)
                self.in_done[slot].record(self.h2d)
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(self.in_done[slot])
                if c >= 2:
                    self.comp.wait_event(self.out_done[slot])      # previous results of this slot have left the device
                qd, kd, vd = (t[slot][:n].unsqueeze(0) for t in (self.dq, self.dk, self.dv))
                if self.backward:
                    qd, kd, vd = (t.detach().requires_grad_() for t in (qd, kd, vd))
                od = flash_attention_n(qd, kd, vd, _bh_offset=lo, **kw)
                if self.backward:
                    od.backward(self.ddo[slot][:n].unsqueeze(0))
                    self.keep[slot] = (od.detach(), qd.grad, kd.grad, vd.grad)
                else:
                    self.keep[slot] = (od.detach(),)
                self.comp_done[slot].record(self.comp)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.comp_done[slot])
                res = self.keep[slot]
                o[lo:hi].copy_(res[0][0], non_blocking=True)
                if self.backward:
                    for dst, src in zip(grads, res[1:]):
                        dst[lo:hi].copy_(src[0], non_blocking=True)
                self.out_done[slot].record(self.d2h)
        cur.wait_stream(self.d2h)
        cur.wait_stream(self.comp)
        self.d2h.synchronize()


def attention_host(query: Tensor, key: Tensor, value: Tensor, dout: Optional[Tensor] = None, *, device=None, chunks: int = 8,
                   pipeline: Optional[HostPipeline] = None, **kw):
    """`flash_attention_n` for host tensors of shape (B, H, L|S, D): returns O (and dQ, dK, dV when `dout` is given) as
    pinned host tensors.  Keyword arguments are those of `flash_attention_n` (softmax_n_param, scale, dropout_p,
    is_causal, _philox); masks / biases are not supported on this path."""
    if "attn_mask" in kw or "attn_bias" in kw:
        raise NotImplementedError("attention_host does not stream attn_mask / attn_bias")
    B, H, L, D = query.shape
    S = key.shape[2]
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    pin = lambda t: t if t.is_pinned() else t.pin_memory()
    flat = lambda t, n: pin(t.reshape(B * H, n, D))
    q, k, v = flat(query, L), flat(key, S), flat(value, S)
    do = flat(dout, L) if dout is not None else None
    o = torch.empty(B * H, L, D, dtype=query.dtype).pin_memory()
    grads = tuple(torch.empty(B * H, n, D, dtype=query.dtype).pin_memory() for n in (L, S, S)) if dout is not None else None
    if pipeline is None:
        pipeline = HostPipeline(B * H, L, S, D, query.dtype, device, chunks, backward=dout is not None)
    pipeline.run(q, k, v, do, o, grads, **kw)
    out = o.reshape(B, H, L, D)
    if dout is None:
        return out
    return (out,) + tuple(g.reshape(B, H, n, D) for g, n in zip(grads, (L, S, S)))
