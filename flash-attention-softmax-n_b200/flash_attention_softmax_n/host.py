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
