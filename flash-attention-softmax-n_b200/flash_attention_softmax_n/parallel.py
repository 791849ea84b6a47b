"""Batch x head sharding of attention over the GPUs of one node (one process per GPU, torch.distributed).

The reference has no distributed code; this is the single data-parallel step the
north star asks for.  Every (batch, head) pair is an independent attention problem (the reference's grid axis
`off_hz`, flash_attn_triton.py:46,274), so the flattened unit axis is cut into contiguous slabs, one per
rank, with NO collective on the data path of the kernels.  Two ways to use it:

* resident slabs (data-parallel training: each rank already owns its units): call `local_attention`;
* root-held tensors: `sharded_attention` scatters contiguous slabs of Q/K/V from the root, runs the local kernel, and
  gathers O back.  Two transports: "p2p" = point-to-point sends of the process group (NCCL over NVLink/NVSwitch on GPUs,
  gloo on CPU: what the world-size-2 CPU tests run), "ipc" = the root's copy engines write the slabs straight into
  staging buffers of the peers that are mapped into the root process through CUDA IPC handles, and the peers write their
  outputs straight into the root's result tensor; NCCL only carries a 4-byte barrier per pipeline step.  The scatter is
  bound by the root's NVLink egress either way; the copy-engine route leaves the SMs to the
  attention kernels and reaches a higher share of the peer-copy rate.

Dropout masks are keyed by the GLOBAL unit index, so results do not depend on the world size.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def partition_units(n_units: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal [start, stop) slabs of the flattened (batch, head) axis; the first
    `n_units % world_size` ranks get one extra unit."""
    if n_units < 0 or world_size <= 0:
        raise ValueError("n_units must be >= 0 and world_size > 0")
    base, extra = divmod(n_units, world_size)
    out, start = [], 0
    for r in range(world_size):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def scatter_units(full: Optional[Tensor], unit_shape: Tuple[int, ...], dtype: torch.dtype, device: torch.device,
                  n_units: int, root: int = 0, group=None) -> Tensor:
    """Root holds `full` of shape (n_units, *unit_shape) (contiguous); every rank returns its slab."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = partition_units(n_units, world)
    lo, hi = parts[rank]
    if rank == root:
        assert full is not None and full.shape[0] == n_units and full.is_contiguous()
        reqs = []
        for r, (a, b) in enumerate(parts):
            if r != root and b > a:
                reqs.append(dist.isend(full[a:b], dst=dist.get_global_rank(group, r) if group else r, group=group))
        local = full[lo:hi].clone()
        for q in reqs:
            q.wait()
        return local
    local = torch.empty((hi - lo, *unit_shape), dtype=dtype, device=device)
    if hi > lo:
        dist.recv(local, src=dist.get_global_rank(group, root) if group else root, group=group)
    return local


def gather_units(local: Tensor, n_units: int, root: int = 0, group=None) -> Optional[Tensor]:
    """Inverse of `scatter_units`: the root returns (n_units, *unit_shape), other ranks None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = partition_units(n_units, world)
    lo, hi = parts[rank]
    assert local.shape[0] == hi - lo
    if rank == root:
        full = torch.empty((n_units, *local.shape[1:]), dtype=local.dtype, device=local.device)
        reqs = []
        for r, (a, b) in enumerate(parts):
            if r != root and b > a:
                reqs.append(dist.irecv(full[a:b], src=dist.get_global_rank(group, r) if group else r, group=group))
        full[lo:hi].copy_(local)
        for q in reqs:
            q.wait()
        return full
    if hi > lo:
        dist.send(local.contiguous(), dst=dist.get_global_rank(group, root) if group else root, group=group)
    return None


def local_attention(q_units: Tensor, k_units: Tensor, v_units: Tensor, unit_offset: int, *,
                    attn_fn: Optional[Callable] = None, **kwargs) -> Tensor:
    """Attention over a resident slab.  `*_units` are (U, S, D): U independent (batch, head) units whose
    global indices start at `unit_offset`.  Returns (U, L, D).
    With dropout, pass the same `_philox=(seed, offset)` on every rank (and for every chunk): the keep mask is a function of
    (seed, offset, global unit index, row, column), so only then is it independent of how the units are partitioned;
    without it every call draws its own stream from the local generator.  `sharded_attention` does this for you."""
    if attn_fn is None:
        from flash_attention_softmax_n.core.flash_attn import flash_attention_n as attn_fn
        kwargs = dict(kwargs, _bh_offset=unit_offset)
    if q_units.shape[0] == 0:
        return q_units.new_empty(q_units.shape)
    out = attn_fn(q_units.unsqueeze(0), k_units.unsqueeze(0), v_units.unsqueeze(0), **kwargs)
    return out.squeeze(0)


def _peer_copy(dst: Tensor, src: Tensor, stream: "torch.cuda.Stream") -> None:
    """dst <- src on `stream` through the C ABI's fasn_copy_async (cudaMemcpyAsync): either tensor may live in another
    process's device memory (IPC-mapped).  Both must be contiguous and of equal size."""
    from flash_attention_softmax_n import _native
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() * dst.element_size() == src.numel() * src.element_size()
    _native.check(_native.load().fasn_copy_async(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(), stream.cuda_stream),
                  "fasn_copy_async")


class IpcSlabs:
    """Device buffers of one (shape, chunks) problem, mapped across the processes of the group with CUDA IPC handles.

    Every non-root rank owns two staging slots for a (q, k, v) piece; the root owns the (n_units, L, D) result.  After
    `__init__` (collective: handles are exchanged with all_gather_object) the root holds views of every peer's slots and
    every peer holds a view of the root's result, all living in the OTHER process's device memory: a `copy_` between such
    a view and a local tensor is one peer-to-peer cudaMemcpyAsync over NVLink, executed by a copy engine."""

    def __init__(self, shape: Tuple[int, int, int, int, int], dtype: torch.dtype, device: torch.device, chunks: int,
                 root: int = 0, group=None):
        from torch.multiprocessing.reductions import reduce_tensor
        B, H, L, S, D = shape
        self.shape, self.dtype, self.device, self.chunks, self.root, self.group = shape, dtype, device, chunks, root, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n_units = B * H
        self.parts = partition_units(n_units, self.world)
        self.sub = [partition_units(b - a, chunks) for a, b in self.parts]
        per = max((hi - lo) for pr in self.sub for lo, hi in pr) if n_units else 0
        on_root = self.rank == root
        self.slots = None
        self.out_full = torch.empty((n_units, L, D), dtype=dtype, device=device) if on_root else None
        if not on_root:
            self.slots = [tuple(torch.empty((per, n, D), dtype=dtype, device=device) for n in (L, S, S)) for _ in range(2)]
        mine = reduce_tensor(self.out_full) if on_root else [[reduce_tensor(t) for t in slot] for slot in self.slots]
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=group)
        self.peer_slots = {}            # root: rank -> [slot][q,k,v] views into that rank's staging memory
        self.root_out = None            # peers: view of the root's result tensor
        if on_root:
            for r in range(self.world):
                if r != root:
                    self.peer_slots[r] = [[fn(*args) for fn, args in slot] for slot in allh[r]]
        else:
            fn, args = allh[root]
            self.root_out = fn(*args)
        self.copy_streams = [torch.cuda.Stream(device) for _ in range(max(1, self.world - 1) if on_root else 1)]
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)

    def piece(self, r: int, c: int) -> Tuple[int, int]:
        return self.parts[r][0] + self.sub[r][c][0], self.parts[r][0] + self.sub[r][c][1]

    def step_barrier(self) -> None:
        """Everything every rank has issued so far (copies included) is complete before anything issued after it starts."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.copy_streams:
            cur.wait_stream(st)
        dist.all_reduce(self.flag, group=self.group)
        for st in self.copy_streams:
            st.wait_stream(cur)


def _sharded_attention_ipc(qf, kf, vf, slabs: "IpcSlabs", attn_fn, kwargs) -> Optional[Tensor]:
    """Pipeline of `chunks` pieces per rank: in step s the root's copy engines write piece s into slot s % 2 of every peer
    while all ranks run the kernels on piece s - 1 and the peers' copy engines write the output of piece s - 1 into the
    root's result.  A 4-byte all-reduce separates the steps."""
    B, H, L, S, D = slabs.shape
    rank, root, chunks = slabs.rank, slabs.root, slabs.chunks
    on_root = rank == root
    cur = torch.cuda.current_stream(slabs.device)
    for st in slabs.copy_streams:
        st.wait_stream(cur)
    keep = []
    for step in range(chunks + 1):
        if step < chunks and on_root:                       # scatter piece `step`: one copy stream per peer
            i = 0
            for r in range(slabs.world):
                if r == root:
                    continue
                a, b = slabs.piece(r, step)
                if b > a:
                    for src, dst in zip((qf, kf, vf), slabs.peer_slots[r][step & 1]):
                        _peer_copy(dst[: b - a], src[a:b], slabs.copy_streams[i])
                i += 1
        c = step - 1                                        # compute piece c (it landed during the previous step)
        if 0 <= c < chunks:
            a, b = slabs.piece(rank, c)
            if b > a:
                if on_root:
                    qc, kc, vc = qf[a:b], kf[a:b], vf[a:b]       # views: the root's own slab is never copied
                else:
                    qc, kc, vc = (t[: b - a] for t in slabs.slots[c & 1])
                oc = local_attention(qc, kc, vc, a, attn_fn=attn_fn, **kwargs)
                if on_root:
                    slabs.out_full[a:b].copy_(oc)
                else:
                    st = slabs.copy_streams[0]
                    st.wait_stream(cur)
                    _peer_copy(slabs.root_out[a:b], oc.contiguous(), st)      # peer write into the root's memory
                    keep.append(oc)                          # alive until the final barrier
        slabs.step_barrier()
    keep.clear()
    return slabs.out_full.reshape(B, H, L, D) if on_root else None


def sharded_attention(query: Optional[Tensor], key: Optional[Tensor], value: Optional[Tensor], *,
                      shape: Tuple[int, int, int, int, int], dtype: torch.dtype, device: torch.device,
                      root: int = 0, group=None, attn_fn: Optional[Callable] = None, chunks: int = 1,
                      transport: str = "p2p", slabs: Optional["IpcSlabs"] = None, **kwargs) -> Optional[Tensor]:
    """Scatter (B,H,L,D)/(B,H,S,D) tensors held by `root` over the group by (batch, head) slabs, run attention
    on every rank, gather O on the root.  `shape` = (B, H, L, S, D) must be passed on every rank.
    Returns (B,H,L,D) on the root and None elsewhere.  attn_mask / attn_bias are not sharded here.

    `transport` = "ipc" (CUDA only) moves the slabs with copy engines through IPC-mapped peer memory (see `IpcSlabs`; pass a
    prebuilt `slabs` to reuse the mapped buffers across calls); "p2p" uses the group's point-to-point sends.
    With dropout and no `_philox`, the root draws one (seed, offset) and broadcasts it, so the mask does not depend on
    the world size or on `chunks`.

    `chunks` > 1 cuts every rank's slab into that many pieces and software-pipelines them: while the kernels work on
    piece c, piece c+1 is in flight from the root and the output of piece c-1 is in flight back (one batched
    point-to-point group per step, NCCL on its own stream), so a step costs about max(transfer, kernels) instead of
    their sum.  The root's own slab is never copied."""
    if "attn_mask" in kwargs or "attn_bias" in kwargs:
        raise NotImplementedError("sharded_attention does not distribute attn_mask / attn_bias")
    if transport not in ("p2p", "ipc"):
        raise ValueError("transport must be 'p2p' or 'ipc'")
    B, H, L, S, D = shape
    n_units = B * H
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_root = rank == root
    if kwargs.get("dropout_p", 0.0) > 0.0 and kwargs.get("_philox") is None and attn_fn is None:
        # one (seed, offset) for the whole logical call, drawn on the root: with the global unit index it makes the dropout
        # mask independent of the world size and of the chunking
        from flash_attention_softmax_n.core.flash_attn import _next_philox
        box = [_next_philox(device) if on_root else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, root) if group else root, group=group)
        kwargs = dict(kwargs, _philox=tuple(box[0]))
    qf = query.reshape(n_units, L, D) if on_root else None
    kf = key.reshape(n_units, S, D) if on_root else None
    vf = value.reshape(n_units, S, D) if on_root else None
    if on_root:
        qf, kf, vf = qf.contiguous(), kf.contiguous(), vf.contiguous()
    if transport == "ipc":
        if device.type != "cuda":
            raise NotImplementedError("transport='ipc' maps CUDA device memory between processes; use 'p2p' on CPU")
        if slabs is None:
            slabs = IpcSlabs(shape, dtype, device, max(1, chunks), root, group)
        return _sharded_attention_ipc(qf, kf, vf, slabs, attn_fn, kwargs)
    if chunks <= 1:
        ql = scatter_units(qf, (L, D), dtype, device, n_units, root, group)
        kl = scatter_units(kf, (S, D), dtype, device, n_units, root, group)
        vl = scatter_units(vf, (S, D), dtype, device, n_units, root, group)
        lo, _ = partition_units(n_units, world)[rank]
        ol = local_attention(ql, kl, vl, lo, attn_fn=attn_fn, **kwargs)
        full = gather_units(ol.contiguous(), n_units, root, group)
        return full.reshape(B, H, L, D) if on_root else None

    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    parts = partition_units(n_units, world)
    # piece c of rank r: global units [parts[r][0] + sub[r][c][0], parts[r][0] + sub[r][c][1])
    sub = [partition_units(b - a, chunks) for a, b in parts]
    piece = lambda r, c: (parts[r][0] + sub[r][c][0], parts[r][0] + sub[r][c][1])
    out_full = torch.empty((n_units, L, D), dtype=dtype, device=device) if on_root else None
    inbuf: List[Optional[Tuple[Tensor, Tensor, Tensor]]] = [None] * chunks
    outbuf: List[Optional[Tensor]] = [None] * chunks
    reqs_in: List[list] = [[] for _ in range(chunks)]
    reqs_out: list = []
    waited = set()

    for step in range(chunks + 2):
        ops = []
        c_in, c_out = step, step - 2
        if c_in < chunks:                                   # scatter piece c_in
            if on_root:
                for r in range(world):
                    a, b = piece(r, c_in)
                    if r != root and b > a:
                        ops += [dist.P2POp(dist.isend, t[a:b], peer(r), group) for t in (qf, kf, vf)]
                a, b = piece(root, c_in)
                inbuf[c_in] = (qf[a:b], kf[a:b], vf[a:b])     # views: the root's own slab is not copied
            else:
                a, b = piece(rank, c_in)
                bufs = tuple(torch.empty((b - a, n, D), dtype=dtype, device=device) for n in (L, S, S))
                inbuf[c_in] = bufs
                if b > a:
                    ops += [dist.P2POp(dist.irecv, t, peer(root), group) for t in bufs]
        if 0 <= c_out < chunks:                             # gather the output of piece c_out
            if on_root:
                for r in range(world):
                    a, b = piece(r, c_out)
                    if r != root and b > a:
                        ops.append(dist.P2POp(dist.irecv, out_full[a:b], peer(r), group))
            else:
                a, b = piece(rank, c_out)
                if b > a:
                    ops.append(dist.P2POp(dist.isend, outbuf[c_out], peer(root), group))
        posted = dist.batch_isend_irecv(ops) if ops else []
        if c_in < chunks:
            reqs_in[c_in] = posted                          # (a group completes as a whole: its gather ops ride along)
        reqs_out += posted
        c = step - 1                                        # compute piece c while the group above is in flight
        if 0 <= c < chunks:
            for w in reqs_in[c]:
                w.wait()
                waited.add(id(w))
            a, b = piece(rank, c)
            qc, kc, vc = inbuf[c]
            oc = local_attention(qc, kc, vc, a, attn_fn=attn_fn, **kwargs).contiguous()
            if on_root:
                out_full[a:b].copy_(oc)
            else:
                outbuf[c] = oc
            inbuf[c] = None
    for w in reqs_out:
        if id(w) not in waited:       # (a second wait() on a completed gloo work blocks)
            w.wait()
    return out_full.reshape(B, H, L, D) if on_root else None
