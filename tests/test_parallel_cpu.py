"""Batch x head sharding logic (flash_attention_softmax_n/parallel.py) on CPU: world_size-2 gloo processes.
The attention function is injected (the package's eager slow_attention_n) because the fused kernel needs a GPU;
what is under test is the partition / scatter / gather plumbing and its world-size invariance."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-softmax-n_b200")


def test_partition_units():
    from flash_attention_softmax_n.parallel import partition_units
    assert partition_units(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert partition_units(2560, 8) == [(i * 320, (i + 1) * 320) for i in range(8)]      # BASELINE config 4
    assert partition_units(1, 2) == [(0, 1), (1, 1)]
    for n, w in [(0, 3), (7, 7), (5, 8), (128, 3)]:
        parts = partition_units(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition_units(4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, H, L, S, D, ret):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flash_attention_softmax_n import slow_attention_n
        from flash_attention_softmax_n.parallel import sharded_attention, scatter_units, gather_units
        torch.manual_seed(7)
        q = torch.randn(B, H, L, D, dtype=torch.float64)
        k = torch.randn(B, H, S, D, dtype=torch.float64)
        v = torch.randn(B, H, S, D, dtype=torch.float64)
        kw = dict(softmax_n_param=0.5, is_causal=True)
        out = sharded_attention(q if rank == 0 else None, k if rank == 0 else None, v if rank == 0 else None,
                                shape=(B, H, L, S, D), dtype=torch.float64, device=torch.device("cpu"),
                                attn_fn=slow_attention_n, **kw)
        out3 = sharded_attention(q if rank == 0 else None, k if rank == 0 else None, v if rank == 0 else None,
                                 shape=(B, H, L, S, D), dtype=torch.float64, device=torch.device("cpu"),
                                 attn_fn=slow_attention_n, chunks=3, **kw)      # software-pipelined pieces
        # scatter followed by gather is the identity
        flat = q.reshape(B * H, L, D).contiguous()
        loc = scatter_units(flat if rank == 0 else None, (L, D), torch.float64, torch.device("cpu"), B * H)
        back = gather_units(loc, B * H)
        if rank == 0:
            ref = slow_attention_n(q, k, v, **kw)
            ret["err"] = (out - ref).abs().max().item()
            ret["roundtrip"] = bool(torch.equal(back, flat))
            ret["err_chunked"] = (out3 - ref).abs().max().item()
        else:
            assert out is None and back is None and out3 is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,H", [(2, 3), (1, 1), (1, 5)])
def test_sharded_attention_world2_gloo(B, H):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), B, H, 12, 20, 8, ret), nprocs=2, join=True)
    assert ret["roundtrip"] is True
    assert ret["err"] < 1e-12
    assert ret["err_chunked"] < 1e-12
