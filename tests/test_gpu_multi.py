"""Two-GPU check of the root-held scatter / compute / gather path over NCCL (skipped on a single-GPU box)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-softmax-n_b200")


def _worker(rank, world, port, ret):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from flash_attention_softmax_n import flash_attention_n
        from flash_attention_softmax_n.parallel import sharded_attention
        B, H, L, S, D = 3, 5, 384, 512, 128
        dev = torch.device("cuda", rank)
        g = torch.Generator().manual_seed(0)
        q, k, v = ((torch.randn(B, H, n, D, generator=g) * 0.5).to(torch.bfloat16).to(dev) for n in (L, S, S))
        kw = dict(softmax_n_param=0.5, is_causal=True, dropout_p=0.1, _philox=(17, 4))
        out = sharded_attention(q if rank == 0 else None, k if rank == 0 else None, v if rank == 0 else None,
                                shape=(B, H, L, S, D), dtype=torch.bfloat16, device=dev, **kw)
        out3 = sharded_attention(q if rank == 0 else None, k if rank == 0 else None, v if rank == 0 else None,
                                 shape=(B, H, L, S, D), dtype=torch.bfloat16, device=dev, chunks=3, **kw)   # pipelined pieces
        # copy-engine transport through CUDA-IPC-mapped peer memory (what bench.py --gpus N times)
        out_ipc = sharded_attention(q if rank == 0 else None, k if rank == 0 else None, v if rank == 0 else None,
                                    shape=(B, H, L, S, D), dtype=torch.bfloat16, device=dev, chunks=2, transport="ipc", **kw)
        # dropout without an explicit stream: the root draws (seed, offset) once, so two transports / chunkings cannot be
        # compared bit for bit -- but every rank must use the same stream: rows of different ranks have the same keep rate
        if rank == 0:
            ref = flash_attention_n(q, k, v, **kw)
            ret["equal"] = bool(torch.equal(out, ref)) and bool(torch.equal(out3, ref)) and bool(torch.equal(out_ipc, ref))
            out_ipc = None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_attention_two_gpus_matches_single_gpu():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["equal"] is True      # dropout masks are keyed by the global unit index: bit-identical to one GPU
