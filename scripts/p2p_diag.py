"""Peer-copy diagnostics (2+ GPUs): copy rates between two devices of one process, and from rank 0 into IPC-mapped memory of rank 1."""
import os, sys, time
import torch, torch.distributed as dist

def rate(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    for d in range(torch.cuda.device_count()): torch.cuda.synchronize(d)
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

if "RANK" not in os.environ:
    print("can_access_peer 0->1", torch.cuda.can_device_access_peer(0, 1))
    n = 256 << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda:0"); b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    print("single process copy_ 0->1: %.1f GB/s" % rate(lambda: b.copy_(a, non_blocking=True), n))
    print("single process copy_ 1->0: %.1f GB/s" % rate(lambda: a.copy_(b, non_blocking=True), n))
    sys.exit(0)

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from torch.multiprocessing.reductions import reduce_tensor
n = 256 << 20
buf = torch.empty(n, dtype=torch.uint8, device=dev)
hs = [None] * world
dist.all_gather_object(hs, reduce_tensor(buf))
if rank == 0:
    fn, args = hs[1]
    peer = fn(*args)
    print("peer tensor device", peer.device, "ptr", hex(peer.data_ptr()))
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    print("IPC push copy_ 0->1 (default streams): %.1f GB/s" % rate(lambda: peer.copy_(src, non_blocking=True), n), flush=True)
    st = torch.cuda.Stream(dev)
    def f():
        with torch.cuda.stream(st): peer.copy_(src, non_blocking=True)
    print("IPC push copy_ 0->1 (side stream):     %.1f GB/s" % rate(f, n), flush=True)
    print("IPC pull copy_ 1->0:                  %.1f GB/s" % rate(lambda: src.copy_(peer, non_blocking=True), n), flush=True)
    # raw runtime call through ctypes: cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, stream)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    def g(): assert rt.cudaMemcpyAsync(peer.data_ptr(), src.data_ptr(), n, 4, torch.cuda.current_stream().cuda_stream) == 0
    print("IPC push raw cudaMemcpyAsync:          %.1f GB/s" % rate(g, n), flush=True)
    def h(): assert rt.cudaMemcpyPeerAsync(ctypes.c_void_p(peer.data_ptr()), 1, ctypes.c_void_p(src.data_ptr()), 0, ctypes.c_size_t(n), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    print("IPC push raw cudaMemcpyPeerAsync:      %.1f GB/s" % rate(h, n), flush=True)
dist.barrier()
dist.destroy_process_group()
