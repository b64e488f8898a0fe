"""Probe of the device-to-device paths available for the streamed gather (2 ranks): in-process peer copy, a CUDA IPC
mapping filled by the peer process, and NCCL send/recv (optionally with NCCL_P2P_USE_CUDA_MEMCPY=1)."""
import os, sys, time, torch, torch.distributed as dist
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 1 << 29   # 2 GiB of float32
x = torch.ones(N, dtype=torch.float32, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return N * 4 * reps / (time.perf_counter() - t) / 1e9


if rank == 1:   # A: one process, two devices
    y0 = torch.empty(N, dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        y0.copy_(x, non_blocking=True)
    torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    print(f"A in-process peer copy: {N * 4 * 5 / (time.perf_counter() - t) / 1e9:.1f} GB/s, p2p {torch.cuda.can_device_access_peer(1, 0)}", flush=True)
    del y0
dist.barrier()
# B: IPC mapping of rank 0's buffer in rank 1
from torch.multiprocessing.reductions import reduce_tensor
h = [None]
if rank == 0:
    ring = torch.zeros(N, dtype=torch.float32, device=dev)
    h = [reduce_tensor(ring)]
dist.broadcast_object_list(h, src=0)
if rank == 1:
    fn, args = h[0]
    view = fn(*args)
    print("view device", view.device, flush=True)
    s = torch.cuda.Stream()
    def push():
        with torch.cuda.stream(s):
            view.copy_(x, non_blocking=True)
        s.synchronize()
    push(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        push()
    print(f"B IPC view copy: {N * 4 * 5 / (time.perf_counter() - t) / 1e9:.1f} GB/s", flush=True)
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    print("B root sees", float(ring[:10].sum()), float(ring[-10:].sum()), flush=True)
# C: NCCL send/recv
def nccl():
    ops = [dist.P2POp(dist.isend, x, 0)] if rank == 1 else [dist.P2POp(dist.irecv, x, 1)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
g = timed(nccl)
if rank == 0:
    print(f"C NCCL send/recv (NCCL_P2P_USE_CUDA_MEMCPY={os.environ.get('NCCL_P2P_USE_CUDA_MEMCPY')}): {g:.1f} GB/s", flush=True)
dist.destroy_process_group()
