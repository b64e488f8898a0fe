"""Second probe: a peer's buffer reached through torch symmetric memory (cuMem / VMM handles, the transport NCCL
reports as P2P/CUMEM) filled with a plain cudaMemcpyAsync from the local GPU."""
import os, sys, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 1 << 29
x = torch.full((N,), float(rank + 1), dtype=torch.float32, device=dev)
t = symm_mem.empty(N, dtype=torch.float32, device=dev)
t.zero_()
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier()
if rank == 1:
    view = hdl.get_buffer(0, (N,), torch.float32)
    print("peer view device", view.device, flush=True)
    s = torch.cuda.Stream()
    def push():
        with torch.cuda.stream(s):
            view.copy_(x, non_blocking=True)
        s.synchronize()
    push()
    t0 = time.perf_counter()
    for _ in range(5):
        push()
    print(f"D symmetric-memory peer copy: {N * 4 * 5 / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    print("D root sees", float(t[:10].sum()), float(t[-10:].sum()), flush=True)
dist.barrier()
dist.destroy_process_group()
