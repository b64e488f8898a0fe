import os, time, torch, torch.distributed as dist
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")   # 1 GiB
for name in ("sendrecv", "allreduce", "batched"):
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
        if name == "sendrecv":
            if rank == 1: dist.send(x, dst=0)
            elif rank == 0: dist.recv(x, src=1)
        elif name == "batched":
            ops = [dist.P2POp(dist.isend, x, 0)] if rank == 1 else [dist.P2POp(dist.irecv, x, 1)]
            for w in dist.batch_isend_irecv(ops): w.wait()
        else:
            dist.all_reduce(x)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    if rank == 0: print(name, f"{x.numel()*4/dt/1e9:.1f} GB/s", flush=True)
if rank == 0:
    print("p2p access", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
