"""BASELINE.json configs[3]: a streaming "video" of 2048^2 frames sharded by frame across the GPUs of one node,
with the final gather of the height maps onto rank 0 over NCCL.

    torchrun --nproc-per-node N scripts/stream_video.py [frames_total] [chunk]

Every rank processes its contiguous frame range (fcd_b200.shard_range) chunk by chunk; there is no collective on
the hot path.  Two ways of gathering are timed: (a) streamed -- while a rank computes chunk k+1 it sends the maps
of chunk k to rank 0 with NCCL point-to-point, which receives them into a two-slot ring (20k maps = 335 GB do not
fit one GPU; a consumer would drain the ring to host or disk); NCCL's copy kernels then compete for SMs with the
persistent FCD kernels, which is why (b) is the recommended order -- (b) compute everything into the rank's own
HBM first (2,500 maps per GPU at 8 GPUs = 42 GB), then gather in a separate phase at NVLink speed.  Also checks
the library's chunked gather (`gather_height_maps`) bit for bit against rank-local results."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import torch
import torch.distributed as dist
import fcd_b200
from fcd_b200.engine import shard_range, gather_height_maps
from bench import make_frames_gpu, SEED
from fcd_b200 import synthetic as o

total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
n = 2048
a, b = shard_range(total, rank, ws)
plan = fcd_b200.HeightMapPlan((n, n), chunk, dev)
ref, frames = make_frames_gpu(n, chunk, SEED + rank, dev)          # one synthetic chunk per rank, replayed
plan.bind(ref, square_size=o.board_square_size(n), height=1.0)
outs = [torch.empty((chunk, n, n), dtype=torch.float32, device=dev) for _ in range(2)]
ring = [torch.empty((chunk, n, n), dtype=torch.float32, device=dev) for _ in range(2)] if rank == 0 else None


def run(gather):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    pending = []
    nchunks = (b - a + chunk - 1) // chunk
    for k in range(nchunks):
        cnt = min(chunk, b - a - k * chunk)
        buf = outs[k & 1]
        for w in pending[:-1]:
            w.wait()
        pending = pending[-1:]
        plan.execute(frames[:cnt], out=buf[:cnt])
        if gather:
            if rank != 0:
                pending.append(dist.isend(buf[:cnt], dst=0))
            else:
                for src in range(1, ws):
                    sa, sb = shard_range(total, src, ws)
                    scnt = min(chunk, sb - sa - k * chunk)
                    if scnt > 0:
                        pending.append(dist.irecv(ring[k & 1][:scnt], src=src))
    for w in pending:
        w.wait()
    torch.cuda.synchronize(); dist.barrier()
    return time.perf_counter() - t0


run(False)
t_plain = run(False)
t_gather = run(True)
# (b) separate gather phase: one chunk per non-root rank per step into the ring, no compute running
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
nchunks = (b - a + chunk - 1) // chunk
for k in range(nchunks):
    ops = []
    if rank != 0:
        ops.append(dist.P2POp(dist.isend, outs[k & 1], 0))
    else:
        for src in range(1, ws):
            ops.append(dist.P2POp(dist.irecv, ring[k & 1], src))      # a real consumer drains ring[k & 1] here
    for w in dist.batch_isend_irecv(ops):
        w.wait()
torch.cuda.synchronize(); dist.barrier()
t_phase = time.perf_counter() - t0
# small-case check of the chunked gather used by the library API
small = plan.execute(frames[:6])
lo, hi = shard_range(6 * ws, rank, ws)
full = gather_height_maps(small, 6 * ws, dst=0, chunk_frames=4)
ok = True
if rank == 0:
    ok = bool(torch.equal(full[lo:hi], small)) and full.shape[0] == 6 * ws
    print(json.dumps({"world": ws, "frames_total": total, "chunk": chunk,
                      "frames_per_s_no_gather": total / t_plain, "frames_per_s_with_streamed_gather": total / t_gather,
                      "streamed_gather_gbs_into_rank0": (total - (b - a)) * n * n * 4 / t_gather / 1e9,
                      "separate_gather_phase_s": t_phase,
                      "separate_gather_gbs_into_rank0": nchunks * (ws - 1) * chunk * n * n * 4 / t_phase / 1e9,
                      "frames_per_s_compute_then_gather": total / (t_plain + t_phase),
                      "small_gather_bit_exact": ok}))
dist.destroy_process_group()
