"""torchrun --nproc-per-node N scripts/temporal_sharded_check.py: frame-sharded temporal analysis over NCCL
(all-reduce of the harmonic sums, all-to-all frame shards -> row bands for the f0 estimate) against the same
analysis of the whole stack on one GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import numpy as np
import torch
import torch.distributed as dist
from fcd_b200 import temporal as tp
from fcd_b200.engine import shard_range
from oracle import temporal_oracle as to

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, size, nb = 200, 512, 64
maps = torch.from_numpy(to.synthetic_maps(n, (size, size), nb, seed=9)).cuda()
a, b = shard_range(n, rank, ws)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
res = tp.block_amplitudes(maps[a:b].contiguous(), mode=3, num_blocks=nb, group=dist.group.WORLD, n_total=n)
torch.cuda.synchronize(); dist.barrier()
ev[0].record()
res = tp.block_amplitudes(maps[a:b].contiguous(), mode=3, num_blocks=nb, group=dist.group.WORLD, n_total=n)
ev[1].record(); torch.cuda.synchronize()
full = tp.block_amplitudes(maps, mode=3, num_blocks=nb)
same_f0 = res.f0 == full.f0
z = lambda r: torch.nan_to_num(r.amps * torch.exp(1j * r.phases))
err = float((z(res) - z(full)).abs().max() / z(full).abs().max())
nan_same = bool(torch.equal(torch.isnan(res.amps), torch.isnan(full.amps)))
ok = torch.tensor([int(same_f0 and nan_same and err < 1e-12)], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": ws, "frames": n, "size": size, "blocks": nb, "f0_identical": same_f0, "nan_identical": nan_same,
                      "max_rel_err_vs_single_gpu": err, "all_ranks_ok": bool(ok.item()), "sharded_ms": ev[0].elapsed_time(ev[1])}))
dist.destroy_process_group()
