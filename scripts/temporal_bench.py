"""Times the temporal harmonic analysis (analyze.block_amplitude for all 64 blocks) on a
device-resident stack, next to the reference algorithm (oracle port) on one block on the CPU."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import numpy as np
import torch
from fcd_b200 import temporal as tp
from oracle import temporal_oracle as to

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
t = torch.arange(n, device=dev, dtype=torch.float32) / 500.0
maps = torch.cos(2 * np.pi * 37.109375 * t)[:, None, None] * torch.rand((1, size, size), device=dev)
for c0 in range(0, n, 64):
    maps[c0:c0 + 64] += 0.05 * torch.randn((min(64, n - c0), size, size), device=dev)
plan = tp.get_plan((64, 64), 1, dev)
out = {"size": size, "frames": n, "stack_gb": maps.numel() * 4 / 1e9}
for name, kw in (("f0_estimated", {}), ("f0_given", {"f0": 37.109375})):
    tp.block_amplitudes(maps[:64], mode=3, num_blocks=64, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); res = tp.block_amplitudes(maps, mode=3, num_blocks=64, **kw); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = {"seconds": dt, "stack_reads_gbs": out["stack_gb"] * (2 if not kw else 1) / dt,
                 "f0_block0": float(res.f0[0])}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
first = maps[0].clone()
e0.record(); tp.mean_spectra(maps, first, 0.0, size // 8, 8, 8, plan); e1.record(); torch.cuda.synchronize()
out["mean_spectrum_kernel"] = {"ms": e0.elapsed_time(e1), "gbs": out["stack_gb"] / e0.elapsed_time(e1) * 1e3}
# CPU: the reference algorithm on ONE of the 64 blocks (it would be called 64 times, re-reading the files each time)
host = maps[:, :size // 8, :size // 8].cpu().numpy()
t0 = time.perf_counter(); to.block_amplitude(host, mode=3, num_blocks=1, block_index=0); dt = time.perf_counter() - t0
out["cpu_reference_algorithm"] = {"seconds_one_block": dt, "seconds_64_blocks_extrapolated": 64 * dt,
                                  "note": "in-memory stack, no file reads (the reference re-reads every map per block)"}
print(json.dumps(out))
