"""Tiny driver for ncu / timing of the reliability-guided unwrap: noisy wrapping 2048^2 maps (residues everywhere)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import torch
from fcd_b200 import HeightMapPlan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
maps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
plan = HeightMapPlan((n, n), 1, dev)
g = torch.Generator(device="cuda").manual_seed(3)
y = torch.arange(n, device="cuda", dtype=torch.float32)[:, None]
x = torch.arange(n, device="cuda", dtype=torch.float32)[None, :]
smooth = 40.0 * torch.exp(-((y - n / 2) ** 2 + (x - n / 2.5) ** 2) / (2 * (n / 5) ** 2)) + 0.01 * x
ph = smooth[None] + 0.7 * torch.randn((maps, n, n), device="cuda", generator=g)
w = torch.atan2(torch.sin(ph), torch.cos(ph)).contiguous()
u = plan.unwrap_phase(w)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    u = plan.unwrap_phase(w)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / max(reps, 1)
print(f"unwrap {maps} maps {n}x{n}: {dt * 1e3:.2f} ms = {dt * 1e3 / maps:.3f} ms per map; residues of map 0: {plan.count_residues(w[:1])}")
