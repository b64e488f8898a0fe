"""Times the f0 estimate (mean spectra of all 64 blocks) and the harmonic sums on a LONG series: 20,000 maps of
512 x 512 (config-4 frame count; 20,000 = 160 x 125 takes the two-level transform)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import numpy as np
import torch
from fcd_b200 import temporal as tp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
t = torch.arange(n, device=dev, dtype=torch.float32)
cyc = 1234
maps = torch.empty((n, size, size), device=dev)
amp = 1 + torch.rand((1, size, size), device=dev)
for c0 in range(0, n, 500):
    c1 = min(n, c0 + 500)
    maps[c0:c1] = torch.cos(2 * np.pi * cyc * t[c0:c1] / n)[:, None, None] * amp + 0.05 * torch.randn((c1 - c0, size, size), device=dev)
tp.block_amplitudes(maps[:64], mode=3, num_blocks=64); torch.cuda.synchronize()
out = {"frames": n, "size": size, "stack_gb": maps.numel() * 4 / 1e9}
t0 = time.perf_counter(); res = tp.block_amplitudes(maps, mode=3, num_blocks=64, tasa=500); torch.cuda.synchronize()
out["f0_estimated_s"] = time.perf_counter() - t0
out["f0_ok"] = bool(all(abs(f - cyc * 500 / n) < 1e-6 for f in res.f0))
t0 = time.perf_counter(); tp.block_amplitudes(maps, f0=cyc * 500 / n, mode=3, num_blocks=64, tasa=500); torch.cuda.synchronize()
out["f0_given_s"] = time.perf_counter() - t0
# the reference algorithm on one block (numpy FFT of every pixel's series), extrapolated to 64 blocks
bs = size // 8
host = maps[:, :bs, :bs].cpu().numpy()
t0 = time.perf_counter()
spec = np.abs(np.fft.fft(np.transpose(host, (1, 2, 0)), axis=-1)).mean(axis=(0, 1))
out["cpu_numpy_fft_one_block_s"] = time.perf_counter() - t0
out["cpu_64_blocks_extrapolated_s"] = 64 * out["cpu_numpy_fft_one_block_s"]
print(json.dumps(out))
