"""Times the reliability-guided unwrap (fcd_unwrap_phase) and the guided pipeline mode at 2048^2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import torch
import fcd_b200
from bench import make_frames_gpu, SEED
from fcd_b200 import synthetic as o

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
plan = fcd_b200.HeightMapPlan((n, n), 16, dev)
ref, frames = make_frames_gpu(n, 16, SEED, dev, peak_range=(4.0, 5.0))
plan.bind(ref, square_size=o.board_square_size(n), height=1.0)
noisy = frames + 0.25 * torch.randn_like(frames)
for name, fr in (("clean", frames), ("noisy", noisy)):
    _, w = plan.execute(fr, phases=True, unwrap=False)
    res = plan.count_residues(w)
    for mode in ("scan", "herraez", "auto"):
        plan.execute(fr, unwrap=mode); torch.cuda.synchronize()
        r0 = plan.launch_count
        t = time.perf_counter(); plan.execute(fr, unwrap=mode); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{name} {mode:8s} {1e3 * dt / fr.shape[0]:8.2f} ms/frame   launches {plan.launch_count - r0}  residues/map {sum(res) / len(res):.0f}")
    torch.cuda.synchronize(); t = time.perf_counter(); plan.unwrap_phase(w); torch.cuda.synchronize()
    print(f"{name} unwrap_phase alone {1e3 * (time.perf_counter() - t) / (2 * fr.shape[0]):.2f} ms/map")
