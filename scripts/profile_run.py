"""Tiny driver for ncu: a few 2048^2 frames through the pipeline (two warm chunks, then one more)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import torch
from bench import make_frames_gpu, SEED
from fcd_b200 import HeightMapPlan
from fcd_b200 import synthetic as o

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
fpl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
plan = HeightMapPlan((n, n), fpl, dev)
ref, frames = make_frames_gpu(n, 3 * fpl, SEED, dev)
plan.bind(ref, square_size=o.board_square_size(n), height=1.0)
out = plan.execute(frames)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
