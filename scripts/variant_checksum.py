"""Bitwise fingerprint of the pipeline's outputs for an A/B of two builds (FCD_B200_LIB selects the library):
heights, phases and the auto-mode bookkeeping of a seeded batch that holds clean, wrapping and noisy frames.  Two
builds whose kernels differ only in schedule / structure must print the same digests."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import numpy as np, torch
from fcd_b200 import HeightMapPlan, synthetic as sy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
ref = sy.rotated_board(n)
clean = sy.synthetic_frames(n, 5, seed=11, peak_range=(0.2, 0.8))[1]
wrap = sy.synthetic_frames(n, 4, seed=12, peak_range=(12.0, 25.0))[1]
frames = np.concatenate([clean, wrap, clean[:2]])
rng = np.random.default_rng(3)
frames[-1] += (0.25 * rng.standard_normal((n, n))).astype(np.float32)      # residues: goes the guided way in auto
plan = HeightMapPlan((n, n), 4, dev)                                        # 4 frames per wave: ragged last wave
plan.bind(ref, square_size=sy.board_square_size(n), height=1.0)
d = torch.from_numpy(frames).to(dev)
out = {"size": n, "lib": os.path.basename(os.environ.get("FCD_B200_LIB", "libfcd_b200.so"))}
for mode in ("scan", "off", "auto"):
    ph = torch.empty((len(frames), 2, n, n), dtype=torch.float32, device=dev)
    h, _ = plan.execute(d, unwrap=mode, phases=ph)
    torch.cuda.synchronize()
    out[mode] = {"heights": hashlib.sha256(h.cpu().numpy().tobytes()).hexdigest()[:16],
                 "phases": hashlib.sha256(ph.cpu().numpy().tobytes()).hexdigest()[:16],
                 "max_abs_phase": float(ph.abs().max())}
    if mode == "auto":
        out[mode]["flagged"] = int(plan.last_flagged_frames); out[mode]["guided"] = list(plan.last_guided_frames)
print(json.dumps(out))
