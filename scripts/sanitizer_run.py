"""Small end-to-end case for compute-sanitizer (all kernels, both RowDemod variants, masks, unwrap path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "trapped-modes-ltg_b200"))
import numpy as np, torch
from fcd_b200 import HeightMapPlan
from oracle import fcd_oracle as o, mask_oracle as mo

g = np.load(os.path.join(ROOT, "tests", "golden", "golden_fcd.npz"))
for n, a in ((256, 15.0), (512, 30.0)):
    ref = o.rotated_board(n, a=a, b=1.0)
    frames = []
    for peak in (12.0 * n / 256, 0.5):      # frame 0 wraps (unwrap path); frame 1 is the masked parity check
        _, uy, ux = o.gaussian_bump_displacement(n, (n * 0.47, n * 0.55), n / 6, peak)
        frames.append(o.rotated_board(n, a=a, b=1.0, uy=uy, ux=ux))
    frames = torch.from_numpy(np.stack(frames * 2)).cuda()
    plan = HeightMapPlan((n, n), 3)
    plan.bind(ref, square_size=o.board_square_size(n, a), height=1.0)
    mask = torch.zeros((n, n), dtype=torch.bool, device="cuda"); mask[n // 3: n // 2, n // 4: n // 2] = True
    h, ph = plan.execute(frames, phases=True, mask=mask)
    hm, _, _ = o.compute_height_map(ref, np.where(mask.cpu().numpy(), ref, frames[1].cpu().numpy()), o.board_square_size(n, a), height=1.0)
    hm *= ~mask.cpu().numpy()
    err = float(np.linalg.norm(h[1].cpu().numpy() - hm) / np.linalg.norm(hm))
    img = mo.synthetic_structure((n, n), 3)
    m = plan.structure_mask(np.stack([img] * 3), 15)
    c = plan.mask_center(m)
    print(n, "rel l2", err, "center", c[0], "residues", plan.count_residues(ph[0]))
    assert err < 1e-5
    plan.close()
# non-pruned RowDemod path (wide band): small period -> large radius
n = 256
ref = o.rotated_board(n, a=40.0, b=3.0)
plan = HeightMapPlan((n, n), 2)
plan.bind(ref, square_size=1.0, height=1.0)
print("band columns", plan.band_columns)
out = plan.execute(torch.from_numpy(np.stack([ref, ref])).cuda())
print("ok", float(out.abs().max()))
