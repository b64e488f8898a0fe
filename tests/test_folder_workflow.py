"""GPU: the batch driver analyze.folder (pydata/analyze.py:141-286) -- .tif frames in, *_map.npy /
centers.txt / calibration_factor.npy out -- against the oracles run frame by frame the way the
reference loops (mask -> np.where(mask, reference, frame) -> compute_height_map -> *= ~mask)."""
import os

import numpy as np
import pytest

from oracle import fcd_oracle as o
from oracle import mask_oracle as mo

pytestmark = pytest.mark.gpu

LAYERS = [[5.7e-3, 1.0003], [1.2e-2, 1.48899], [4.3e-2, 1.34], [80e-2, 1.0003]]    # examples/fcd_example.py:17


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a).astype(np.float64) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def write_dataset(d, n=256, frames=5, structure=False, dtype=np.uint8):
    import cv2
    scale = 200 if dtype == np.uint8 else 900
    ref = np.round(o.rotated_board(n, a=15.0, b=1.0) * scale).astype(dtype)
    cv2.imwrite(os.path.join(d, "reference.tif"), ref)
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:n, 0:n]
    out = []
    for k in range(frames):
        _, uy, ux = o.gaussian_bump_displacement(n, (110.0 + 8 * k, 140.0 - 5 * k), 28.0, 0.5 + 0.1 * k)
        fr = o.rotated_board(n, a=15.0, b=1.0, uy=uy, ux=ux) * scale
        if structure:                                           # a dark floating ring with a cavity inside
            rr = np.hypot(yy - (120 + 3 * k), xx - (125 - 2 * k))
            fr = np.where((rr > 42) & (rr < 60), 0.05 * scale, fr)
        fr = np.round(fr + rng.normal(0, 0.5, fr.shape)).clip(0, np.iinfo(dtype).max).astype(dtype)
        cv2.imwrite(os.path.join(d, f"img_{k:04d}.tif"), fr)
        out.append(fr)
    return ref, out


def test_folder_plain(tmp_path):
    from pydata.analyze import analyze
    import fcd_b200
    ref, frames = write_dataset(str(tmp_path), dtype=np.uint16)
    sq = o.board_square_size(256, 15.0)
    analyze.folder(str(tmp_path / "reference.tif"), str(tmp_path), LAYERS, sq, batch=2)
    maps = sorted(f for f in os.listdir(tmp_path / "maps") if f.endswith("_map.npy"))
    assert maps == [f"img_{k:04d}_map.npy" for k in range(5)]
    reff = ref.astype(np.float32)
    assert np.array_equal(analyze.load_image(str(tmp_path / "img_0003.tif")), frames[3].astype(np.float32))
    cal = np.load(tmp_path / "maps" / "calibration_factor.npy")
    for k, name in enumerate(maps):
        got = np.load(tmp_path / "maps" / name)
        assert got.dtype == np.float32 and got.shape == (256, 256)
        want, _, calo = o.compute_height_map(reff, frames[k].astype(np.float32), sq, LAYERS)
        assert rel_l2(got, want) < 1e-4
        assert cal.shape == (1,) and cal[0] == calo
    # resume (analyze.py:181-182): everything exists -> nothing is recomputed
    plan = fcd_b200.get_plan((256, 256), 2)
    n0 = plan.launch_count
    analyze.folder(str(tmp_path / "reference.tif"), str(tmp_path), LAYERS, sq, batch=2)
    assert plan.launch_count == n0
    os.remove(tmp_path / "maps" / maps[-1])
    analyze.folder(str(tmp_path / "reference.tif"), str(tmp_path), LAYERS, sq, batch=2)
    assert os.path.exists(tmp_path / "maps" / maps[-1]) and plan.launch_count > n0


def test_folder_masked_workflow(tmp_path):
    from pydata.analyze import analyze
    ref, frames = write_dataset(str(tmp_path), frames=3, structure=True)
    sq = o.board_square_size(256, 15.0)
    analyze.folder(str(tmp_path / "reference.tif"), str(tmp_path), LAYERS, sq, smoothed=14, batch=4)
    reff = ref.astype(np.float32)
    lines = open(tmp_path / "maps" / "centers.txt").read().splitlines()
    assert len(lines) == 3
    for k in range(3):
        fr = frames[k].astype(np.float32)
        m = mo.mask(fr, 14)
        c = mo.center(m)
        assert lines[k] == f"{k}\t{c}"
        want, _, _ = o.compute_height_map(reff, np.where(m == 1, reff, fr), sq, LAYERS)
        want *= ~m
        got = np.load(tmp_path / "maps" / f"img_{k:04d}_map.npy")
        assert np.all(got[m] == 0)
        assert rel_l2(got, want) < 1e-4
    with pytest.raises(NotImplementedError):
        analyze.folder(str(tmp_path / "reference.tif"), str(tmp_path), LAYERS, sq, polar=True)
