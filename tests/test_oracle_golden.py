"""CPU: pin the oracle restatement (oracle/fcd_oracle.py) against vectors produced by the
unmodified reference (oracle/make_golden.py) and against the reference's own known answers."""
import os

import numpy as np
import pytest

from oracle import fcd_oracle as o

REF_PICS = "/root/reference/examples/Pictures"


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.mark.parametrize("case", ["small", "wrap"])
def test_synthetic_height_map_matches_reference(golden, case):
    g = lambda k: golden[f"synth256_{case}.{k}"]
    hm, ph, cal = o.compute_height_map(g("ref"), g("frame"), float(g("square_size")), height=1.0)
    assert cal == float(g("cal"))
    assert rel_l2(hm, g("height_map")) < 1e-12
    assert np.allclose(ph[:, ::4, ::4], g("phases_sub"), rtol=0, atol=1e-11)
    carriers, _ = o.compute_carriers(g("ref").astype(np.float64), float(g("square_size")))
    assert np.array_equal(np.array([c.pixels for c in carriers]), g("pixels"))
    assert np.allclose(np.array([c.frequencies for c in carriers]), g("freqs"), rtol=1e-15)
    assert carriers[0].radius == float(g("radius"))


def test_wrap_case_really_wraps_and_scan_unwrap_agrees(golden):
    g = lambda k: golden[f"synth256_wrap.{k}"]
    assert g("phase_minmax")[1] > np.pi and abs(g("wrapped_minmax")[1]) <= np.pi
    assert list(g("residues")) == [0, 0]
    ref = g("ref").astype(np.float64)
    carriers, cal = o.compute_carriers(ref, float(g("square_size")))
    hm_h, ph_h = o.height_map_from_carriers(g("frame"), carriers, cal, 1.0, True, o.unwrap_phase)
    hm_s, ph_s = o.height_map_from_carriers(g("frame"), carriers, cal, 1.0, True, o.unwrap_scan)
    # identical up to one global 2*pi*k per map; height map invariant to it
    for i in range(2):
        d = (ph_h[i] - ph_s[i]) / (2 * np.pi)
        assert np.allclose(d, np.rint(d.flat[0]), atol=1e-9)
    assert rel_l2(hm_s, hm_h) < 1e-11


def test_val_validator_case(golden):
    hm, ph, cal = o.compute_height_map(golden["val256.I0"], golden["val256.I"], 256 / 30, height=1)
    assert cal == float(golden["val256.cal"])
    assert rel_l2(hm, golden["val256.height_map"]) < 1e-12
    err = np.max(np.abs(hm - golden["val256.h"])) * 100 / np.max(np.abs(hm))
    assert err < 1.5  # the quantity examples/val_example.py:41 prints
    # skipping the unwrap is badly wrong on this input
    hm_nw, _, _ = o.compute_height_map(golden["val256.I0"], golden["val256.I"], 256 / 30, height=1, unwrap=False)
    assert np.max(np.abs(hm_nw - golden["val256.h"])) * 100 / np.max(np.abs(hm_nw)) > 20


@pytest.mark.parametrize("i", [0, 1, 2])
def test_carrier_search_on_camera_like_reference(golden, i):
    img = golden[f"noisy{i}.image"].astype(np.float64)
    cal, peaks = o.compute_calibration_factor(0.0022, img)
    assert np.array_equal(np.array(peaks), golden[f"noisy{i}.peaks"])
    assert cal == float(golden[f"noisy{i}.cal"])


def test_height_from_layers(golden):
    assert o.height_from_layers(golden["layers.example"].tolist()) == float(golden["layers.height"])
    assert float(golden["layers.height"]) == pytest.approx(0.0329956084466042, rel=1e-14)
    with pytest.raises(Warning):
        o.resolve_height(layers=[[1, 1]], height=1.0)
    assert o.resolve_height() == 1


@pytest.mark.parametrize("shape", [(64, 64), (64, 128)])
def test_integrate_in_fourier_quirk(golden, shape):
    key = f"integrate{shape[0]}x{shape[1]}"
    h = o.integrate_in_fourier(golden[key + ".gx"], golden[key + ".gy"], 0.37)
    assert np.allclose(h, golden[key + ".h"], rtol=0, atol=1e-13)
    assert abs(h.mean()) < 1e-12


def test_label8_raster_order_and_connectivity():
    b = np.zeros((6, 8), bool)
    b[1, 5] = b[2, 4] = True          # diagonal neighbours -> one blob (8-connectivity)
    b[1, 1] = True                    # separate blob, same row, earlier column
    b[4, 2] = True
    lab, n = o.label8(b)
    assert n == 3
    assert lab[1, 1] == 1 and lab[1, 5] == 2 and lab[2, 4] == 2 and lab[4, 2] == 3


def test_disk_is_strict_and_clipped():
    m = o.disk_mask((10, 10), (2, 8), 3.0)
    assert m[2, 8] and not m[2, 5] and m[2, 6] and m[0, 9]
    assert m.sum() == sum(1 for r in range(10) for c in range(10) if (r - 2) ** 2 + (c - 8) ** 2 < 9)


def test_unwrap_herraez_on_ramp():
    y, x = np.mgrid[0:40, 0:50]
    true = 0.31 * x + 0.22 * y - 7.0
    w = np.angle(np.exp(1j * true))
    u = o.unwrap_phase(w)
    d = (u - true) / (2 * np.pi)
    assert np.allclose(d, np.rint(d[0, 0]), atol=1e-9)
    assert o.count_residues(w) == 0


@pytest.mark.skipif(not os.path.isdir(REF_PICS), reason="reference fixtures not on this machine")
def test_known_answers_from_reference_fixtures():
    import cv2
    load = lambda p: cv2.imread(os.path.join(REF_PICS, p), cv2.IMREAD_UNCHANGED).astype(np.float64)
    cal, peaks = o.compute_calibration_factor(0.002, load("reference_df.tif"))
    golden_cal = np.load(os.path.join(REF_PICS, "mask", "maps", "calibration_factor.npy"))
    assert cal == golden_cal[0]  # the only numeric golden in the reference repo, bit-exact
    assert [p.tolist() for p in peaks] == [[564, 566], [458, 564]]
    assert [p.tolist() for p in o.find_peaks(load("reference_2.png"))] == [[443, 590], [434, 443]]
    assert [p.tolist() for p in o.find_peaks(load("prueba1_20250317_122608_C1S0001000001.tif"))] == [[557, 558], [466, 557]]


@pytest.mark.parametrize("shape,seed", [((24, 37), 1), ((40, 32), 2), ((33, 33), 3)])
def test_unwrap_oracle_is_integration_along_the_minimum_spanning_tree(shape, seed):
    """Second, independent statement of what oracle/unwrap_herraez.c computes (and of what the CUDA unwrapper relies
    on, csrc/fcd_unwrap.cuh): merging neighbour pairs in ascending order of summed reliabilities and skipping pairs
    inside a group is Kruskal's algorithm, so the unwrapped map is the integral of the wrapped differences along the
    minimum spanning tree under the strict order (cost, edge index: horizontal edges before vertical, raster order).
    Here the tree comes from scipy (edge weights = ranks in that order, hence unique) and the integration from a
    breadth-first walk; reliabilities are recomputed in numpy with the oracle's border filler."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import breadth_first_order, minimum_spanning_tree
    rng = np.random.default_rng(seed)
    H, W = shape
    y, x = np.mgrid[0:H, 0:W]
    smooth = 7.0 * np.exp(-((y - H / 2) ** 2 + (x - W / 3) ** 2) / (2 * (H / 4) ** 2)) + 0.11 * x
    w = np.angle(np.exp(1j * (smooth + 1.2 * rng.standard_normal(shape))))
    assert o.count_residues(w) > 10
    got = o.unwrap_phase(w)

    def wrap(d):
        return np.where(d > np.pi, d - 2 * np.pi, np.where(d < -np.pi, d + 2 * np.pi, d))

    # reliabilities: wrapped second differences inside, the oracle's deterministic filler on the border
    n = H * W
    lcg, rel = 0x9E3779B97F4A7C15, np.empty(n)
    for i in range(n):
        lcg = (lcg * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        rel[i] = 9999999.0 + (lcg >> 11) / 9007199254740992.0
    rel = rel.reshape(shape)
    c = w[1:-1, 1:-1]
    h = wrap(w[1:-1, :-2] - c) - wrap(c - w[1:-1, 2:])
    v = wrap(w[:-2, 1:-1] - c) - wrap(c - w[2:, 1:-1])
    d1 = wrap(w[:-2, :-2] - c) - wrap(c - w[2:, 2:])
    d2 = wrap(w[:-2, 2:] - c) - wrap(c - w[2:, :-2])
    rel[1:-1, 1:-1] = h * h + v * v + d1 * d1 + d2 * d2
    idx = np.arange(n).reshape(shape)
    p = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()])          # horizontal edges first, raster order
    q = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()])
    cost = rel.ravel()[p] + rel.ravel()[q]
    rank = np.empty(len(p)); rank[np.argsort(cost, kind="stable")] = np.arange(1, len(p) + 1)
    tree = minimum_spanning_tree(coo_matrix((rank, (p, q)), shape=(n, n)).tocsr())
    sym = tree + tree.T
    order, pred = breadth_first_order(sym, 0, directed=False)
    flat, k = w.ravel(), np.zeros(n, dtype=np.int64)
    for node in order[1:]:
        par = pred[node]
        d = flat[par] - flat[node]                                           # continue the parent's branch
        k[node] = k[par] + (1 if d > np.pi else (-1 if d < -np.pi else 0))
    want = flat + 2 * np.pi * k
    diff = (got.ravel() - want) / (2 * np.pi)
    assert np.abs(diff - np.round(diff)).max() < 1e-9
    assert np.unique(np.round(diff)).size == 1                               # equal up to one global 2*pi*k
