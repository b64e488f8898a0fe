"""CPU: the product's kernel sources, compiled for CPU emulation (tests/emul_lib.py), against
the oracle and the reference-derived golden vectors.  This is what guards the kernel
arithmetic and index maps in the GPU-less build container; the same comparisons run on the
real device in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import fcd_oracle as o
from tests.emul_lib import EmulPlan, lib
from fcd_b200 import _native


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a).astype(np.float64) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def bind_like_reference(plan, ref, square_size, height=1.0):
    peaks = plan.find_peaks(ref)
    k = o.pixel_to_wavenumber(ref.shape, [peaks[0], peaks[1]])
    cal = 2 * square_size / (2 * np.pi / np.mean(np.abs(k)))
    radius = np.linalg.norm(peaks[0] - peaks[1]) / 2
    plan.bind(ref, peaks, radius, cal, height)
    return peaks, radius, cal


def phase_offset_dev(ph, ref):
    k = np.rint(np.median((ph - ref) / (2 * np.pi)))
    return np.abs(ph - ref - 2 * np.pi * k).max()


@pytest.mark.parametrize("case", ["small", "wrap"])
def test_golden_synthetic(golden, case):
    g = lambda k: golden[f"synth256_{case}.{k}"]
    plan = EmulPlan((256, 256))
    peaks, radius, cal = bind_like_reference(plan, g("ref").astype(np.float64), float(g("square_size")))
    assert np.array_equal(np.array(peaks), g("pixels"))          # identical carrier pixels
    assert cal == float(g("cal")) and radius == float(g("radius"))
    hm, ph = plan.execute(np.stack([g("frame")] * 3), phases=True)   # 3 frames: two launch chunks
    for i in range(3):
        assert rel_l2(hm[i], g("height_map")) < 1e-5                 # budget 1e-4 (north star)
    assert np.array_equal(hm[0], hm[1]) and np.array_equal(hm[0], hm[2])
    _, pho, _ = o.compute_height_map(g("ref"), g("frame"), float(g("square_size")), height=1.0)
    for i in range(2):
        assert phase_offset_dev(ph[0, i], pho[i]) < 2e-5
    plan.close()


def test_validator_case_with_given_carriers(golden):
    """val.py's axis-aligned board: its 'rightmost' choice is a 1-ulp tie, so the
    reference's carriers are passed in; the surface wraps, exercising the unwrap kernels."""
    I0, I = golden["val256.I0"], golden["val256.I"]
    pix = golden["val256.pixels"]
    plan = EmulPlan((256, 256), 1)
    plan.bind(I0, pix, float(golden["val256.radius"]), float(golden["val256.cal"]), 1.0)
    hm = plan.execute(I.astype(np.float32))
    # float32 frame vs the reference's float64 frame: compare with the oracle on the same input
    hmo, _, _ = o.compute_height_map(I0, I.astype(np.float32), 256 / 30, height=1)
    assert rel_l2(hm[0], hmo) < 1e-5
    assert rel_l2(hm[0], golden["val256.height_map"]) < 1e-4
    hm_nw = plan.execute(I.astype(np.float32), unwrap=False)
    assert rel_l2(hm_nw[0], hmo) > 0.1
    plan.close()


@pytest.mark.parametrize("i", [0, 1, 2])
def test_carrier_search_camera_like(golden, i):
    img = golden[f"noisy{i}.image"]
    plan = EmulPlan(img.shape)
    for dtype in (np.float64, np.float32):
        peaks = plan.find_peaks(img.astype(dtype))
        assert np.array_equal(np.array(peaks), golden[f"noisy{i}.peaks"])
    spec, mx = plan.highpass_spectrum(img.astype(np.float64))
    ospec = o.highpassed_spectrum(img.astype(np.float64))
    assert np.allclose(spec, ospec, rtol=0, atol=1e-9 * ospec.max())
    assert np.array_equal(spec == 0, ospec == 0)                   # same high-pass support
    locs = plan.peak_locations(ospec, 0.5 * ospec.max(), 4)
    assert [l.tolist() for l in locs] == [l.tolist() for l in o.find_peak_locations(ospec, 0.5 * ospec.max(), 4)]
    plan.close()


def test_peak_locations_semantics():
    """ascending sort, first four, border lines cleared, 8-connectivity, first maximum."""
    img = np.zeros((64, 64))
    img[10, 10], img[11, 11], img[10, 12] = 5.0, 7.0, 7.0      # one blob (diagonals), max 7 first at (10,12)
    img[30, 5] = 9.0
    img[40, 40] = 3.0
    img[50, 20] = 4.0
    img[20, 50] = 8.0
    img[0, 30] = img[63, 7] = img[17, 0] = img[33, 63] = 100.0   # on the border: ignored
    plan = EmulPlan((64, 64))
    got = [l.tolist() for l in plan.peak_locations(img, 1.0, 4)]
    want = [l.tolist() for l in o.find_peak_locations(img, 1.0, 4)]
    assert got == want == [[40, 40], [50, 20], [10, 12], [20, 50]]
    assert plan.peak_locations(img, 1000.0, 4) == []
    with pytest.raises(ValueError):
        plan.find_peaks(np.zeros((64, 64)))                      # reference: min() of empty list
    plan.close()


@pytest.mark.parametrize("shape", [(64, 64), (64, 128), (128, 64), (512, 256)])
def test_rectangular_and_small_shapes(shape):
    n0, n1 = shape
    a0, b0 = n0 * 15.0 / 256, 1.0
    y = np.arange(n0)[:, None].astype(np.float64)
    x = np.arange(n1)[None, :].astype(np.float64)
    ky, kx = 2 * np.pi * round(a0) / n0, 2 * np.pi * round(n1 * 15.0 / 256) / n1
    cy, cx, s = 0.45 * n0, 0.55 * n1, min(n0, n1) / 7.0
    hgt = 0.5 * s * np.exp(-((y - cy) ** 2 + (x - cx) ** 2) / (2 * s * s))
    uy, ux = (y - cy) / s ** 2 * hgt, (x - cx) / s ** 2 * hgt

    def board(yy, xx):
        return (0.5 + 0.25 * (1.1 * np.cos(ky * yy + 0.3 * kx * xx * 0) * 0 + 1.1 * np.cos(ky * yy + b0 * 2 * np.pi * xx / n1)
                              + np.cos(kx * xx - b0 * 2 * np.pi * yy / n0)) / 2.1).astype(np.float32)

    ref, frame = board(y, x), board(y - uy, x - ux)
    sq = 3.3
    plan = EmulPlan(shape, 1)
    peaks, radius, cal = bind_like_reference(plan, ref.astype(np.float64), sq, height=0.7)
    assert [p.tolist() for p in peaks] == [p.tolist() for p in o.find_peaks(ref.astype(np.float64))]
    hm, ph = plan.execute(frame, phases=True)
    hmo, pho, calo = o.compute_height_map(ref, frame, sq, height=0.7)
    assert cal == calo
    assert rel_l2(hm[0], hmo) < 2e-5
    plan.close()


def test_carrier_attributes_and_fft2(golden):
    ref = golden["synth256_small.ref"].astype(np.float64)
    sq = float(golden["synth256_small.square_size"])
    plan = EmulPlan((256, 256))
    bind_like_reference(plan, ref, sq)
    cars, _ = o.compute_carriers(ref, sq)
    for i in range(2):
        assert np.array_equal(plan.mask(i), cars[i].mask)
        assert np.abs(plan.ccsgn(i) - cars[i].ccsgn).max() < 1e-13
        assert np.abs(plan.ccsgn(i, c128=False) - cars[i].ccsgn).max() < 1e-6
    rng = np.random.default_rng(0)
    z = rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256))
    assert np.abs(plan.fft2(z) - np.fft.fft2(z)).max() < 1e-10
    assert np.abs(plan.fft2(z, inverse=True) - np.fft.ifft2(z)).max() < 1e-14
    plan.close()


def test_mask_workflow(golden):
    """analyze.py:229-234,254-255: frame := where(mask, reference, frame); height *= ~mask."""
    g = lambda k: golden[f"synth256_small.{k}"]
    ref, frame, sq = g("ref"), g("frame"), float(g("square_size"))
    mask = np.zeros((256, 256), bool)
    mask[100:140, 90:150] = True
    plan = EmulPlan((256, 256))
    bind_like_reference(plan, ref, sq)
    hm = plan.execute(frame, mask=mask)
    sub = np.where(mask, ref, frame)
    hmo, _, _ = o.compute_height_map(ref, sub, sq, height=1.0)
    hmo *= ~mask
    assert rel_l2(hm[0], hmo) < 1e-5
    assert np.all(hm[0][mask] == 0)
    hm2 = plan.execute(np.stack([frame, frame]), mask=np.stack([mask, np.zeros_like(mask)]))
    assert np.array_equal(hm2[0], hm[0]) and rel_l2(hm2[1], g("height_map")) < 1e-5
    plan.close()


def test_error_behaviour():
    l = lib()
    import ctypes
    h = ctypes.c_void_p()
    assert l.fcd_plan_create(100, 256, 1, ctypes.byref(h)) == _native.FCD_OK               # not a power of two: a generic plan
    assert l.fcd_plan_is_fused(h) == 0
    f = np.zeros((1, 100, 256), np.float32)
    assert l.fcd_execute(h, f.ctypes.data, 1, f.ctypes.data, None, None, 0, 1, None) == _native.FCD_ERR_INVALID
    assert b"powers of two" in l.fcd_last_error()                                            # the fused pipeline is not for it
    assert l.fcd_plan_destroy(h) == _native.FCD_OK
    assert l.fcd_plan_create(256, 8192, 1, ctypes.byref(h)) == _native.FCD_ERR_INVALID       # beyond the largest plan
    assert l.fcd_plan_create(3000, 100, 1, ctypes.byref(h)) == _native.FCD_ERR_INVALID       # generic plans stop at 2048
    assert b"powers of two" in l.fcd_last_error()
    plan = EmulPlan((64, 64))
    f = np.zeros((1, 64, 64), np.float32)
    rc = l.fcd_execute(plan.h, f.ctypes.data, 1, f.ctypes.data, None, None, 0, 1, None)
    assert rc == _native.FCD_ERR_STATE                                                       # execute before bind
    plan.close()


def test_integer_camera_frames_and_set_height(golden):
    """uint8 / uint16 frames are widened in the first kernel (analyze.load_image's astype)."""
    g = lambda k: golden[f"synth256_small.{k}"]
    ref8 = np.round(g("ref") * 200).astype(np.uint8)
    frm8 = np.round(g("frame") * 200).astype(np.uint8)
    sq = float(g("square_size"))
    plan = EmulPlan((256, 256))
    bind_like_reference(plan, ref8.astype(np.float32), sq)
    h_f32 = plan.execute(frm8.astype(np.float32))
    h_u8 = plan.execute(frm8)
    h_u16 = plan.execute(frm8.astype(np.uint16) * 1)
    assert np.array_equal(h_u8, h_f32) and np.array_equal(h_u16, h_f32)
    hmo, _, _ = o.compute_height_map(ref8.astype(np.float32), frm8.astype(np.float32), sq, height=1.0)
    assert rel_l2(h_u8[0], hmo) < 1e-5
    plan.set_height(0.25)
    assert rel_l2(plan.execute(frm8)[0], hmo / 0.25) < 1e-5
    plan.close()


def test_residue_guard(golden):
    g = lambda k: golden[f"synth256_wrap.{k}"]
    plan = EmulPlan((256, 256))
    bind_like_reference(plan, g("ref").astype(np.float64), float(g("square_size")))
    _, ph = plan.execute(g("frame"), phases=True, unwrap=False)
    assert plan.count_residues(ph[0]) == [0, 0] == list(g("residues"))
    rng = np.random.default_rng(3)
    noisy = rng.uniform(-np.pi, np.pi, (2, 256, 256)).astype(np.float32)
    got = plan.count_residues(noisy)
    assert got == [o.count_residues(noisy[0]), o.count_residues(noisy[1])] and got[0] > 1000
    plan.close()


def test_thread_order_independence(golden):
    """Racecheck by construction: inside a phase there is no barrier, so running the emulated
    threads descending or odd-first must reproduce the ascending run bit for bit (pipeline with
    unwrap path, masks, phases output; carrier search; mask/centre kernels)."""
    from oracle import mask_oracle as mo
    l = lib()
    g = lambda k: golden[f"synth256_wrap.{k}"]
    ref, frame, sq = g("ref").astype(np.float64), g("frame"), float(g("square_size"))
    mask = np.zeros((256, 256), bool); mask[90:120, 60:200] = True
    img = mo.synthetic_structure((256, 256), 2)

    def run(order):
        l.fcd_emul_set_thread_order(order)
        try:
            plan = EmulPlan((256, 256), 2)
            peaks, radius, cal = bind_like_reference(plan, ref, sq)
            h, ph = plan.execute(np.stack([frame, golden["synth256_small.frame"], frame]), phases=True)
            hm = plan.execute(frame, mask=mask)
            sm = plan.structure_mask(img, 14)
            c = plan.mask_center(sm)
            cc = plan.ccsgn(0, c128=False)
            plan.close()
            return [np.array(peaks), h, ph, hm, sm, np.array(c), cc]
        finally:
            l.fcd_emul_set_thread_order(0)

    base = run(0)
    for order in (1, 2):
        for a, b in zip(base, run(order)):
            assert np.array_equal(a, b)


def test_empty_batches_and_bad_modes():
    """Edge cases of the C ABI: empty batches are no-ops, an unknown unwrap mode is an argument error."""
    from fcd_b200 import synthetic as syn
    plan = EmulPlan((64, 64), 2)
    ref = syn.rotated_board(64, a=6.0, b=1.0).astype(np.float64)
    bind_like_reference(plan, ref, syn.board_square_size(64, 6.0))
    assert plan.execute(np.zeros((0, 64, 64), np.float32)).shape == (0, 64, 64)
    assert plan.unwrap_phase(np.zeros((0, 64, 64), np.float32)).shape == (0, 64, 64)
    assert plan.structure_mask(np.zeros((0, 64, 64), np.float32)).shape == (0, 64, 64)
    assert plan.mask_center(np.zeros((0, 64, 64), np.uint8)) == []
    with pytest.raises(_native.FcdError) as e:
        plan.execute(np.zeros((1, 64, 64), np.float32), unwrap=4)
    assert e.value.code == _native.FCD_ERR_INVALID
    # the reference frame itself demodulates to a flat surface (zero phases), with every unwrap mode
    for mode in (0, 1, 2, 3):
        h = plan.execute(ref.astype(np.float32), unwrap=mode)
        assert np.abs(h).max() < 1e-4
    plan.close()


# ----------------------------------------------------------------------------- shapes that are not powers of two
@pytest.mark.parametrize("shape", [(60, 100), (75, 64), (64, 96), (33, 47)])
def test_generic_plan_fft2_any_shape(shape):
    """A plan of any shape is GENERIC: fcd_fft2_c128 runs Bluestein's chirp convolution on padded power-of-two
    transforms and must equal scipy's fft2 / ifft2 (the reference calls them on whatever shape it is given)."""
    import scipy.fft as sfft
    rng = np.random.default_rng(2)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    plan = EmulPlan(shape)
    f = plan.fft2(x)
    assert np.abs(f - sfft.fft2(x)).max() < 1e-11 * np.abs(f).max()
    assert np.abs(plan.fft2(f, inverse=True) - x).max() < 1e-12
    with pytest.raises(_native.FcdError) as e:                      # the fused pipeline needs powers of two
        plan.execute(np.zeros((1,) + shape, np.float32))
    assert e.value.code in (_native.FCD_ERR_INVALID, _native.FCD_ERR_STATE)
    plan.close()


def test_generic_plan_carriers_match_oracle():
    """Carrier search, disk masks and ccsgn on a 120 x 150 board (odd half sizes exercise the fftshift bookkeeping)."""
    from fcd_b200 import synthetic as syn
    n0, n1 = 120, 150
    y = np.arange(n0, dtype=np.float64)[:, None]
    x = np.arange(n1, dtype=np.float64)[None, :]
    ref = 0.5 + 0.25 * (1.1 * np.cos(2 * np.pi * (12 * y / n0 - 10 * x / n1)) - np.cos(2 * np.pi * (9 * y / n0 + 14 * x / n1))) / 1.05
    plan = EmulPlan((n0, n1))
    spec, mx = plan.highpass_spectrum(ref)
    want = o.highpassed_spectrum(ref)
    assert np.abs(spec - want).max() < 1e-9 * want.max() and abs(mx - want.max()) < 1e-9 * mx
    peaks = plan.find_peaks(ref)
    opeaks = o.find_peaks(ref)
    assert np.array_equal(np.array(peaks), np.array(opeaks))
    carriers, cal = o.compute_carriers(ref, 3.0)
    plan.bind(ref, peaks, carriers[0].radius, cal)
    for i in range(2):
        assert np.array_equal(plan.mask(i), carriers[i].mask)
        assert np.abs(plan.ccsgn(i) - carriers[i].ccsgn).max() < 1e-12
    # the unwrap and the residue count are shape-agnostic
    w = np.angle(np.exp(1j * (0.15 * x + 0.05 * y + 3 * np.exp(-((y - 60) ** 2 + (x - 70) ** 2) / 800)))).astype(np.float32)
    u = plan.unwrap_phase(w)
    assert plan.count_residues(w[None]) == [0]
    assert np.abs((u - w) / (2 * np.pi) - np.round((u - w) / (2 * np.pi))).max() < 1e-5
    assert np.abs(u - o.unwrap_phase(w.astype(np.float64)) - (u - o.unwrap_phase(w.astype(np.float64)))[0, 0]).max() < 1e-4
    plan.close()
