"""GPU (B200): the CUDA path, called through the product's Python surface and the C ABI,
against the float64 oracle and the reference-derived golden vectors.  Tolerance from the
north star: relative L2 of the float32 height map <= 1e-4, identical carrier pixels."""
import numpy as np
import pytest

from oracle import fcd_oracle as o

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def phase_dev(ph, ref):
    k = np.rint(np.median((ph - ref) / (2 * np.pi)))
    return np.abs(ph - ref - 2 * np.pi * k).max()


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    from pyfcd.fcd import fcd, fourier, Carrier
    import fcd_b200
    return dict(fcd=fcd, fourier=fourier, Carrier=Carrier, eng=fcd_b200, torch=torch)


def test_native_library_is_loaded(api):
    api["eng"].get_plan((64, 64))
    assert any("libfcd_b200.so" in l for l in open("/proc/self/maps"))


@pytest.mark.parametrize("case", ["small", "wrap"])
def test_drop_in_compute_height_map_golden(api, golden, case):
    g = lambda k: golden[f"synth256_{case}.{k}"]
    hm, ph, cal = api["fcd"].compute_height_map(g("ref"), g("frame"), float(g("square_size")), height=1.0)
    assert hm.dtype == np.float64 and hm.flags.writeable and ph.shape == (2, 256, 256)
    assert cal == float(g("cal"))
    assert rel_l2(hm, g("height_map")) < TOL
    assert rel_l2(hm, g("height_map")) < 1e-5        # what float32 actually achieves
    _, pho, _ = o.compute_height_map(g("ref"), g("frame"), float(g("square_size")), height=1.0)
    for i in range(2):
        assert phase_dev(ph[i], pho[i]) < 5e-5
    hm *= ~np.zeros(hm.shape, bool)                   # callers do this in place (analyze.py:255)


def test_carriers_identical_to_reference(api, golden):
    g = lambda k: golden[f"synth256_small.{k}"]
    carriers, cal = api["fcd"].compute_carriers(g("ref").astype(np.float64), float(g("square_size")))
    assert np.array_equal(np.array([c.pixels for c in carriers]), g("pixels"))
    assert np.allclose(np.array([c.frequencies for c in carriers]), g("freqs"), rtol=1e-15)
    assert carriers[0].radius == float(g("radius")) and cal == float(g("cal"))
    oc, _ = o.compute_carriers(g("ref").astype(np.float64), float(g("square_size")))
    for a, b in zip(carriers, oc):
        assert np.array_equal(a.mask, b.mask)
        assert np.abs(a.ccsgn - b.ccsgn).max() < 1e-13
    calf, peaks = api["fcd"].compute_calibration_factor(float(g("square_size")), g("ref"))
    assert calf == cal


@pytest.mark.parametrize("i", [0, 1, 2])
def test_carrier_search_camera_like(api, golden, i):
    img = golden[f"noisy{i}.image"]
    for dtype in (np.float32, np.float64):
        peaks = api["fourier"].find_peaks(img.astype(dtype))
        assert np.array_equal(np.array(peaks), golden[f"noisy{i}.peaks"])
    cal, _ = api["fcd"].compute_calibration_factor(0.0022, img.astype(np.float64))
    assert cal == float(golden[f"noisy{i}.cal"])
    spec = o.highpassed_spectrum(img.astype(np.float64))
    got = api["fourier"].find_peak_locations(spec, 0.5 * spec.max(), 4)
    assert [p.tolist() for p in got] == [p.tolist() for p in o.find_peak_locations(spec, 0.5 * spec.max(), 4)]


def test_validator_surface_with_unwrap(api, golden):
    """pyval/val.py surface (phases reach +-pi): carriers given (its tie is round-off decided)."""
    I0, I = golden["val256.I0"], golden["val256.I"].astype(np.float32)
    plan = api["eng"].HeightMapPlan((256, 256), 1)
    pix = golden["val256.pixels"]
    plan.bind(I0, calibration_factor=float(golden["val256.cal"]), height=1, peaks=(pix[0], pix[1]),
              radius=float(golden["val256.radius"]))
    hm = plan.execute(api["torch"].from_numpy(I).cuda()).cpu().numpy()
    hmo, _, _ = o.compute_height_map(I0, I, 256 / 30, height=1)
    assert rel_l2(hm, hmo) < 1e-5
    assert rel_l2(hm, golden["val256.height_map"]) < TOL
    hm_nw = plan.execute(api["torch"].from_numpy(I).cuda(), unwrap=False).cpu().numpy()
    assert rel_l2(hm_nw, hmo) > 0.1
    plan.close()


def test_batched_api_matches_oracle_and_is_order_independent(api):
    n = 256
    rng = np.random.default_rng(5)
    ref = o.rotated_board(n, a=15.0, b=1.0)
    frames = []
    for _ in range(7):
        cy, cx = rng.uniform(0.35 * n, 0.65 * n, 2)
        _, uy, ux = o.gaussian_bump_displacement(n, (cy, cx), rng.uniform(n / 12, n / 6), rng.uniform(0.2, 6.0))
        frames.append(o.rotated_board(n, a=15.0, b=1.0, uy=uy, ux=ux))
    frames = np.stack(frames)
    sq = o.board_square_size(n, 15.0)
    hm, ph, cal = api["eng"].compute_height_maps(ref, frames, sq, height=0.5, return_phases=True, frames_per_launch=3)
    hm = hm.cpu().numpy()
    carriers, calo = o.compute_carriers(ref.astype(np.float64), sq)
    assert cal == calo
    for i in range(7):
        hmo, _ = o.height_map_from_carriers(frames[i], carriers, calo, 0.5)
        assert rel_l2(hm[i], hmo) < 1e-5
    # same frames, different batch positions / chunking -> bitwise identical (sharding invariant)
    perm = np.array([3, 0, 6, 1, 5, 2, 4])
    hm2, _, _ = api["eng"].compute_height_maps(ref, frames[perm], sq, height=0.5, frames_per_launch=2)
    assert np.array_equal(hm2.cpu().numpy(), hm[perm])


def test_config1_1024(api):
    """BASELINE.json configs[0]: single 1024^2 synthetic checkerboard + Gaussian bump."""
    n = 1024
    ref, frames, truth = o.synthetic_frames(n, 1)
    sq = o.board_square_size(n)
    hm, ph, cal = api["fcd"].compute_height_map(ref, frames[0], sq, height=1.0)
    hmo, pho, calo = o.compute_height_map(ref, frames[0], sq, height=1.0)
    assert cal == calo == 1.0
    assert rel_l2(hm, hmo) < 1e-5
    carriers, _ = api["fcd"].compute_carriers(ref, sq)
    assert [c.pixels.tolist() for c in carriers] == [[512 + 57, 512 + 63], [512 - 63, 512 + 57]]
    t = truth[0] - truth[0].mean()
    assert rel_l2(hm, t) < 0.05                       # FCD's own accuracy on this surface


def test_full_size_2048_properties(api):
    """BASELINE.json configs[1] shape.  One frame against the oracle, the rest through
    size-independent properties: zero mean, reference -> flat, batch invariance, truth."""
    torch = api["torch"]
    n = 2048
    ref, frames, truth = o.synthetic_frames(n, 3)
    sq = o.board_square_size(n)
    plan = api["eng"].HeightMapPlan((n, n), 2)
    cal = plan.bind(ref, square_size=sq, height=1.0)
    assert cal == 1.0
    assert [p.tolist() for p in plan.peaks] == [[1138, 1150], [898, 1138]]
    batch = torch.from_numpy(np.concatenate([frames, ref[None], frames[:1]])).cuda()
    hm = plan.execute(batch).cpu().numpy()
    hmo, _, _ = o.compute_height_map(ref, frames[0], sq, height=1.0)
    assert rel_l2(hm[0], hmo) < 1e-5
    assert np.array_equal(hm[0], hm[4])                                 # batch position invariance
    assert np.abs(hm[3]).max() < 1e-4 * np.abs(hm[0]).max()             # reference frame -> flat
    for i in range(3):
        assert abs(hm[i].mean()) < 1e-5 * np.abs(hm[i]).max()          # DC bin is zero (fourier.py:130)
        t = truth[i] - truth[i].mean()
        assert rel_l2(hm[i], t) < 0.05
    plan.close()


def test_full_size_4096_properties(api):
    """BASELINE.json configs[2] shape (4096^2; the non-pruned demodulation kernel and the 16x16x16 radix
    plan): size-independent properties only -- carrier pixels, zero mean, reference -> flat, batch
    invariance, linearity in 1/height, agreement with the analytic surface, guided == scan without residues."""
    torch = api["torch"]
    n = 4096
    ref, frames, truth = o.synthetic_frames(n, 2)
    sq = o.board_square_size(n)
    plan = api["eng"].HeightMapPlan((n, n), 2)
    cal = plan.bind(ref, square_size=sq, height=1.0)
    assert cal == 1.0
    assert [p.tolist() for p in plan.peaks] == [[2048 + 228, 2048 + 252], [2048 - 252, 2048 + 228]]
    batch = torch.from_numpy(np.concatenate([frames, ref[None], frames[:1]])).cuda()
    hm_t = plan.execute(batch)
    hm = hm_t.cpu().numpy()
    assert np.array_equal(hm[0], hm[3])
    assert np.abs(hm[2]).max() < 1e-4 * np.abs(hm[0]).max()
    for i in range(2):
        assert abs(hm[i].mean()) < 1e-5 * np.abs(hm[i]).max()
        t = truth[i] - truth[i].mean()
        assert rel_l2(hm[i], t) < 0.05
    plan.set_height(height=0.5)
    assert rel_l2(plan.execute(batch[:1]).cpu().numpy(), hm[:1] * 2) < 1e-6
    plan.set_height(height=1.0)
    assert rel_l2(plan.execute(batch[:1], unwrap="herraez").cpu().numpy(), hm[:1]) < 1e-5
    plan.close()


def test_full_size_4096_vs_oracle(api):
    """BASELINE.json configs[2] shape against the float64 oracle itself (one frame; the oracle takes ~20 s):
    the 16x16x16 radix plan and its pruned first pass must land where 2048^2 does."""
    torch = api["torch"]
    n = 4096
    ref, frames, _ = o.synthetic_frames(n, 1)
    sq = o.board_square_size(n)
    hmo, pho, calo = o.compute_height_map(ref, frames[0], sq, height=1.0)
    plan = api["eng"].HeightMapPlan((n, n), 1)
    assert plan.bind(ref, square_size=sq, height=1.0) == calo
    fr = torch.from_numpy(frames[:1]).cuda()
    for mode in ("auto", "scan", "herraez"):
        hm, ph = plan.execute(fr, phases=True, unwrap=mode)
        assert rel_l2(hm[0].cpu().numpy(), hmo) < 1e-5, mode
        for i in range(2):
            assert phase_dev(ph[0, i].cpu().numpy().astype(np.float64), pho[i]) < 1e-4
    plan.close()


def test_full_size_2048_masked_typed_auto_vs_oracle(api):
    """2048^2 through the default mode of the drop-in (unwrap="auto") with everything on at once: 16-bit camera
    frames widened in K1, a per-frame mask (reference substituted before, zeros after: analyze.py:231,255), one
    clean frame, one wrapping frame, one noisy wrapping frame (residues -> reliability-guided)."""
    torch = api["torch"]
    n = 2048
    ref, frames, _ = o.synthetic_frames(n, 1)
    sq = o.board_square_size(n)
    h, uy, ux = o.gaussian_bump_displacement(n, (900.0, 1100.0), 260.0, 4.5)
    wrapping = o.rotated_board(n, uy=uy, ux=ux)
    noisy = wrapping + 0.25 * np.random.default_rng(5).standard_normal((n, n)).astype(np.float32)
    noisy[800:1000, 700:1100] = 0.5
    stack = np.stack([frames[0], wrapping, noisy]).astype(np.float32)
    # the camera's integers: the oracle gets exactly the values the device widens (load_image's astype, analyze.py:40)
    cam = np.clip(np.rint(stack * 60000.0), 0, 65535).astype(np.uint16)
    ref_cam = np.clip(np.rint(ref * 60000.0), 0, 65535).astype(np.uint16).astype(np.float32)
    mask = np.zeros((3, n, n), bool)
    mask[:, 1200:1500, 300:800] = True
    mask[2, 100:180, :] = True
    plan = api["eng"].HeightMapPlan((n, n), 2)
    plan.bind(ref_cam, square_size=sq, height=1.0)
    hm = plan.execute(torch.from_numpy(cam).cuda(), mask=torch.from_numpy(mask).cuda(), unwrap="auto").cpu().numpy()
    assert 2 in plan.last_guided_frames and 0 not in plan.last_guided_frames and plan.last_flagged_frames >= 2
    for i in range(3):
        fr = np.where(mask[i], ref_cam, cam[i].astype(np.float32))
        want, _, _ = o.compute_height_map(ref_cam, fr, sq, height=1.0)
        want *= ~mask[i]
        assert rel_l2(hm[i], want) < 1e-5, i
        assert not hm[i][mask[i]].any()
    plan.close()


def test_drop_in_cache_survives_other_binds(api, golden):
    """The drop-in keeps per-reference state between calls; any other bind of the shared plan (folder, batched
    API, a direct bind) must invalidate it."""
    ga = lambda k: golden[f"synth256_small.{k}"]
    gb = lambda k: golden[f"synth256_wrap.{k}"]
    fcd, eng = api["fcd"], api["eng"]
    ref_a, sq_a = ga("ref"), float(ga("square_size"))
    ref_b = np.roll(ga("ref"), 3, axis=1) * 0.9 + 0.02            # same carriers, different pixels / ccsgn
    hm_a0, _, _ = fcd.compute_height_map(ref_a, ga("frame"), sq_a, height=1.0)
    eng.compute_height_maps(ref_b, ga("frame")[None], sq_a, height=1.0, frames_per_launch=1)   # rebinds the cached plan
    hm_a1, _, _ = fcd.compute_height_map(ref_a, ga("frame"), sq_a, height=1.0)
    assert np.array_equal(hm_a0, hm_a1)
    eng.get_plan((256, 256), 1).bind(ref_b, square_size=sq_a, height=1.0)                     # direct bind
    hm_a2, _, _ = fcd.compute_height_map(ref_a, ga("frame"), sq_a, height=1.0)
    assert np.array_equal(hm_a0, hm_a2)
    # truthy non-bool `unwrap` means the reference's unwrapper (fcd.py:119 does `if unwrap:`)
    hm_w, _, _ = fcd.compute_height_map(gb("ref"), gb("frame"), float(gb("square_size")), height=1.0, unwrap=np.True_)
    hm_w1, _, _ = fcd.compute_height_map(gb("ref"), gb("frame"), float(gb("square_size")), height=1.0, unwrap=1)
    assert np.array_equal(hm_w, hm_w1) and rel_l2(hm_w, gb("height_map")) < 1e-5


def plane_wave_board(n0, n1, p, q, uy=0.0, ux=0.0, eps=0.1):
    """Exactly periodic two-carrier pattern on an n0 x n1 grid: integer cycle counts p = (rows, cols), q likewise."""
    y = np.arange(n0, dtype=np.float64)[:, None] - uy
    x = np.arange(n1, dtype=np.float64)[None, :] - ux
    P = 2 * np.pi * (p[0] * y / n0 + p[1] * x / n1)
    Q = 2 * np.pi * (q[0] * y / n0 + q[1] * x / n1)
    return 0.5 + 0.25 * ((1 + eps) * np.cos(P) - np.cos(Q)) / (1 + eps / 2)


@pytest.mark.parametrize("shape", [(600, 800), (375, 250)])
def test_drop_in_any_shape(api, shape):
    """The reference takes any image shape (pyfcd/fcd.py:14).  Shapes that are not powers of two run the float64
    stage-level path (Bluestein transforms on the hand-written kernels, fcd_b200/generic.py): same carriers, same
    calibration factor, height map and phases as the float64 oracle, clean and with residues, masked and batched."""
    torch = api["torch"]
    n0, n1 = shape
    p, q = (round(0.055 * n0), round(0.005 * n1)), (-round(0.005 * n0), round(0.055 * n1))
    ref = plane_wave_board(n0, n1, p, q)
    y = np.arange(n0, dtype=np.float64)[:, None] - 0.45 * n0
    x = np.arange(n1, dtype=np.float64)[None, :] - 0.55 * n1
    sg = min(shape) / 7.0
    g = np.exp(-(y * y + x * x) / (2 * sg * sg))
    frames = []
    for peak in (0.6, 4.5):                                       # no wrapping / wrapping
        amp = peak * sg * np.exp(0.5)
        frames.append(plane_wave_board(n0, n1, p, q, uy=amp * y / (sg * sg) * g, ux=amp * x / (sg * sg) * g))
    noisy = frames[1] + 0.25 * np.random.default_rng(3).standard_normal(shape)
    noisy[n0 // 3:n0 // 3 + n0 // 9, n1 // 4:n1 // 4 + n1 // 8] = 0.5
    frames.append(noisy)
    sq = 7.5
    fcd, eng = api["fcd"], api["eng"]
    carriers, cal = fcd.compute_carriers(ref, sq)
    oc, ocal = o.compute_carriers(ref, sq)
    assert cal == ocal and [c.pixels.tolist() for c in carriers] == [list(map(int, c.pixels)) for c in oc]
    for a, b in zip(carriers, oc):
        assert np.array_equal(a.mask, b.mask) and np.abs(a.ccsgn - b.ccsgn).max() < 1e-12
    for k, fr in enumerate(frames):
        hm, ph, c = fcd.compute_height_map(ref, fr, sq, height=0.8)
        hmo, pho, _ = o.compute_height_map(ref, fr, sq, height=0.8)
        assert c == ocal and hm.shape == shape and hm.dtype == np.float64 and hm.flags.writeable
        assert rel_l2(hm, hmo) < (1e-9 if k < 2 else 1e-5), (k, rel_l2(hm, hmo))
        for i in range(2):
            assert phase_dev(ph[i], pho[i]) < (1e-8 if k < 2 else 1e-4)
    plan = eng.get_plan(shape, 1)
    assert not plan.fused
    # batched API with a per-frame mask: reference substituted before, zeros after (analyze.py:231,255)
    mask = np.zeros((3,) + shape, bool)
    mask[:, n0 // 2:n0 // 2 + 40, n1 // 5:n1 // 5 + 60] = True
    stack32 = np.stack(frames).astype(np.float32)              # camera frames are float32 (analyze.load_image)
    maps, _, _ = eng.compute_height_maps(ref, stack32, sq, height=0.8, mask=torch.from_numpy(mask).cuda(), plan=plan)
    guided = plan.last_guided_frames                            # the mask edge can add residues to the wrapping frame
    assert 2 in guided and 0 not in guided
    for k in range(3):
        want, _, _ = o.compute_height_map(ref, np.where(mask[k], ref, stack32[k].astype(np.float64)), sq, height=0.8)
        want *= ~mask[k]
        assert rel_l2(maps[k].cpu().numpy(), want) < (1e-5 if k in guided else 1e-9), k
        assert not maps[k].cpu().numpy()[mask[k]].any()
    with pytest.raises(Exception):
        plan.structure_mask(torch.from_numpy(frames[0].astype(np.float32)).cuda())        # fused plans only


def test_plot_helpers_run_with_a_stub_matplotlib(api, golden, monkeypatch):
    """examples/fcd_example.py:21-22 call fcd.fft_peaks(reference) and compute_calibration_factor(..., plot=True).
    matplotlib is not installed here; a recording stub shows that both run and draw what the reference draws
    (pyfcd/fcd.py:90-99, 157-175: the chosen peaks and the candidate list)."""
    import sys
    import types
    calls = []

    class Rec:
        def __init__(self, name):
            self._n = name

        def __getattr__(self, k):
            def f(*a, **kw):
                calls.append((self._n + "." + k, a, kw))
                return (Rec("fig"), Rec("ax")) if k == "subplots" else None
            return f

    plt = types.ModuleType("matplotlib.pyplot")
    rec = Rec("plt")
    for name in ("subplots", "legend", "tight_layout", "show"):
        setattr(plt, name, getattr(rec, name))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    g = lambda k: golden[f"synth256_small.{k}"]
    ref = g("ref")
    cal, peaks = api["fcd"].compute_calibration_factor(float(g("square_size")), ref, plot=True)
    assert cal == float(g("cal")) and np.array_equal(np.array(peaks), g("pixels"))
    assert any(c[0] == "ax.imshow" for c in calls) and any(c[0] == "plt.show" for c in calls)
    calls.clear()
    api["fcd"].fft_peaks(ref)
    drawn = [c for c in calls if c[0] == "ax.plot"]
    want = [tuple(p) for p in o.find_peak_locations(o.highpassed_spectrum(ref.astype(np.float64)),
                                                    0.5 * o.highpassed_spectrum(ref.astype(np.float64)).max(), 4)]
    assert [(c[1][1], c[1][0]) for c in drawn[:len(want)]] == want          # candidates in the reference's order
    assert (drawn[-2][1][1], drawn[-2][1][0]) == tuple(g("pixels")[0])      # rightmost
    assert (drawn[-1][1][1], drawn[-1][1][0]) == tuple(g("pixels")[1])      # perpendicular
    shown = [c for c in calls if c[0] == "ax.imshow"][0][1][0]
    spec = np.log1p(np.fft.fftshift(np.abs(np.fft.fft2(ref.astype(np.float64) - ref.astype(np.float64).mean()))))
    assert np.allclose(shown, spec, rtol=1e-9, atol=1e-9)


def test_mask_workflow(api, golden):
    g = lambda k: golden[f"synth256_small.{k}"]
    torch = api["torch"]
    ref, frame, sq = g("ref"), g("frame"), float(g("square_size"))
    mask = np.zeros((256, 256), bool)
    mask[100:140, 90:150] = True
    hm, _, _ = api["eng"].compute_height_maps(ref, frame[None], sq, height=1.0, mask=torch.from_numpy(mask))
    hm = hm.cpu().numpy()[0]
    hmo, _, _ = o.compute_height_map(ref, np.where(mask, ref, frame), sq, height=1.0)
    hmo *= ~mask
    assert rel_l2(hm, hmo) < 1e-5 and np.all(hm[mask] == 0)


@pytest.mark.parametrize("shape", [(64, 128), (512, 256), (4096, 64)])
def test_rectangular(api, shape):
    n0, n1 = shape
    y = np.arange(n0)[:, None].astype(np.float64)
    x = np.arange(n1)[None, :].astype(np.float64)
    ky, kx = 2 * np.pi * round(n0 * 15.0 / 256) / n0, 2 * np.pi * round(n1 * 15.0 / 256) / n1
    cy, cx, s = 0.45 * n0, 0.55 * n1, min(n0, n1) / 7.0
    hgt = 0.5 * s * np.exp(-((y - cy) ** 2 + (x - cx) ** 2) / (2 * s * s))
    uy, ux = (y - cy) / s ** 2 * hgt, (x - cx) / s ** 2 * hgt
    board = lambda yy, xx: (0.5 + 0.25 * (1.1 * np.cos(ky * yy + 2 * np.pi * xx / n1)
                                          + np.cos(kx * xx - 2 * np.pi * yy / n0)) / 2.1).astype(np.float32)
    ref, frame = board(y, x), board(y - uy, x - ux)
    hm, ph, cal = api["fcd"].compute_height_map(ref, frame, 3.3, height=0.7)
    hmo, pho, calo = o.compute_height_map(ref, frame, 3.3, height=0.7)
    assert cal == calo and rel_l2(hm, hmo) < 2e-5


def test_stage_level_methods(api, golden):
    g = lambda k: golden[f"synth256_wrap.{k}"]
    ref, frame, sq = g("ref").astype(np.float64), g("frame").astype(np.float64), float(g("square_size"))
    fcd, fourier = api["fcd"], api["fourier"]
    carriers, cal = fcd.compute_carriers(ref, sq)
    ph = fcd.compute_phases(np.fft.fft2(frame), carriers)
    oc, _ = o.compute_carriers(ref, sq)
    pho = o.compute_phases(np.fft.fft2(frame), oc)
    for i in range(2):
        assert phase_dev(ph[i], pho[i]) < 1e-9
    uv = fcd.compute_displacement_field(pho, carriers)
    assert np.allclose(uv, o.compute_displacement_field(pho, oc), rtol=1e-12, atol=1e-12)
    key = "integrate64x128"
    h = fourier.integrate_in_fourier(golden[key + ".gx"], golden[key + ".gy"], 0.37)
    assert np.allclose(h, golden[key + ".h"], rtol=0, atol=1e-12)
    kx, ky = fourier.wavenumber_meshgrid((64, 128))
    okx, oky = o.wavenumber_meshgrid((64, 128))
    assert np.array_equal(kx, okx) and np.array_equal(ky, oky)


def test_errors_like_the_reference(api):
    fcd = api["fcd"]
    ref = o.rotated_board(256, a=15.0, b=1.0)
    with pytest.raises(Warning):
        fcd.compute_height_map(ref, ref, 1.0, layers=[[1, 1], [1, 1.3], [1, 1.3], [1, 1]], height=1.0)
    with pytest.raises(ValueError):
        fcd.compute_height_map(np.zeros((256, 256), np.float32), ref, 1.0)      # no carrier peak
    with pytest.raises(Exception):
        fcd.compute_height_map(np.zeros((100, 256), np.float32), np.zeros((100, 256), np.float32), 1.0)
    layers = [[5.7e-2, 1.0003], [1.2e-2, 1.48899], [4.3e-2, 1.34], [80e-2, 1.0003]]
    assert fcd.height_from_layers(layers) == o.height_from_layers(layers)
    h1, _, _ = fcd.compute_height_map(ref, ref * 0.9 + 0.01, 1.0, layers)
    assert h1.shape == (256, 256)


def test_cufft_pipeline_agrees(api):
    """The cuFFT (torch.fft) version of the pipeline used as the benchmark comparator."""
    torch = api["torch"]
    from fcd_b200.cufft_pipeline import CufftPipeline
    n = 512
    ref = o.rotated_board(n, a=30.0, b=2.0)
    frames = []
    for peak in (0.5, 9.0):
        _, uy, ux = o.gaussian_bump_displacement(n, (250.0, 270.0), 60.0, peak)
        frames.append(o.rotated_board(n, a=30.0, b=2.0, uy=uy, ux=ux))
    frames = torch.from_numpy(np.stack(frames)).cuda()
    plan = api["eng"].HeightMapPlan((n, n), 2)
    plan.bind(ref, square_size=o.board_square_size(n, 30.0), height=1.0)
    ours = plan.execute(frames)
    theirs = CufftPipeline(plan).execute(frames)
    for i in range(2):
        assert float(torch.linalg.vector_norm(ours[i] - theirs[i]) / torch.linalg.vector_norm(theirs[i])) < 2e-5
    plan.close()


def test_integer_frames_residues_and_reference_cache(api, golden):
    torch = api["torch"]
    g = lambda k: golden[f"synth256_small.{k}"]
    ref8 = np.round(g("ref") * 200).astype(np.uint8)
    frm8 = np.round(g("frame") * 200).astype(np.uint8)
    sq = float(g("square_size"))
    hm8, _, cal = api["eng"].compute_height_maps(ref8.astype(np.float32), frm8[None], sq, height=1.0)
    hmf, _, _ = api["eng"].compute_height_maps(ref8.astype(np.float32), frm8[None].astype(np.float32), sq, height=1.0)
    assert torch.equal(hm8, hmf)
    hmo, _, _ = o.compute_height_map(ref8.astype(np.float32), frm8.astype(np.float32), sq, height=1.0)
    assert rel_l2(hm8[0].cpu().numpy(), hmo) < 1e-5
    # residue guard on wrapped phases
    plan = api["eng"].HeightMapPlan((256, 256), 1)
    gw = lambda k: golden[f"synth256_wrap.{k}"]
    plan.bind(gw("ref"), square_size=float(gw("square_size")), height=1.0)
    _, ph = plan.execute(torch.from_numpy(gw("frame")).cuda(), phases=True, unwrap=False)
    assert plan.count_residues(ph) == [0, 0]
    noisy = torch.rand((256, 256), device="cuda") * 6.2 - 3.1
    assert plan.count_residues(noisy) == [o.count_residues(noisy.cpu().numpy())]
    plan.close()
    # drop-in: second call with the same reference reuses the bound state, a new height is honoured
    fcd = api["fcd"]
    a, _, _ = fcd.compute_height_map(g("ref"), g("frame"), sq, height=1.0)
    n0 = api["eng"].get_plan((256, 256), 1).launch_count
    b, _, _ = fcd.compute_height_map(g("ref").copy(), g("frame"), sq, height=0.5)
    n1 = api["eng"].get_plan((256, 256), 1).launch_count
    assert n1 - n0 <= 8                      # only the per-frame kernels ran (7 stages + the residue count)
    assert np.allclose(b, a / 0.5, rtol=1e-6, atol=1e-9)
    c, _, _ = fcd.compute_height_map(g("ref") * 0.5, g("frame") * 0.5, sq, height=1.0)   # different reference -> re-bind
    assert rel_l2(c, a) < 1e-5


def test_structure_mask_and_center_bit_exact(api):
    """analyze.mask / analyze.center on the device: byte / integer results."""
    import os
    from oracle import mask_oracle as mo
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_mask.npz"))
    sys_path_has_pydata = __import__("pydata.analyze", fromlist=["analyze"]).analyze
    for i in range(4):
        shape = tuple(int(v) for v in g[f"case{i}.shape"])
        img = mo.synthetic_structure(shape, int(g[f"case{i}.seed"]))
        smoothed = int(g[f"case{i}.smoothed"])
        m_gold = np.unpackbits(g[f"case{i}.mask"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
        plan = api["eng"].HeightMapPlan(shape, 1)
        flipped = np.ascontiguousarray(img[::-1, ::-1])
        m = plan.structure_mask(np.stack([img, flipped, img, img, flipped]), smoothed)
        assert np.array_equal(m[0].cpu().numpy(), m_gold) and np.array_equal(m[3].cpu().numpy(), m_gold)
        assert np.array_equal(m[4].cpu().numpy(), mo.mask(flipped, smoothed))
        c = plan.mask_center(m)
        assert c[0] == tuple(int(v) for v in g[f"case{i}.center"]) and c[1] == mo.center(m[1].cpu().numpy())
        plan.close()
        mm, cc = sys_path_has_pydata.mask(img, smoothed=smoothed, find_center=True)      # drop-in classmethods
        assert np.array_equal(mm, m_gold) and cc == c[0]
    big = mo.synthetic_structure((2048, 2048), 11)
    plan = api["eng"].HeightMapPlan((2048, 2048), 1)
    m = plan.structure_mask(big, 15)
    assert np.array_equal(m.cpu().numpy(), mo.mask(big, 15))
    assert plan.mask_center(m)[0] == mo.center(m.cpu().numpy())
    plan.close()
    with pytest.raises(UnboundLocalError):
        sys_path_has_pydata.center(np.ones((64, 64), bool))
