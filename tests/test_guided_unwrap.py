"""Reliability-guided unwrap (csrc/fcd_unwrap.cuh; skimage.restoration.unwrap_phase at
pyfcd/fcd.py:119) against oracle/unwrap_herraez.c.

The device builds the minimum spanning tree that the sequential algorithm's merge order
implies, so on the SAME wrapped input the integer 2*pi field must be identical to the oracle's
up to one global constant -- with or without residues.  CPU tests run the kernel sources in
emulation; the gpu-marked tests run them on the device through the product's Python surface."""
import numpy as np
import pytest

from oracle import fcd_oracle as o

TWO_PI = 2 * np.pi


def wrapped_cases(shape, seed):
    rng = np.random.default_rng(seed)
    H, W = shape
    y, x = np.mgrid[0:H, 0:W]
    smooth = 9.0 * np.exp(-((y - H / 2) ** 2 + (x - W / 2.5) ** 2) / (2 * (H / 5) ** 2)) + 0.05 * x
    for noise in (0.0, 0.8, 2.5):
        yield noise, np.angle(np.exp(1j * (smooth + noise * rng.standard_normal(shape)))).astype(np.float32)


def assert_same_integer_field(unwrapped, wrapped):
    """unwrapped (device, float32) vs the oracle on the same wrapped map: one global 2*pi*k apart."""
    ref = o.unwrap_phase(wrapped.astype(np.float64))
    k = (unwrapped.astype(np.float64) - ref) / TWO_PI
    assert np.abs(k - np.round(k)).max() < 1e-5
    assert np.unique(np.round(k)).size == 1
    kk = (unwrapped.astype(np.float64) - wrapped) / TWO_PI
    assert np.abs(kk - np.round(kk)).max() < 1e-5 and np.round(kk)[0, 0] == 0     # pixel (0,0) keeps its value


def noisy_wrap_frame(golden):
    """The wrapping golden frame plus camera noise and a pattern-free blob: residues in both maps."""
    g = lambda k: golden[f"synth256_wrap.{k}"]
    frame = g("frame").copy()
    frame += 0.25 * np.random.default_rng(11).standard_normal(frame.shape).astype(np.float32)
    frame[100:130, 80:120] = 0.5
    return g("ref").astype(np.float64), frame, float(g("square_size"))


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a).astype(np.float64) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


# ----------------------------------------------------------------------------- CPU emulation
@pytest.mark.parametrize("shape", [(64, 64), (64, 128), (128, 64)])
def test_emulated_unwrap_matches_oracle(shape):
    from tests.emul_lib import EmulPlan
    plan = EmulPlan(shape)
    for noise, w in wrapped_cases(shape, 5):
        assert (o.count_residues(w.astype(np.float64)) > 50) == (noise > 0)
        assert_same_integer_field(plan.unwrap_phase(w), w)
    both = np.stack([w for _, w in wrapped_cases(shape, 6)])          # several maps share the Boruvka rounds
    got = plan.unwrap_phase(both)
    for i in range(both.shape[0]):
        assert_same_integer_field(got[i], both[i])
    plan.close()


def test_emulated_pipeline_with_residues(golden):
    from tests.emul_lib import EmulPlan
    from tests.test_emulated_kernels import bind_like_reference
    ref, frame, sq = noisy_wrap_frame(golden)
    plan = EmulPlan((256, 256), 2)
    bind_like_reference(plan, ref, sq)
    _, w = plan.execute(frame, phases=True, unwrap=0)
    assert min(plan.count_residues(w[0])) > 0
    h2, p2 = plan.execute(np.stack([frame, frame, frame]), phases=True, unwrap=2)   # 3 frames, 2 per launch
    for i in range(2):
        assert_same_integer_field(p2[0, i], w[0, i])
    assert np.array_equal(h2[0], h2[2]) and np.array_equal(p2[0], p2[2])
    assert np.array_equal(plan.execute(frame, unwrap=2)[0], h2[0])                    # internal phase workspace
    ho, _, _ = o.compute_height_map(ref, frame, sq, height=1.0)
    assert rel_l2(h2[0], ho) < 1e-4
    h1 = plan.execute(frame, unwrap=1)
    assert rel_l2(h1[0], ho) > 1e-2          # the scan path is path-dependent here: that is why mode 2 exists
    plan.close()


def test_emulated_auto_mode_flags_probes_and_redoes_only_what_it_must(golden):
    """unwrap = 3 inside the C ABI: a frame that cannot wrap is never looked at again, a frame that wraps without
    residues is probed and keeps its scan result, a frame with residues ends up bit-identical to mode 2."""
    from tests.emul_lib import EmulPlan
    from tests.test_emulated_kernels import bind_like_reference
    ref, noisy, sq = noisy_wrap_frame(golden)
    wrapping = golden["synth256_wrap.frame"]                      # 4-5 px displacement: 2*pi jumps, no residues
    flat = ref.astype(np.float32)                                 # the reference itself: zero phases
    frames = np.stack([flat, noisy, wrapping, noisy, noisy, flat, wrapping])
    plan = EmulPlan((256, 256), 4)                                # 7 frames, 4 per launch: two waves
    bind_like_reference(plan, ref, sq)
    h3, p3 = plan.execute(frames, phases=True, unwrap=3)
    flagged, guided = plan.last_auto()
    assert flagged == 5 and guided == [1, 3, 4]
    h1, p1 = plan.execute(frames, phases=True, unwrap=1)
    h2, p2 = plan.execute(frames, phases=True, unwrap=2)
    for i in (0, 2, 5, 6):
        assert np.array_equal(h3[i], h1[i]) and np.array_equal(p3[i], p1[i])
    for i in (1, 3, 4):
        assert np.array_equal(h3[i], h2[i]) and np.array_equal(p3[i], p2[i])
    # flagged frames that are consecutive are probed where they lie (no gather): same results
    h3c, p3c = plan.execute(frames[1:5], phases=True, unwrap=3)
    assert plan.last_auto() == (4, [0, 2, 3])
    assert np.array_equal(h3c, h3[1:5]) and np.array_equal(p3c, p3[1:5])
    # without a phases buffer the same height maps come out, and a per-frame mask follows its frame
    assert np.array_equal(plan.execute(frames, unwrap=3), h3)
    mask = np.zeros(frames.shape, np.uint8)
    mask[:, 100:140, 60:90] = 1
    mask[3, 10:20, :] = 1
    hm3 = plan.execute(frames, mask=mask, unwrap=3)
    _, gm = plan.last_auto()
    hm2 = plan.execute(frames, mask=mask, unwrap=2)
    hm1 = plan.execute(frames, mask=mask, unwrap=1)
    assert {1, 3, 4} <= set(gm) and not {0, 5} & set(gm)          # the mask edge adds residues to the wrapping frames
    for i in range(7):
        assert np.array_equal(hm3[i], hm2[i] if i in gm else hm1[i])
        assert not hm3[i][mask[i] != 0].any()
    plan.close()


# ----------------------------------------------------------------------------- device
@pytest.mark.gpu
def test_gpu_unwrap_matches_oracle():
    import torch
    import fcd_b200
    for shape in [(64, 128), (256, 256)]:
        plan = fcd_b200.HeightMapPlan(shape, 1)
        maps = np.stack([w for _, w in wrapped_cases(shape, 7)])
        got = plan.unwrap_phase(torch.from_numpy(maps).cuda()).cpu().numpy()
        for i in range(maps.shape[0]):
            assert_same_integer_field(got[i], maps[i])
        plan.close()


@pytest.mark.gpu
def test_gpu_pipeline_with_residues_auto_mode(golden):
    import torch
    import fcd_b200
    from pyfcd.fcd import fcd
    ref, frame, sq = noisy_wrap_frame(golden)
    ho, pho, _ = o.compute_height_map(ref, frame, sq, height=1.0)
    # drop-in surface: unwrap=True follows the reference's unwrapper
    hm, ph, _ = fcd.compute_height_map(ref, frame, sq, height=1.0)
    assert rel_l2(hm, ho) < 1e-4
    # batched API: clean frame + noisy frame; only the noisy one takes the guided path
    clean = golden["synth256_wrap.frame"]
    plan = fcd_b200.HeightMapPlan((256, 256), 4)
    plan.bind(ref, square_size=sq, height=1.0)
    frames = torch.from_numpy(np.stack([clean, frame, clean])).cuda()
    h_auto, p_auto = plan.execute(frames, phases=True, unwrap="auto")
    assert plan.last_guided_frames == [1]
    h_scan = plan.execute(frames, unwrap=True)
    assert torch.equal(h_auto[0], h_scan[0]) and torch.equal(h_auto[2], h_scan[2])
    assert rel_l2(h_auto[1].cpu().numpy(), ho) < 1e-4 and rel_l2(h_scan[1].cpu().numpy(), ho) > 1e-2
    _, w = plan.execute(frames[1], phases=True, unwrap=False)
    for i in range(2):
        assert_same_integer_field(p_auto[1, i].cpu().numpy(), w[i].cpu().numpy())
    h_g = plan.execute(frames, unwrap="herraez")
    assert torch.equal(h_g[1], h_auto[1])
    assert rel_l2(h_g[0].cpu().numpy(), h_scan[0].cpu().numpy()) < 1e-5     # no residues: both unwrappers agree
    plan.close()


@pytest.mark.gpu
def test_gpu_guided_unwrap_full_size():
    """2048^2 with residues: size-independent properties (integer field, residue-free result
    on a residue-free map equals the scan unwrap)."""
    import torch
    import fcd_b200
    n = 2048
    plan = fcd_b200.HeightMapPlan((n, n), 1)
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.arange(n, device="cuda", dtype=torch.float32)[:, None]
    x = torch.arange(n, device="cuda", dtype=torch.float32)[None, :]
    smooth = 40.0 * torch.exp(-((y - n / 2) ** 2 + (x - n / 2.5) ** 2) / (2 * (n / 5) ** 2)) + 0.01 * x
    for noise in (0.0, 0.7):
        ph = smooth + noise * torch.randn((n, n), device="cuda", generator=g)
        w = torch.atan2(torch.sin(ph), torch.cos(ph))
        u = plan.unwrap_phase(w)
        k = (u - w) / TWO_PI
        assert float((k - torch.round(k)).abs().max()) < 1e-4
        if noise == 0.0:
            assert plan.count_residues(w) == [0]
            d = torch.round((u - ph) / TWO_PI)
            assert float(d.min()) == float(d.max())          # the true phase up to one global 2*pi*k
        else:
            assert plan.count_residues(w)[0] > 1000
            # every tree edge is continuous and the result is deterministic
            assert torch.equal(u, plan.unwrap_phase(w))
    plan.close()


# ----------------------------------------------------------------------------- the reference's own example data
def real_pair():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_real.npz"))
    return g, g["ref"].astype(np.float32), g["frame"].astype(np.float32), float(g["square_size"]), g["layers"].tolist()


def test_emulated_real_camera_pair_needs_and_gets_the_guided_unwrap():
    """examples/fcd_example.py's pair (central 512^2 crop, golden = output of the unmodified reference):
    the wrapped phases contain residues, the scan unwrap is 20 % off, the guided unwrap matches."""
    from tests.emul_lib import EmulPlan
    from tests.test_emulated_kernels import bind_like_reference
    g, ref, frm, sq, layers = real_pair()
    plan = EmulPlan((512, 512), 1)
    _, _, cal = bind_like_reference(plan, ref, sq, height=o.height_from_layers(layers))
    assert cal == float(g["cal"])
    _, w = plan.execute(frm, phases=True, unwrap=0)
    assert plan.count_residues(w[0]) == list(g["residues"]) and min(g["residues"]) > 0
    assert rel_l2(plan.execute(frm, unwrap=2)[0], g["height_map"]) < 1e-5
    assert rel_l2(plan.execute(frm, unwrap=1)[0], g["height_map"]) > 0.05
    plan.close()


@pytest.mark.gpu
def test_gpu_real_camera_pair_through_the_drop_in():
    import torch
    import fcd_b200
    from pyfcd.fcd import fcd
    g, ref, frm, sq, layers = real_pair()
    hm, ph, cal = fcd.compute_height_map(ref, frm, sq, layers)              # examples/fcd_example.py:23
    assert cal == float(g["cal"]) and hm.dtype == np.float64 and hm.flags.writeable
    assert rel_l2(hm, g["height_map"]) < 1e-5
    # batched API on the integer camera frames, default unwrap = "auto"
    maps, _, _ = fcd_b200.compute_height_maps(ref, np.stack([g["frame"], g["frame"]]), sq, layers=layers)
    assert rel_l2(maps[1].cpu().numpy(), g["height_map"]) < 1e-5 and torch.equal(maps[0], maps[1])
