"""Temporal harmonic analysis (analyze.block_amplitude / block_split, pydata/analyze.py:542-641,
365-417): oracle against goldens produced by the unmodified reference, the kernels in CPU
emulation against both, the frame-shard -> row-band exchange on gloo, and (gpu) the product
surface on the device."""
import os

import numpy as np
import pytest

from oracle import temporal_oracle as to
from oracle.make_golden_temporal import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_temporal.npz")


@pytest.fixture(scope="module")
def gt():
    return np.load(GOLD)


def case_maps(ci):
    n, shape, nb, mode, zero, given = CASES[ci]
    return to.synthetic_maps(n, shape, nb, seed=ci + 1), n, nb, mode, zero, (37.5 if given else None)


def assert_block_close(amps, phases, g_amps, g_phases):
    assert amps.shape == g_amps.shape
    nan = np.isnan(g_amps)
    assert np.array_equal(np.isnan(amps), nan) and np.array_equal(np.isnan(phases), np.isnan(g_phases))
    assert np.all(amps[..., -1] == 0) and np.all(phases[..., -1] == 0)      # mode+1 planes, the last never filled
    z = np.where(nan, 0, amps * np.exp(1j * phases))
    zg = np.where(nan, 0, g_amps.astype(np.float64) * np.exp(1j * g_phases.astype(np.float64)))
    assert np.abs(z - zg).max() <= 2e-5 * np.abs(zg).max()


# ------------------------------------------------------------------------------- oracle
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_oracle_matches_reference_golden(gt, ci):
    maps, n, nb, mode, zero, f0 = case_maps(ci)
    for b in range(nb):
        h, a, p, f = to.block_amplitude(maps, f0=f0, tasa=500, mode=mode, num_blocks=nb, block_index=b, zero=zero)
        assert np.array_equal(np.array(h), gt[f"case{ci}.block{b}.harmonics"]) and f == float(gt[f"case{ci}.block{b}.f0"])
        assert np.array_equal(a.astype(np.float32), gt[f"case{ci}.block{b}.amps"], equal_nan=True)
        assert np.array_equal(p.astype(np.float32), gt[f"case{ci}.block{b}.phases"], equal_nan=True)


def test_oracle_degenerate_block_has_no_peak():
    maps = np.ones((64, 64, 64), np.float32)            # constant in time: only DC, no local maximum
    out = to.block_amplitude(maps, num_blocks=4, block_index=0, mode=2)
    assert len(out) == 5 and out[3] is None


# ------------------------------------------------------------------------------- emulated kernels
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_emulated_kernels_match_golden(gt, ci):
    from fcd_b200 import temporal as tp
    from tests.emul_lib import EmulPlan
    maps, n, nb, mode, zero, f0 = case_maps(ci)
    bpr = int(np.sqrt(nb))
    plan = EmulPlan((64, 64))
    freqs = tp.positive_frequencies(n, 500)
    if f0 is None:
        mean, valid = plan.temporal_mean_spectrum(maps, maps[0], zero, bpr)
        for b in range(nb):
            ref = to.mean_spectrum(maps, nb, b, zero)
            assert np.abs(mean[b] - ref).max() <= 1e-6 * ref.max()
        f0s = [tp.pick_f0(mean[b], freqs) for b in range(nb)]
    else:
        f0s = [f0] * nb
    assert f0s == [float(gt[f"case{ci}.block{b}.f0"]) for b in range(nb)]
    bins = np.array([tp.harmonic_bins(f, mode, freqs)[1] for f in f0s], np.int32)
    acc = plan.temporal_harmonics(maps, bins, zero=zero, bpr=bpr)
    acc_chunked = plan.temporal_harmonics(maps, bins, zero=zero, bpr=bpr, chunks=(7, 30))
    assert np.allclose(acc, acc_chunked, rtol=0, atol=1e-12 * np.abs(acc).max())
    # additive over frame shards (what the multi-GPU all-reduce relies on)
    half = n // 2
    part = (plan.temporal_harmonics(maps[:half], bins, n_total=n, t0=0, zero=zero, bpr=bpr)
            + plan.temporal_harmonics(maps[half:], bins, n_total=n, t0=half, zero=zero, bpr=bpr))
    assert np.allclose(acc, part, rtol=0, atol=1e-12 * np.abs(acc).max())
    amps, phases = plan.temporal_finalize(acc, n, (64, 64), first=maps[0])
    bs = 64 // bpr
    for b in range(nb):
        i, j = divmod(b, bpr)
        sl = (slice(i * bs, (i + 1) * bs), slice(j * bs, (j + 1) * bs))
        assert_block_close(amps[sl], phases[sl], gt[f"case{ci}.block{b}.amps"], gt[f"case{ci}.block{b}.phases"])
    plan.close()


@pytest.mark.parametrize("n,n1", [(48, 8), (35, 7), (90, 10)])
def test_emulated_two_level_transform(n, n1):
    """Long series take N = N1 * N2 in two levels of chirp-convolution transforms; forced here on short ones."""
    from tests.emul_lib import EmulPlan
    maps = to.synthetic_maps(n, (64, 64), 4, seed=3)
    plan = EmulPlan((64, 64))
    mean, valid = plan.temporal_mean_spectrum(maps, maps[0], 0.01, 2, n1=n1)
    direct, _ = plan.temporal_mean_spectrum(maps, maps[0], 0.01, 2)
    for b in range(4):
        ref = to.mean_spectrum(maps, 4, b, 0.01)
        assert np.abs(mean[b] - ref).max() <= 1e-6 * ref.max()
        assert np.abs(mean[b] - direct[b]).max() <= 1e-6 * ref.max()
    plan.close()


def test_unsupported_series_length_is_an_error():
    from fcd_b200 import _native
    from tests.emul_lib import EmulPlan, lib
    assert lib().fcd_temporal_frames_supported(4096) and lib().fcd_temporal_frames_supported(2048)
    assert lib().fcd_temporal_frames_supported(1000) and lib().fcd_temporal_frames_supported(3000)
    assert lib().fcd_temporal_frames_supported(20000)                   # 160 x 125
    assert not lib().fcd_temporal_frames_supported(2 * 2053)            # 2 x a prime above 2048
    plan = EmulPlan((64, 64))
    with pytest.raises(_native.FcdError):
        plan.temporal_harmonics(np.zeros((4, 64, 64), np.float32), np.zeros((4, 9), np.int32))    # > 8 bins
    plan.close()


def _exchange_worker(rank, ws, port, n_total, q):
    import torch
    import torch.distributed as dist
    from fcd_b200 import temporal as tp
    from fcd_b200.engine import shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    full = torch.arange(n_total * 8 * 6, dtype=torch.float32).view(n_total, 8, 6)
    a, b = shard_range(n_total, rank, ws)
    band = tp.frames_to_row_bands(full[a:b].contiguous(), n_total)
    ok = torch.equal(band, full[:, rank * 4:(rank + 1) * 4])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_frames_to_row_bands_two_ranks_gloo():
    """The only real exchange of the repo (frame shards -> pixel bands for the temporal FFT)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# ------------------------------------------------------------------------------- device
@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_gpu_block_amplitudes_match_golden(gt, ci):
    import torch
    from fcd_b200 import temporal as tp
    maps, n, nb, mode, zero, f0 = case_maps(ci)
    res = tp.block_amplitudes(torch.from_numpy(maps).cuda(), f0=f0, tasa=500, mode=mode, num_blocks=nb, zero=zero,
                              chunk_frames=23)
    for b in range(nb):
        h, a, p, f = res.block(b)
        assert np.array_equal(np.array(h), gt[f"case{ci}.block{b}.harmonics"]) and f == float(gt[f"case{ci}.block{b}.f0"])
        assert_block_close(a, p, gt[f"case{ci}.block{b}.amps"], gt[f"case{ci}.block{b}.phases"])


@pytest.mark.gpu
def test_gpu_drop_in_block_amplitude_from_folder(gt, tmp_path):
    from pydata.analyze import analyze
    ci = 0
    maps, n, nb, mode, zero, f0 = case_maps(ci)
    for t in range(n):
        np.save(tmp_path / f"img_{t:05d}_map.npy", maps[t])
    np.save(tmp_path / "calibration_factor.npy", np.array([1.0]))
    for b in range(nb):
        h, a, p, f = analyze.block_amplitude(str(tmp_path), tasa=500, mode=mode, num_blocks=nb, block_index=b)
        assert f == float(gt[f"case{ci}.block{b}.f0"])
        assert_block_close(a, p, gt[f"case{ci}.block{b}.amps"], gt[f"case{ci}.block{b}.phases"])
    sp = analyze.block_split(str(tmp_path), num_blocks=nb, block_index=nb - 1)
    assert np.array_equal(sp, to.block_split(maps, None, nb, nb - 1), equal_nan=True)


@pytest.mark.gpu
def test_gpu_full_size_series_properties():
    """512 frames of 512^2 (64 blocks): every block's f0 is recovered exactly, amplitudes follow
    the synthetic mode shape, and a 1000-frame series (chirp-convolution path) agrees with numpy."""
    import torch
    from fcd_b200 import temporal as tp
    n, H, nb = 512, 512, 64
    t = torch.arange(n, device="cuda", dtype=torch.float32) / 500.0
    bpr, bs = 8, 64
    maps = torch.empty((n, H, H), device="cuda")
    cyc = {}
    for b in range(nb):
        i, j = divmod(b, bpr)
        cyc[b] = 11 + 3 * b
        f = cyc[b] * 500.0 / n
        maps[:, i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = (0.5 + (1 + 0.01 * b) * torch.cos(2 * np.pi * f * t + 0.1 * b))[:, None, None]
    maps += 0.01 * torch.randn_like(maps)
    res = tp.block_amplitudes(maps, tasa=500, mode=2, num_blocks=nb)
    for b in range(nb):
        assert abs(res.f0[b] - cyc[b] * 500.0 / n) < 1e-9
        i, j = divmod(b, bpr)
        a = res.amps[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
        assert abs(float(a[..., 0].mean()) - 0.5) < 1e-3          # first entry: the DC bin (0 * f0)
        assert abs(float(a[..., 1].mean()) - (1 + 0.01 * b)) < 1e-3
    # a long series: 2500 = 50 x 50 frames takes the two-level path on its own
    tt = torch.arange(2500, device="cuda", dtype=torch.float32)
    long = (torch.cos(2 * np.pi * 137 * tt / 2500)[:, None, None] * (1 + torch.rand((1, 64, 64), device="cuda"))
            + 0.05 * torch.randn((2500, 64, 64), device="cuda"))
    res_long = tp.block_amplitudes(long, tasa=500, mode=2, num_blocks=4)
    assert all(abs(f - 137 * 500 / 2500) < 1e-9 for f in res_long.f0)
    ref = to.mean_spectrum(long.cpu().numpy(), 4, 2)
    assert np.abs(res_long.mean_spectrum[2] - ref).max() <= 5e-6 * ref.max()
    m1000 = maps[:, :64, :64].repeat(2, 1, 1)[:1000].contiguous()
    mean, _ = tp.mean_spectra(m1000, m1000[0].clone(), 0.0, 32, 2, 2, tp.get_plan((64, 64), 1))
    ref = to.mean_spectrum(m1000.cpu().numpy(), 4, 3)
    assert np.abs(mean[3] - ref).max() <= 2e-6 * ref.max()
