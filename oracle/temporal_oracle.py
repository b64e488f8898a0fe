"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's temporal harmonic analysis,
``analyze.block_amplitude`` (pydata/analyze.py:542-641) and ``analyze.block_split``
(pydata/analyze.py:365-417), on an in-memory stack instead of a folder of ``*_map.npy`` files.

Pinned: ``oracle/make_golden_temporal.py`` runs the UNMODIFIED reference on a temporary folder
holding the same maps and stores its outputs in ``tests/golden/golden_temporal.npz``; this
restatement reproduces them exactly (tests/test_temporal.py).  Only ``tests/`` and ``bench.py``'s
CPU legs may import this module."""
from __future__ import annotations

import numpy as np
from scipy.signal import find_peaks


def synthetic_maps(n_frames: int, shape=(64, 64), num_blocks: int = 4, tasa: float = 500.0, seed: int = 1,
                   dtype=np.float32) -> np.ndarray:
    """Standing-wave height maps: every spatial block oscillates at its own frequency (plus a
    weaker second harmonic, a static offset and noise); a few pixels of the first map are
    exactly zero (the reference treats those as masked, analyze.py:568)."""
    rng = np.random.default_rng(seed)
    H, W = shape
    bpr = int(np.sqrt(num_blocks))
    bs = H // bpr
    y, x = np.mgrid[0:H, 0:W]
    t = np.arange(n_frames) / tasa
    maps = np.zeros((n_frames, H, W))
    for b in range(bpr * bpr):
        i, j = divmod(b, bpr)
        cyc = 5 + 3 * b                                   # whole cycles per record -> exact bin
        f = cyc * tasa / n_frames
        sl = (slice(i * bs, (i + 1) * bs), slice(j * bs, (j + 1) * bs))
        shape_b = np.cos(np.pi * (y[sl] - i * bs) / bs) * np.cos(2 * np.pi * (x[sl] - j * bs) / bs)
        ph = 0.3 * b
        maps[(slice(None),) + sl] = (0.02 * (b + 1) + shape_b[None] * np.cos(2 * np.pi * f * t + ph)[:, None, None]
                                     + 0.3 * shape_b[None] ** 2 * np.cos(4 * np.pi * f * t + 1.0)[:, None, None])
    maps += 0.05 * rng.standard_normal(maps.shape)
    maps = maps.astype(dtype)
    maps[0, 5:9, 3:12] = 0.0                              # masked pixels (zero in the first map)
    maps[:, H - 4:, W - 6:] = 0.0
    return maps


def block_split(maps: np.ndarray, t_limit=None, num_blocks: int = 64, block_index: int = 0) -> np.ndarray:
    """analyze.py:365-417 on a stack [N, H, W] (files in sorted order)."""
    maps = maps[:t_limit]
    initial_map = maps[0]
    H, W = initial_map.shape
    mask_validos = ~(initial_map == 0)
    blocks_per_row = int(np.sqrt(num_blocks))
    block_size = H // blocks_per_row
    i = block_index // blocks_per_row
    j = block_index % blocks_per_row
    out = []
    for m in maps:
        block = m[i * block_size:(i + 1) * block_size, j * block_size:(j + 1) * block_size]
        mask_block = mask_validos[i * block_size:(i + 1) * block_size, j * block_size:(j + 1) * block_size]
        out.append(np.where(mask_block, block, np.nan))
    return np.transpose(np.stack(out, axis=0), (1, 2, 0))


def block_amplitude(maps: np.ndarray, f0=None, tasa=500, mode=1, num_blocks=64, block_index=0, zero=0):
    """analyze.py:542-641 on a stack [N, H, W].  Returns (harmonics, amps, phases, f0) like the
    reference (its docstring promises five values; the code returns these four)."""
    initial_map = maps[0]
    H, W = initial_map.shape
    mask_validos = ~(initial_map == 0)
    blocks_per_row = int(np.sqrt(num_blocks))
    block_size = H // blocks_per_row
    i = block_index // blocks_per_row
    j = block_index % blocks_per_row
    stack = []
    for m0 in maps:
        m = m0 - zero
        block = m[i * block_size:(i + 1) * block_size, j * block_size:(j + 1) * block_size]
        mask_block = mask_validos[i * block_size:(i + 1) * block_size, j * block_size:(j + 1) * block_size]
        stack.append(np.where(mask_block, block, np.nan))
    stack = np.transpose(np.stack(stack, axis=0), (1, 2, 0))
    ny, nx, N = stack.shape
    dt = 1 / tasa
    fft_vals = np.fft.fft(stack, axis=-1)
    fft_freqs = np.fft.fftfreq(N, d=dt)
    pos_freqs = fft_freqs >= 0
    fft_vals = fft_vals[:, :, pos_freqs]
    fft_freqs = fft_freqs[pos_freqs]
    if f0 is None:
        with np.errstate(all="ignore"):
            mean_spectrum = np.nanmean(np.abs(fft_vals), axis=(0, 1))
        peaks, _ = find_peaks(mean_spectrum)
        if len(peaks) == 0:
            return (np.zeros(mode), np.full((ny, nx, mode), None, dtype=object),
                    np.full((ny, nx, mode), None, dtype=object), None, None)
        max_peak_index = peaks[np.argmax(mean_spectrum[peaks])]
        f0 = fft_freqs[max_peak_index]
    harmonics = [f0 * n for n in range(0, mode)]
    indices = [np.argmin(np.abs(fft_freqs - f)) for f in harmonics]
    amps = np.zeros((ny, nx, mode + 1))
    phases = np.zeros((ny, nx, mode + 1))
    for k, idx in enumerate(indices):
        harmonic_vals = fft_vals[:, :, idx]
        if k == 0:
            amps[:, :, k] = np.abs(harmonic_vals) / N
        else:
            amps[:, :, k] = 2 * np.abs(harmonic_vals) / N
        phases[:, :, k] = np.angle(harmonic_vals)
    return harmonics, amps, phases, f0


def mean_spectrum(maps: np.ndarray, num_blocks=64, block_index=0, zero=0):
    """The quantity f0 is estimated from (analyze.py:603-614) and its frequency axis."""
    stack = block_split(maps - zero if zero else maps, None, num_blocks, block_index)
    if zero:                                   # the mask comes from the raw first map (analyze.py:568)
        stack = block_split(maps, None, num_blocks, block_index) - zero
    N = stack.shape[-1]
    fft_vals = np.fft.fft(stack, axis=-1)
    pos = np.fft.fftfreq(N) >= 0
    with np.errstate(all="ignore"):
        return np.nanmean(np.abs(fft_vals[:, :, pos]), axis=(0, 1))
