"""Temporal harmonic analysis of a device-resident stack of height maps, all spatial blocks
at once.  Host side of csrc/fcd_temporal.cuh; replaces analyze.block_amplitude /
analyze.block_split (pydata/analyze.py:542-641, 365-417), which re-read every map file once
per block.

Frame-sharded stacks (one rank per GPU): the harmonic sums are additive over
frames, so every rank accumulates its own frames and the [bins][2][H*W] float64 sums are
all-reduced (NCCL) before the amplitudes are formed.  Estimating f0 needs whole time series
per pixel: the stack is redistributed from frame shards to row bands with one all-to-all
(`frames_to_row_bands`), each rank transforms its band's blocks, and the per-block f0 are
all-gathered."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch
from scipy.signal import find_peaks

from ._native import check
from .engine import HeightMapPlan, _ptr, _stream_ptr, get_plan, shard_range


def positive_frequencies(n_frames: int, tasa: float) -> np.ndarray:
    """np.fft.fftfreq(N, d=1/tasa) restricted to >= 0 (analyze.py:605-609)."""
    f = np.fft.fftfreq(n_frames, d=1 / tasa)
    return f[f >= 0]


def pick_f0(mean_spectrum: np.ndarray, fft_freqs: np.ndarray):
    """analyze.py:612-618: highest local maximum of the block's mean spectrum, or None."""
    if not np.all(np.isfinite(mean_spectrum)):
        return None
    peaks, _ = find_peaks(mean_spectrum)
    if len(peaks) == 0:
        return None
    return fft_freqs[peaks[np.argmax(mean_spectrum[peaks])]]


def harmonic_bins(f0, mode: int, fft_freqs: np.ndarray):
    """analyze.py:620-621 (the list starts at 0 * f0)."""
    harmonics = [f0 * n for n in range(0, mode)]
    return harmonics, [int(np.argmin(np.abs(fft_freqs - f))) for f in harmonics]


def frames_to_row_bands(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Frame shards -> row bands: rank r ends up with rows [r*H/ws, (r+1)*H/ws) of ALL n_total
    frames, in time order.  One all_to_all_single (NCCL over NVLink on GPUs; gloo in CPU tests)."""
    import torch.distributed as dist

    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    n_loc, H, W = local.shape
    if H % ws:
        raise ValueError("rows must divide evenly among the ranks")
    band = H // ws
    counts = [shard_range(n_total, r, ws)[1] - shard_range(n_total, r, ws)[0] for r in range(ws)]
    if counts[rank] != n_loc:
        raise ValueError("local frame count does not match shard_range")
    send = local.view(n_loc, ws, band, W).permute(1, 0, 2, 3).contiguous().view(ws * n_loc, band, W)   # [dst][frame]
    recv = torch.empty((n_total, band, W), dtype=local.dtype, device=local.device)
    dist.all_to_all_single(recv, send, output_split_sizes=counts, input_split_sizes=[n_loc] * ws, group=group)
    return recv


class BlockAmplitudes:
    """Result for the whole image: per-block f0 / harmonics, amplitude and phase planes
    [H, W, mode + 1] float64 (CUDA), mean spectra [blocks, npos] (when f0 was estimated)."""

    def __init__(self, harmonics, amps, phases, f0, mean_spectrum, fft_freqs, blocks_per_row, block_size):
        self.harmonics, self.amps, self.phases, self.f0 = harmonics, amps, phases, f0
        self.mean_spectrum, self.fft_freqs = mean_spectrum, fft_freqs
        self.blocks_per_row, self.block_size = blocks_per_row, block_size

    def block(self, block_index: int):
        """(harmonics, amps, phases, f0) of one block as numpy, like analyze.block_amplitude returns."""
        i, j = divmod(int(block_index), self.blocks_per_row)
        bs = self.block_size
        if self.f0[block_index] is None:                  # analyze.py:614-615
            mode = self.amps.shape[-1] - 1
            return (np.zeros(mode), np.full((bs, bs, mode), None, dtype=object),
                    np.full((bs, bs, mode), None, dtype=object), None, None)
        sl = (slice(i * bs, (i + 1) * bs), slice(j * bs, (j + 1) * bs))
        return (self.harmonics[block_index], self.amps[sl].cpu().numpy(), self.phases[sl].cpu().numpy(),
                self.f0[block_index])


def mean_spectra(maps: torch.Tensor, first_map: Optional[torch.Tensor], zero: float, block_size: int,
                 block_rows: int, block_cols: int, plan: HeightMapPlan):
    """nanmean(|fft(block stack)|) at the non-negative frequencies for every block of a
    device-resident stack [N, rows, cols] -> (numpy [blocks, npos], valid-pixel counts)."""
    n, rows, cols = (int(s) for s in maps.shape)
    if not plan.lib.fcd_temporal_frames_supported(n):
        raise ValueError(f"f0 estimation needs a frame count that is a power of two <= 4096, at most 2048, or a product of "
                         f"two factors <= 2048 (got {n}); drop a frame or pass f0 explicitly")
    nblk = block_rows * block_cols
    npos = n // 2 if n % 2 == 0 else (n + 1) // 2
    mean = np.zeros((nblk, npos), np.float64)
    valid = np.zeros(nblk, np.int32)
    with torch.cuda.device(plan.device):
        check(plan.lib, plan.lib.fcd_temporal_mean_spectrum(
            plan._h, _ptr(maps), n, rows, cols, _ptr(first_map), float(zero), int(block_size), int(block_rows),
            int(block_cols), mean.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            valid.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _stream_ptr()))
    return mean, valid


def block_amplitudes(maps: torch.Tensor, f0=None, tasa=500, mode=1, num_blocks=64, zero=0,
                     plan: Optional[HeightMapPlan] = None, group=None, n_total: Optional[int] = None,
                     first_map: Optional[torch.Tensor] = None, chunk_frames: int = 2048) -> BlockAmplitudes:
    """analyze.block_amplitude for every block of ``maps`` ([n, H, W] CUDA float32).  With
    ``group`` (torch.distributed) ``maps`` is this rank's contiguous frame shard
    (engine.shard_range) of an ``n_total``-frame series and every rank gets the full result."""
    if not (isinstance(maps, torch.Tensor) and maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 3):
        raise TypeError("maps must be a CUDA float32 tensor [n, H, W]")
    if not 1 <= mode <= 8:
        raise ValueError("mode must be in [1, 8]")
    maps = maps.contiguous()
    n, H, W = (int(s) for s in maps.shape)
    ws, rank, t0 = 1, 0, 0
    if group is not None:
        import torch.distributed as dist
        ws, rank = dist.get_world_size(group), dist.get_rank(group)
        if n_total is None:
            tot = torch.tensor([n], device=maps.device)
            dist.all_reduce(tot, group=group)
            n_total = int(tot.item())
        t0 = shard_range(n_total, rank, ws)[0]
    n_total = n if n_total is None else int(n_total)
    if plan is None:
        plan = get_plan((64, 64), 1, maps.device)          # only the launcher and the table caches are used
    bpr = int(np.sqrt(num_blocks))                          # analyze.py:572
    bs = H // bpr                                           # analyze.py:573
    nblk = bpr * bpr
    if first_map is None:                                   # the mask comes from the first map of the series
        first_map = maps[0].clone() if t0 == 0 and n > 0 else torch.empty((H, W), dtype=torch.float32, device=maps.device)
        if ws > 1:
            import torch.distributed as dist
            dist.broadcast(first_map, src=dist.get_global_rank(group, 0), group=group)
    fft_freqs = positive_frequencies(n_total, tasa)
    mean = None
    if f0 is None:
        if ws > 1:
            import torch.distributed as dist
            if bpr % ws or H % ws:
                raise ValueError("block rows must divide evenly among the ranks for f0 estimation")
            band = frames_to_row_bands(maps, n_total, group)                    # [n_total, H/ws, W]
            fm = first_map[rank * (H // ws):(rank + 1) * (H // ws)].contiguous()
            loc, _ = mean_spectra(band, fm, zero, bs, bpr // ws, bpr, plan)       # this band's block rows
            loc = torch.from_numpy(loc).to(maps.device)
            allm = [torch.empty_like(loc) for _ in range(ws)]
            dist.all_gather(allm, loc, group=group)
            mean = torch.cat(allm).cpu().numpy()
        else:
            mean, _ = mean_spectra(maps, first_map, zero, bs, bpr, bpr, plan)
        f0s = [pick_f0(mean[b], fft_freqs) for b in range(nblk)]
    else:
        f0s = [f0] * nblk
    harmonics, bins = [], np.zeros((nblk, mode), np.int32)
    for b in range(nblk):
        if f0s[b] is None:
            harmonics.append(None)
            continue
        h, idx = harmonic_bins(f0s[b], mode, fft_freqs)
        harmonics.append(h)
        bins[b] = idx
    acc = torch.zeros((mode, 2, H * W), dtype=torch.float64, device=maps.device)
    with torch.cuda.device(plan.device):
        for c0 in range(0, n, chunk_frames):
            c1 = min(n, c0 + chunk_frames)
            check(plan.lib, plan.lib.fcd_temporal_accumulate(
                plan._h, _ptr(maps[c0:c1]), c1 - c0, t0 + c0, n_total, H, W, float(zero), bs, bpr, bpr,
                bins.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), mode, _ptr(acc), int(c0 == 0), _stream_ptr()))
        if ws > 1:
            import torch.distributed as dist
            dist.all_reduce(acc, group=group)
        amps = torch.empty((H, W, mode + 1), dtype=torch.float64, device=maps.device)
        phases = torch.empty_like(amps)
        check(plan.lib, plan.lib.fcd_temporal_finalize(plan._h, _ptr(acc), mode, n_total, H, W, _ptr(first_map),
                                                       _ptr(amps), _ptr(phases), _stream_ptr()))
    return BlockAmplitudes(harmonics, amps, phases, f0s, mean, fft_freqs, bpr, bs)
