"""Host side of the B200 FCD path: plan objects, the batched device-resident API and frame
sharding.  PyTorch is used for device memory, streams and torch.distributed only; all
arithmetic on the hot path runs in the hand-written kernels behind the C ABI
(include/fcd_b200.h).

Reference call sites this replaces: pyfcd/fcd.py:14-35 (fcd.compute_height_map), looped
per frame by pydata/analyze.py:220-252.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native
from ._native import check, load_library

TWO_PI = 2.0 * np.pi


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("fcd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------------------
# scalar host logic mirrored from the reference
# ---------------------------------------------------------------------------------------
def height_from_layers(layers) -> float:
    """pyfcd/fcd.py:38-51 (effective_height uses the hard-coded index 2)."""
    fluid = layers[-2][1]
    before_camera = layers[-1][1]
    alpha = 1 - before_camera / fluid
    height = 0
    for i in range(len(layers) - 1):
        height += layers[2][1] * (layers[i][0] / layers[i][1])
    return alpha * height


def resolve_height(layers=None, height=None):
    """pyfcd/fcd.py:16-25, including the `raise Warning` when both are given."""
    if height is not None:
        if layers is None:
            return height
        raise Warning("Provide either height or layers, not both.")
    if layers is None:
        return 1
    return height_from_layers(layers)


def wavenumber(size, calibration_factor=1, shifted=False):
    """pyfcd/fourier.py:44-57: fftfreq(size, cal/2pi), optionally fftshift-ed."""
    n = int(size)
    val = 1.0 / (n * (calibration_factor / TWO_PI))
    m = np.empty(n, dtype=np.int64)
    half = (n - 1) // 2 + 1
    m[:half] = np.arange(0, half)
    m[half:] = np.arange(-(n // 2), 0)
    k = m * val
    return np.roll(k, n // 2) if shifted else k


def pixel_to_wavenumber(image_shape, locations, calibration_factor=1):
    """pyfcd/fourier.py:95-113."""
    kr = wavenumber(image_shape[0], calibration_factor, shifted=True)
    kc = wavenumber(image_shape[1], calibration_factor, shifted=True)
    if isinstance(locations[0], np.ndarray):
        return np.array([[kr[loc[0]], kc[loc[1]]] for loc in locations])
    return np.array([kr[locations[0]], kc[locations[1]]])


def calibration_from_peaks(shape, peaks, square_size) -> float:
    """pyfcd/fcd.py:86-88,101."""
    pixel_frequencies = pixel_to_wavenumber(shape, peaks)
    pixel_wavelength = 2 * np.pi / np.mean(np.abs(pixel_frequencies))
    physical_wavelength = 2 * square_size
    return physical_wavelength / pixel_wavelength


# ---------------------------------------------------------------------------------------
# tensors in / out
# ---------------------------------------------------------------------------------------
def to_device_image(a, device, allow_f64=True) -> torch.Tensor:
    """2-D (or 3-D batch) numpy / torch input -> contiguous CUDA tensor, float32 or float64."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64 if allow_f64 else np.float32)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64 if allow_f64 else torch.float32)
    if not allow_f64 and t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.to(device, non_blocking=True).contiguous()


UNWRAP_MODES = {"off": 0, "scan": 1, "herraez": 2, "guided": 2, "auto": 3}


def _mask_bytes(mask, device) -> torch.Tensor:
    """A mask as the kernels read it: one byte per pixel on `device`, zero = keep.  torch.bool storage is exactly
    that, so bool (what analyze.mask returns) and uint8 masks are reinterpreted, not copied: a dtype copy of a
    batch of 2048^2 masks costs 10 us per frame, more than K1 (profiles/launches_r02b_masked.csv)."""
    mk = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mask))
    mk = mk.to(device)
    if mk.dtype == torch.bool:
        mk = mk.contiguous().view(torch.uint8)
    elif mk.dtype != torch.uint8:
        mk = (mk != 0).view(torch.uint8)
    return mk.contiguous()


def unwrap_mode(unwrap) -> int:
    """False / 'off' -> 0;  True / 'scan' -> 1 (row/column scan: the fast path, exact where the
    wrapped phases have no residues);  'herraez' -> 2 (reliability-guided, what
    skimage.restoration.unwrap_phase does, pyfcd/fcd.py:119);  'auto' -> 3 (scan; the frames the
    demodulation kernel flags as able to wrap get a residue count, and those that hold residues are
    redone with 'herraez' -- all inside fcd_execute, include/fcd_b200.h)."""
    if isinstance(unwrap, str):
        try:
            return UNWRAP_MODES[unwrap]
        except KeyError:
            raise ValueError(f"unwrap must be a bool or one of {sorted(UNWRAP_MODES)}") from None
    return 1 if unwrap else 0


class HeightMapPlan:
    """One plan = one frame shape on one GPU.  Owns the workspaces, the twiddle tables and,
    after ``bind``, the per-reference state (carrier disks, ccsgn, integration coefficients).

    Not thread-safe; use from one CUDA stream at a time (the current torch stream)."""

    def __init__(self, shape: Sequence[int], frames_per_launch: int = 16, device=None):
        _require_cuda()
        self.lib = load_library()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (int(shape[0]), int(shape[1]))
        self.frames_per_launch = int(frames_per_launch)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_plan_create(self.shape[0], self.shape[1], self.frames_per_launch,
                                                     ctypes.byref(self._h)))
        self.fused = bool(self.lib.fcd_plan_is_fused(self._h))   # False: a shape that is not a power of two (generic.py)
        self._generic = None
        self.peaks = None
        self.radius = None
        self.calibration_factor = None
        self.height = None
        self._reference = None
        self._dropin_key = None          # fcd.compute_height_map's reference cache (cleared by every bind)
        self.last_guided_frames = []     # frames the last unwrap="auto" call redid reliability-guided
        self.last_flagged_frames = 0     # frames that call looked at twice (|phase| > pi/2 somewhere)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.fcd_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- carrier search ------------------------------------------------------------------
    def _check_image(self, t: torch.Tensor) -> None:
        if tuple(t.shape[-2:]) != self.shape:
            raise ValueError(f"image shape {tuple(t.shape)} does not match plan shape {self.shape}")

    def highpass_spectrum(self, image, return_spectrum=True):
        """fftshift(|fft2(image-mean)|) with the low-k disc zeroed, and its max (fourier.py:18-35)."""
        img = to_device_image(image, self.device)
        self._check_image(img)
        spec = torch.empty(self.shape, dtype=torch.float64, device=self.device) if return_spectrum else None
        mx = ctypes.c_double(0.0)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_highpass_spectrum(self._h, _ptr(img), int(img.dtype == torch.float64),
                                                           _ptr(spec), ctypes.byref(mx), _stream_ptr()))
        return spec, mx.value

    def peak_locations(self, image, threshold, no_peaks):
        """fourier.find_peak_locations (fourier.py:140-168) on a float64 device image."""
        img = to_device_image(image, self.device).to(torch.float64).contiguous()
        self._check_image(img)
        rc = (ctypes.c_int * (2 * max(int(no_peaks), 1)))()
        count = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_peak_locations(self._h, _ptr(img), float(threshold), int(no_peaks), rc,
                                                        ctypes.byref(count), _stream_ptr()))
        return [np.array([rc[2 * i], rc[2 * i + 1]]) for i in range(count.value)]

    def find_peaks(self, image):
        """(rightmost, perpendicular) carrier pixels, shifted coordinates (fourier.py:8-41)."""
        img = to_device_image(image, self.device)
        self._check_image(img)
        out = (ctypes.c_int * 4)()
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_find_peaks(self._h, _ptr(img), int(img.dtype == torch.float64), out,
                                                    _stream_ptr()))
        return np.array([out[0], out[1]]), np.array([out[2], out[3]])

    # ---- per-reference state -------------------------------------------------------------
    def bind(self, reference, square_size=None, layers=None, height=None, peaks=None, radius=None,
             calibration_factor=None, allow_collinear=False):
        """Carrier detection (unless ``peaks`` are given) + per-reference precomputation.
        Mirrors fcd.compute_carriers (fcd.py:54-70).  Returns the calibration factor."""
        ref = to_device_image(reference, self.device)
        self._check_image(ref)
        self._dropin_key = None          # whatever was cached for the drop-in is gone from here on
        h_eff = resolve_height(layers, height)
        if peaks is None:
            peaks = self.find_peaks(ref)
        peaks = (np.asarray(peaks[0]), np.asarray(peaks[1]))
        if calibration_factor is None:
            if square_size is None:
                raise ValueError("square_size or calibration_factor is required")
            calibration_factor = calibration_from_peaks(self.shape, peaks, square_size)
        if radius is None:
            radius = float(np.linalg.norm(peaks[0] - peaks[1]) / 2)
        pk = (ctypes.c_int * 4)(int(peaks[0][0]), int(peaks[0][1]), int(peaks[1][0]), int(peaks[1][1]))
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_bind_reference(self._h, _ptr(ref), int(ref.dtype == torch.float64), pk,
                                                        float(radius), float(calibration_factor), float(h_eff),
                                                        _stream_ptr()))
        self.peaks, self.radius = peaks, float(radius)
        self.calibration_factor, self.height = float(calibration_factor), h_eff
        self._reference = ref
        self._generic = None
        return self.calibration_factor

    def carrier_frequencies(self):
        """[k_row, k_col] of both carriers in calibrated units (carriers.py:12)."""
        return [pixel_to_wavenumber(self.shape, p, self.calibration_factor) for p in self.peaks]

    def carrier_mask(self, i: int) -> torch.Tensor:
        out = torch.empty(self.shape, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_get_carrier_mask(self._h, int(i), _ptr(out), _stream_ptr()))
        return out.bool()

    def carrier_ccsgn(self, i: int, complex128: bool = True) -> torch.Tensor:
        out = torch.empty(self.shape, dtype=torch.complex128 if complex128 else torch.complex64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_get_carrier_ccsgn(self._h, int(i), _ptr(out), int(complex128), _stream_ptr()))
        return out

    # ---- per-frame path ------------------------------------------------------------------
    def execute(self, frames: torch.Tensor, out: Optional[torch.Tensor] = None, phases=False,
                mask: Optional[torch.Tensor] = None, unwrap=True):
        """frames: CUDA float32 [n, H, W] (or [H, W]).  Returns height maps float32 [n, H, W]
        and, when ``phases`` is True or a tensor, phases float32 [n, 2, H, W].
        ``unwrap``: see :func:`unwrap_mode`."""
        mode = unwrap_mode(unwrap)
        kinds = {torch.float32: 0, torch.uint8: 1, torch.uint16: 2}
        if not self.fused:
            kinds[torch.float64] = 3          # the float64 path of generic plans takes float64 frames as they are
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype in kinds):
            raise TypeError("frames must be a CUDA float32 / uint8 / uint16 tensor "
                            "(see compute_height_maps for numpy input)")
        if frames.device != self.device:
            raise ValueError(f"frames are on {frames.device}, the plan is on {self.device}")
        kind = kinds[frames.dtype]
        squeeze = frames.dim() == 2
        fr = frames.unsqueeze(0) if squeeze else frames
        fr = fr.contiguous()
        self._check_image(fr)
        n = fr.shape[0]
        # generic plans (shapes that are not powers of two) compute in float64 and hand out float64 by default
        odt = (torch.float32,) if self.fused else (torch.float64, torch.float32)
        if out is None:
            out = torch.empty(fr.shape, dtype=odt[0], device=fr.device)
        elif not (out.is_cuda and out.device == self.device and out.dtype in odt and out.is_contiguous()
                  and out.numel() == fr.numel()):
            raise ValueError("out must be a contiguous CUDA float32 tensor of the frames' size on the plan's device")
        ph = None
        if isinstance(phases, torch.Tensor):
            ph = phases
            if not (ph.is_cuda and ph.device == self.device and ph.dtype in odt and ph.is_contiguous()
                    and tuple(ph.shape) == (n, 2) + self.shape):
                raise ValueError("phases must be a contiguous CUDA float32 tensor [n, 2, H, W] on the plan's device")
        elif phases:
            ph = torch.empty((n, 2) + self.shape, dtype=odt[0], device=self.device)
        mask_stride = 0
        mk = None
        if mask is not None:
            mk = _mask_bytes(mask, self.device)
            if mk.dim() == 3:
                if mk.shape[0] != n:
                    raise ValueError("per-frame mask count differs from frame count")
                mask_stride = self.shape[0] * self.shape[1]
            self._check_image(mk)
        if not self.fused:
            if self.peaks is None:
                raise _native.FcdError(_native.FCD_ERR_STATE, "execute called before bind")
            if self._generic is None:
                from .generic import GenericPipeline
                self._generic = GenericPipeline(self)
            with torch.cuda.device(self.device):
                guided = self._generic.execute(fr, out.view((n,) + self.shape), ph, mk, mode)
            self.last_guided_frames, self.last_flagged_frames = guided, len(guided)
            if squeeze:
                out = out.view(self.shape)
                ph = ph[0] if ph is not None else None
            return (out, ph) if ph is not None else out
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_execute_typed(self._h, _ptr(fr), kind, int(n), _ptr(out), _ptr(ph),
                                                       _ptr(mk), int(mask_stride), mode, _stream_ptr()))
        if mode == 3:
            flagged, count = ctypes.c_longlong(0), ctypes.c_int(0)
            idx = (ctypes.c_int * max(int(n), 1))()
            check(self.lib, self.lib.fcd_last_auto(self._h, ctypes.byref(flagged), ctypes.byref(count), idx, int(n)))
            self.last_flagged_frames = int(flagged.value)
            self.last_guided_frames = [idx[i] for i in range(count.value)]
        if squeeze:
            out = out.view(self.shape) if out.dim() == 3 else out
            ph = ph[0] if ph is not None else None
        return (out, ph) if ph is not None else out

    def unwrap_phase(self, wrapped: torch.Tensor) -> torch.Tensor:
        """Reliability-guided unwrap of [..., H, W] float32 maps (skimage.restoration.unwrap_phase as
        called at pyfcd/fcd.py:119); pixel (0, 0) of every map keeps its value."""
        w = wrapped.to(self.device).to(torch.float32).contiguous()
        self._check_image(w)
        n = w.numel() // (self.shape[0] * self.shape[1])
        out = torch.empty_like(w)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_unwrap_phase(self._h, _ptr(w), int(n), _ptr(out), _stream_ptr()))
        return out

    def set_height(self, layers=None, height=None) -> None:
        """Change the effective height of the bound reference (fcd.py:16-25) without re-binding."""
        h_eff = resolve_height(layers, height)
        check(self.lib, self.lib.fcd_set_height(self._h, float(h_eff)))
        self.height = h_eff

    def count_residues(self, phases: torch.Tensor) -> list:
        """Residues per phase map ([..., H, W] CUDA float32, e.g. wrapped phases from
        execute(..., phases=True, unwrap=False)); unwrap parity with the reference holds where 0."""
        ph = phases.to(self.device).to(torch.float32).contiguous()
        self._check_image(ph)
        n = ph.numel() // (self.shape[0] * self.shape[1])
        out = (ctypes.c_int * max(n, 1))()
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_count_residues(self._h, _ptr(ph), int(n), out, _stream_ptr()))
        return [out[i] for i in range(n)]

    # ---- floating-structure mask / cavity centre (pydata/analyze.py:43-140) ---------------------
    def structure_mask(self, frames, smoothed: int = 14) -> torch.Tensor:
        """analyze.mask for a batch: frames [n,H,W] (or [H,W]) float32 -> bool masks, bit-exact."""
        fr = to_device_image(frames, self.device, allow_f64=False)
        squeeze = fr.dim() == 2
        fr = (fr.unsqueeze(0) if squeeze else fr).contiguous()
        self._check_image(fr)
        out = torch.empty(fr.shape, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_structure_mask(self._h, _ptr(fr), int(fr.shape[0]), int(smoothed), _ptr(out),
                                                        _stream_ptr()))
        out = out.view(torch.bool)          # the kernels write 0 / 1 bytes: reinterpreted, not copied
        return out[0] if squeeze else out

    def mask_center(self, masks) -> list:
        """analyze.center for a batch of masks: [(cy, cx), ...]; (-1, -1) where there is no enclosed region."""
        mk = _mask_bytes(masks, self.device)
        mk = mk.unsqueeze(0) if mk.dim() == 2 else mk
        self._check_image(mk)
        n = int(mk.shape[0])
        out = (ctypes.c_int * (2 * max(n, 1)))()
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_mask_center(self._h, _ptr(mk), n, out, _stream_ptr()))
        return [(out[2 * i], out[2 * i + 1]) for i in range(n)]

    STAGES = ("row_fwd", "col_band", "row_demod", "row_link", "phase_fix", "col_integrate", "row_inv")

    def set_profiling(self, enable: bool) -> None:
        check(self.lib, self.lib.fcd_set_profiling(self._h, int(bool(enable))))

    def stage_times(self) -> dict:
        """{stage: (milliseconds, launches, frames)} accumulated since set_profiling(True)."""
        ms = (ctypes.c_double * 7)()
        n = (ctypes.c_longlong * 7)()
        fr = (ctypes.c_longlong * 7)()
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_stage_times(self._h, ms, n, fr))
        return {name: (ms[i], n[i], fr[i]) for i, name in enumerate(self.STAGES)}

    @property
    def launch_count(self) -> int:
        return int(self.lib.fcd_launch_count(self._h))

    @property
    def band_columns(self) -> int:
        return int(self.lib.fcd_band_columns(self._h))

    def fft2_c128(self, x: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """Stage-level float64 2-D FFT (scipy.fft.fft2 / ifft2 semantics) on the hand-written kernels."""
        x = x.to(self.device).to(torch.complex128).contiguous()
        self._check_image(x)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            check(self.lib, self.lib.fcd_fft2_c128(self._h, _ptr(x), _ptr(out), 1 if inverse else -1, _stream_ptr()))
        return out


# ---------------------------------------------------------------------------------------
# batched public API
# ---------------------------------------------------------------------------------------
_plan_cache: dict = {}


def get_plan(shape, frames_per_launch: int = 16, device=None) -> HeightMapPlan:
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (int(shape[0]), int(shape[1]), int(frames_per_launch), dev.index)
    plan = _plan_cache.get(key)
    if plan is None:
        plan = HeightMapPlan(shape, frames_per_launch, dev)
        _plan_cache[key] = plan
    return plan


def compute_height_maps(reference, frames, square_size, layers=None, height=None, unwrap="auto",
                        return_phases=False, mask=None, plan: Optional[HeightMapPlan] = None,
                        frames_per_launch: int = 16, out=None):
    """Batched equivalent of calling fcd.compute_height_map(reference, frame, ...) for every
    frame (pydata/analyze.py:220-252) with the per-reference work done once.

    ``frames``: [n, H, W] numpy array (copied host->device) or CUDA float32 tensor (used in
    place).  Returns (height_maps [n,H,W] CUDA float32, phases [n,2,H,W] or None,
    calibration_factor)."""
    _require_cuda()
    shape = tuple(np.shape(reference))
    if plan is None:
        plan = get_plan(shape, frames_per_launch)
    cal = plan.bind(reference, square_size=square_size, layers=layers, height=height)
    if isinstance(frames, torch.Tensor) and frames.is_cuda:
        fr = frames if frames.dtype in (torch.uint8, torch.uint16) else frames.to(torch.float32)
    elif isinstance(frames, np.ndarray) and frames.dtype in (np.uint8, np.uint16):
        fr = torch.from_numpy(np.ascontiguousarray(frames)).to(plan.device)    # widened on the GPU
    else:
        fr = to_device_image(frames, plan.device, allow_f64=not plan.fused)
    res = plan.execute(fr, out=out, phases=return_phases, mask=mask, unwrap=unwrap)
    if return_phases:
        return res[0], res[1], cal
    return res, None, cal


# ---------------------------------------------------------------------------------------
# frame sharding across GPUs: contiguous ranges, no hot-path collective
# ---------------------------------------------------------------------------------------
def shard_range(n_frames: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [start, stop) of rank's frames; the first n_frames % world_size ranks get one more."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(n_frames), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_height_maps(local: torch.Tensor, n_frames: int, dst: int = 0, group=None, chunk_frames: int = 64):
    """Optional final gather of the sharded height maps onto ``dst`` (NCCL over NVLink, or
    gloo on CPU tensors in tests), in chunks so that no rank stages more than ``chunk_frames``
    frames per peer at a time.  Returns the full [n_frames, H, W] tensor on ``dst``, else None."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ranges = [shard_range(n_frames, r, world) for r in range(world)]
    full = None
    if rank == dst:
        full = torch.empty((n_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        a, b = ranges[dst]
        full[a:b].copy_(local)
    for r in range(world):
        if r == dst:
            continue
        a, b = ranges[r]
        for c0 in range(a, b, chunk_frames):
            c1 = min(b, c0 + chunk_frames)
            if rank == r:
                dist.send(local[c0 - a:c1 - a].contiguous(), dst=dst, group=group)
            elif rank == dst:
                dist.recv(full[c0:c1], src=r, group=group)
    return full
