"""CPU oracle for the FCD height-map path -- TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy/scipy restatement of the reference algorithm in
``/root/reference/pyfcd`` (fcd.py, fourier.py, carriers.py).  It exists so that the
CUDA path can be checked against something independent.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it; the
product package (``trapped-modes-ltg_b200/``) never does.

Pinning status
--------------
* Everything that the reference itself computes (peak search, calibration factor,
  carrier construction, phase extraction, 2x2 solve, Fourier integration, layer
  height) is pinned against the *actual reference source* executed in the build
  container through ``oracle/ref_shims.py`` (which only substitutes the four
  scikit-image entry points and matplotlib, none of which are installed).  The
  resulting vectors are committed under ``tests/golden/`` by ``oracle/make_golden.py``.
* The one numeric golden held by the reference repo
  (``examples/Pictures/mask/maps/calibration_factor.npy``) is reproduced bit-exactly.
* Third-party pieces that are NOT under ``/root/reference`` and not installed
  (scikit-image, version unpinned by the reference: ``unwrap_phase``, ``label`` /
  ``regionprops``, ``draw.disk``) are restated from their published algorithms:
  **parity unpinned** for those three boundaries (no reference test or fixture
  exercises them in isolation).  ``unwrap_phase`` follows Herraez et al., Appl. Opt. 41
  (2002) 7437 (``oracle/unwrap_herraez.c``).

Every function cites the reference lines it restates.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import scipy.fft as sfft

_HERE = os.path.dirname(os.path.abspath(__file__))
TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------------------
# wavenumber helpers  (reference: pyfcd/fourier.py:44-73, 95-113)
# --------------------------------------------------------------------------------------
def wavenumber(size: int, calibration_factor: float = 1.0, shifted: bool = False) -> np.ndarray:
    """k = fftfreq(size, cal/2pi); optional fftshift.  fourier.py:44-57."""
    k = sfft.fftfreq(size, calibration_factor / TWO_PI)
    return sfft.fftshift(k) if shifted else k


def wavenumber_meshgrid(shape, calibration_factor: float = 1.0, shifted: bool = False):
    """'ij' meshes: first varies along rows, second along columns.  fourier.py:59-73."""
    kr = wavenumber(shape[0], calibration_factor, shifted)
    kc = wavenumber(shape[1], calibration_factor, shifted)
    return np.meshgrid(kr, kc, indexing="ij")


def pixel_to_wavenumber(shape, locations, calibration_factor: float = 1.0) -> np.ndarray:
    """Index the *shifted* wavenumber vectors at [row], [col].  fourier.py:95-113."""
    kr = wavenumber(shape[0], calibration_factor, shifted=True)
    kc = wavenumber(shape[1], calibration_factor, shifted=True)
    first = locations[0]
    if isinstance(first, np.ndarray):
        return np.array([[kr[p[0]], kc[p[1]]] for p in locations])
    return np.array([kr[locations[0]], kc[locations[1]]])


# --------------------------------------------------------------------------------------
# stand-ins for the scikit-image calls
# --------------------------------------------------------------------------------------
def label8(binary: np.ndarray) -> tuple[np.ndarray, int]:
    """8-connected component labelling, labels numbered by raster order of the first
    pixel of each component (the behaviour of skimage.measure.label with default
    connectivity, used at fourier.py:160).  Breadth-first flood from each seed; the
    inputs here hold a few dozen foreground pixels so a Python loop is fine."""
    binary = np.asarray(binary).astype(bool)
    lab = np.zeros(binary.shape, dtype=np.int32)
    n0, n1 = binary.shape
    current = 0
    for r, c in np.argwhere(binary):  # argwhere is row-major == raster order
        if lab[r, c]:
            continue
        current += 1
        lab[r, c] = current
        stack = [(int(r), int(c))]
        while stack:
            y, x = stack.pop()
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x + dx
                    if 0 <= yy < n0 and 0 <= xx < n1 and binary[yy, xx] and not lab[yy, xx]:
                        lab[yy, xx] = current
                        stack.append((yy, xx))
    return lab, current


def disk_mask(shape, center, radius) -> np.ndarray:
    """Boolean disk ((r-r0)/R)^2 + ((c-c0)/R)^2 < 1 clipped to ``shape`` -- the published
    behaviour of skimage.draw.disk(center, radius, shape=shape) used at carriers.py:19."""
    rr = (np.arange(shape[0], dtype=np.float64)[:, None] - float(center[0])) / float(radius)
    cc = (np.arange(shape[1], dtype=np.float64)[None, :] - float(center[1])) / float(radius)
    return (rr * rr + cc * cc) < 1.0


# -- Herraez unwrap (C) ----------------------------------------------------------------
_unwrap_lib = None


def _load_unwrap_lib():
    global _unwrap_lib
    if _unwrap_lib is not None:
        return _unwrap_lib
    so = os.path.join(_HERE, "libunwrap_herraez.so")
    src = os.path.join(_HERE, "unwrap_herraez.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    lib = ctypes.CDLL(so)
    lib.unwrap_herraez_2d.restype = ctypes.c_int
    lib.unwrap_herraez_2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    _unwrap_lib = lib
    return lib


def unwrap_phase(wrapped: np.ndarray) -> np.ndarray:
    """Reliability-sorted 2-D unwrapping (Herraez et al. 2002), the algorithm behind
    skimage.restoration.unwrap_phase called at fcd.py:119.  Output = input + 2*pi*integer."""
    w = np.ascontiguousarray(wrapped, dtype=np.float64)
    out = np.empty_like(w)
    lib = _load_unwrap_lib()
    rc = lib.unwrap_herraez_2d(w.ctypes.data, out.ctypes.data, w.shape[0], w.shape[1])
    if rc != 0:
        raise RuntimeError(f"unwrap_herraez_2d failed rc={rc}")
    return out


def unwrap_scan(wrapped: np.ndarray, row_ref: int | None = None, col_ref: int | None = None) -> np.ndarray:
    """Path-following unwrap (each row from ``col_ref``, rows linked along column
    ``col_ref`` from ``row_ref``).  Equal to any other correct unwrap, up to one global
    2*pi*k, iff the field has no residues.  This is the path the CUDA kernels follow."""
    w = np.asarray(wrapped, dtype=np.float64)
    n0, n1 = w.shape
    row_ref = n0 // 2 if row_ref is None else row_ref
    col_ref = n1 // 2 if col_ref is None else col_ref
    jumps = np.zeros((n0, n1), dtype=np.int64)
    jumps[:, 1:] = np.rint(np.diff(w, axis=1) / TWO_PI).astype(np.int64)
    c = np.cumsum(jumps, axis=1)
    c -= c[:, col_ref:col_ref + 1]
    col = w[:, col_ref]
    cj = np.zeros(n0, dtype=np.int64)
    cj[1:] = np.rint(np.diff(col) / TWO_PI).astype(np.int64)
    m = np.cumsum(cj)
    m -= m[row_ref]
    return w - TWO_PI * (c + m[:, None])


def count_residues(wrapped: np.ndarray) -> int:
    """Number of 2x2 loops whose wrapped differences do not sum to zero."""
    w = np.asarray(wrapped, dtype=np.float64)

    def wd(a):
        return a - TWO_PI * np.rint(a / TWO_PI)

    d1 = wd(w[:-1, 1:] - w[:-1, :-1])
    d2 = wd(w[1:, 1:] - w[:-1, 1:])
    d3 = wd(w[1:, :-1] - w[1:, 1:])
    d4 = wd(w[:-1, :-1] - w[1:, :-1])
    return int(np.count_nonzero(np.rint((d1 + d2 + d3 + d4) / TWO_PI)))


# --------------------------------------------------------------------------------------
# peak search  (reference: pyfcd/fourier.py:8-41, 140-168)
# --------------------------------------------------------------------------------------
def highpassed_spectrum(image: np.ndarray) -> np.ndarray:
    """fftshift(|fft2(image-mean)|) with bins k^2 <= (4pi/min(shape))^2 zeroed.
    fourier.py:18-23,34."""
    img = np.asarray(image)
    spec = sfft.fftshift(np.abs(sfft.fft2(img - np.mean(img))))
    kr, kc = wavenumber_meshgrid(spec.shape, shifted=True)
    kmin = 4.0 * np.pi / min(img.shape)
    return spec * ((kr ** 2 + kc ** 2) > kmin ** 2)


def find_peak_locations(image: np.ndarray, threshold: float, no_peaks: int):
    """Threshold, clear the border lines, 8-connected blobs, per-blob first-maximum pixel,
    stable ascending sort by that maximum, first ``no_peaks``.  fourier.py:140-168."""
    blob = np.array(image > threshold)
    blob[0, :] = False
    blob[-1, :] = False
    blob[:, 0] = False
    blob[:, -1] = False
    lab, n = label8(blob)
    found = []
    for l in range(1, n + 1):
        coords = np.argwhere(lab == l)  # row-major, like regionprops(...).coords
        vals = image[coords[:, 0], coords[:, 1]]
        j = int(np.argmax(vals))  # first maximum
        found.append((vals[j], coords[j]))
    found.sort(key=lambda t: t[0])  # stable
    return [c for _, c in found[:no_peaks]]


def find_peaks(image: np.ndarray):
    """(rightmost, perpendicular) carrier pixels in shifted coordinates.  fourier.py:8-41."""
    spec = highpassed_spectrum(image)
    thr = 0.5 * np.max(spec)
    locs = find_peak_locations(spec, thr, 4)

    def angle_key(p):
        k = pixel_to_wavenumber(spec.shape, p)
        return abs(np.arctan2(k[0], k[1]))

    rightmost = min(locs, key=angle_key)
    k_first = pixel_to_wavenumber(spec.shape, rightmost)

    def dot_key(p):
        return abs(np.dot(k_first, pixel_to_wavenumber(spec.shape, p)))

    perpendicular = min(locs, key=dot_key)
    return rightmost, perpendicular


# --------------------------------------------------------------------------------------
# carriers  (reference: pyfcd/carriers.py:9-24, pyfcd/fcd.py:54-101)
# --------------------------------------------------------------------------------------
@dataclass
class OracleCarrier:
    pixels: np.ndarray
    frequencies: np.ndarray
    radius: float
    mask: np.ndarray
    ccsgn: np.ndarray


def make_carrier(reference: np.ndarray, calibration_factor: float, peak, radius: float) -> OracleCarrier:
    """carriers.py:10-24: wavevector at the peak (with cal), disk drawn in shifted
    coordinates then ifftshift-ed, ccsgn = conj(ifft2(fft2(ref)*mask))."""
    ref = np.asarray(reference)
    freqs = pixel_to_wavenumber(ref.shape, peak, calibration_factor)
    mask = sfft.ifftshift(disk_mask(ref.shape, peak, radius))
    ccsgn = np.conj(sfft.ifft2(sfft.fft2(ref) * mask))
    return OracleCarrier(np.asarray(peak), freqs, float(radius), mask, ccsgn)


def compute_calibration_factor(square_size: float, reference: np.ndarray):
    """cal = 2*square_size / (2pi / mean|k components|).  fcd.py:73-101."""
    peaks = find_peaks(reference)
    k = pixel_to_wavenumber(np.shape(reference), peaks)
    pixel_wavelength = TWO_PI / np.mean(np.abs(k))
    return (2.0 * square_size) / pixel_wavelength, peaks


def compute_carriers(reference: np.ndarray, square_size: float):
    """fcd.py:54-70."""
    cal, peaks = compute_calibration_factor(square_size, reference)
    radius = np.linalg.norm(peaks[0] - peaks[1]) / 2.0
    return [make_carrier(reference, cal, p, radius) for p in peaks], cal


# --------------------------------------------------------------------------------------
# per-frame path  (reference: pyfcd/fcd.py:14-35, 104-138; pyfcd/fourier.py:76-92, 116-137)
# --------------------------------------------------------------------------------------
def height_from_layers(layers) -> float:
    """fcd.py:38-51 (note the hard-coded index 2 in effective_height)."""
    alpha = 1.0 - layers[-1][1] / layers[-2][1]
    total = 0
    for i in range(len(layers) - 1):
        total += layers[2][1] * (layers[i][0] / layers[i][1])
    return alpha * total


def compute_phases(displaced_fft: np.ndarray, carriers, unwrap=True, unwrapper=None) -> np.ndarray:
    """fcd.py:104-120.  ``unwrapper`` defaults to the Herraez restatement."""
    unwrapper = unwrap_phase if unwrapper is None else unwrapper
    out = np.zeros((2,) + displaced_fft.shape)
    for i, car in enumerate(carriers):
        ang = -np.angle(sfft.ifft2(displaced_fft * car.mask) * car.ccsgn)
        out[i] = unwrapper(ang) if unwrap else ang
    return out


def compute_displacement_field(phases: np.ndarray, carriers) -> np.ndarray:
    """fcd.py:123-138."""
    f0, f1 = carriers[0].frequencies, carriers[1].frequencies
    det = f0[1] * f1[0] - f0[0] * f1[1]
    u = (f1[0] * phases[0] - f0[0] * phases[1]) / det
    v = (f0[1] * phases[1] - f1[1] * phases[0]) / det
    return np.array([u, v])


def integrate_in_fourier(gx: np.ndarray, gy: np.ndarray, calibration_factor: float = 1.0) -> np.ndarray:
    """fourier.py:116-137 with remove_degeneracy (fourier.py:76-92): k2 formed first,
    k2[0,0]=1, then column N1//2+1 of kx and row N0//2+1 of ky zeroed (even sizes)."""
    ky, kx = wavenumber_meshgrid(gx.shape, calibration_factor)
    k2 = kx ** 2 + ky ** 2
    k2[0, 0] = 1
    n0, n1 = gx.shape
    if n1 % 2 == 0:
        kx[:, n1 // 2 + 1] = 0
    if n0 % 2 == 0:
        ky[n0 // 2 + 1, :] = 0
    gxh, gyh = sfft.fft2(gx), sfft.fft2(gy)
    hhat = (-1.0j * kx * gxh + -1.0j * ky * gyh) / k2
    return np.real(sfft.ifft2(hhat))


def resolve_height(layers=None, height=None) -> float:
    """fcd.py:16-25."""
    if height is not None:
        if layers is not None:
            raise Warning("Provide either height or layers, not both.")
        return height
    return 1 if layers is None else height_from_layers(layers)


def height_map_from_carriers(displaced, carriers, cal, height, unwrap=True, unwrapper=None):
    """Per-frame part of fcd.py:28-35 with the per-reference work hoisted."""
    dfft = sfft.fft2(np.asarray(displaced, dtype=np.float64))
    phases = compute_phases(dfft, carriers, unwrap, unwrapper)
    disp = compute_displacement_field(phases, carriers)
    grad = -disp / height
    return integrate_in_fourier(grad[0], grad[1], cal), phases


def compute_height_map(reference, displaced, square_size, layers=None, height=None, unwrap=True,
                       unwrapper=None):
    """fcd.py:14-35.  Inputs are upcast to float64 first so that this is a float64 oracle
    even for float32 frames (scipy would otherwise demodulate in single precision)."""
    h_eff = resolve_height(layers, height)
    ref = np.asarray(reference, dtype=np.float64)
    carriers, cal = compute_carriers(ref, square_size)
    hmap, phases = height_map_from_carriers(displaced, carriers, cal, h_eff, unwrap, unwrapper)
    return hmap, phases, cal


# --------------------------------------------------------------------------------------
# synthetic inputs  (SURVEY.md section 8(d)): shared with bench.py, which may not import oracle/ on its
# product arm -- the generators live in the package's input-generation module
# --------------------------------------------------------------------------------------
import importlib.util as _ilu
import os as _os

_spec = _ilu.spec_from_file_location(
    "fcd_b200_synthetic",
    _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "trapped-modes-ltg_b200", "fcd_b200",
                  "synthetic.py"))
_synthetic = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_synthetic)
rotated_board = _synthetic.rotated_board
board_square_size = _synthetic.board_square_size
gaussian_bump_displacement = _synthetic.gaussian_bump_displacement
synthetic_frames = _synthetic.synthetic_frames
