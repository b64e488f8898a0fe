"""TEST INFRASTRUCTURE ONLY: builds and wraps tests/emul/libfcd_emul.so -- the product's
kernel sources compiled with -DFCD_EMULATE so that every block / phase / thread runs
sequentially on the CPU ("device" pointers are host pointers).  It lets the CPU test-suite
check the exact kernel arithmetic and index maps in a container without a GPU.  The product
package never imports this."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "trapped-modes-ltg_b200")
CSRC = os.path.join(PKG, "csrc")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
LIB = os.path.join(EMUL_DIR, "libfcd_emul.so")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from fcd_b200 import _native  # noqa: E402  (prototypes only; does not load the CUDA library)


def build_emul() -> str:
    srcs = [os.path.join(EMUL_DIR, "fcd_emul.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs.append(os.path.join(ROOT, "include", "fcd_b200.h"))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-DFCD_EMULATE", "-shared", "-fPIC",
                               "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB,
                               os.path.join(EMUL_DIR, "fcd_emul.cpp")])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _native.declare(ctypes.CDLL(build_emul()))
    return _lib


def _p(a):
    return ctypes.c_void_p(0 if a is None else a.ctypes.data)


class EmulPlan:
    """numpy-buffer twin of fcd_b200.engine.HeightMapPlan on the emulation library."""

    def __init__(self, shape, frames_per_launch=2):
        self.lib = lib()
        self.shape = tuple(int(s) for s in shape)
        self.h = ctypes.c_void_p()
        _native.check(self.lib, self.lib.fcd_plan_create(self.shape[0], self.shape[1], frames_per_launch, ctypes.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.fcd_plan_destroy(self.h)
            self.h = ctypes.c_void_p()

    def find_peaks(self, image):
        img = np.ascontiguousarray(image)
        assert img.dtype in (np.float32, np.float64)
        out = (ctypes.c_int * 4)()
        _native.check(self.lib, self.lib.fcd_find_peaks(self.h, _p(img), int(img.dtype == np.float64), out, None))
        return np.array([out[0], out[1]]), np.array([out[2], out[3]])

    def highpass_spectrum(self, image):
        img = np.ascontiguousarray(image)
        spec = np.empty(self.shape, np.float64)
        mx = ctypes.c_double()
        _native.check(self.lib, self.lib.fcd_highpass_spectrum(self.h, _p(img), int(img.dtype == np.float64), _p(spec),
                                                               ctypes.byref(mx), None))
        return spec, mx.value

    def peak_locations(self, image, threshold, no_peaks):
        img = np.ascontiguousarray(image, dtype=np.float64)
        rc = (ctypes.c_int * (2 * max(no_peaks, 1)))()
        cnt = ctypes.c_int()
        _native.check(self.lib, self.lib.fcd_peak_locations(self.h, _p(img), float(threshold), no_peaks, rc,
                                                            ctypes.byref(cnt), None))
        return [np.array([rc[2 * i], rc[2 * i + 1]]) for i in range(cnt.value)]

    def bind(self, reference, peaks, radius, cal, height=1.0):
        ref = np.ascontiguousarray(reference)
        pk = (ctypes.c_int * 4)(int(peaks[0][0]), int(peaks[0][1]), int(peaks[1][0]), int(peaks[1][1]))
        _native.check(self.lib, self.lib.fcd_bind_reference(self.h, _p(ref), int(ref.dtype == np.float64), pk,
                                                            float(radius), float(cal), float(height), None))

    def execute(self, frames, phases=False, mask=None, unwrap=True):
        fr = np.ascontiguousarray(frames)
        if fr.dtype not in (np.uint8, np.uint16):
            fr = np.ascontiguousarray(fr, dtype=np.float32)
        kind = {np.dtype(np.float32): 0, np.dtype(np.uint8): 1, np.dtype(np.uint16): 2}[fr.dtype]
        if fr.ndim == 2:
            fr = fr[None]
        out = np.zeros(fr.shape, np.float32)
        ph = np.zeros((fr.shape[0], 2) + self.shape, np.float32) if phases else None
        mk, stride = None, 0
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.uint8)
            stride = self.shape[0] * self.shape[1] if mk.ndim == 3 else 0
        _native.check(self.lib, self.lib.fcd_execute_typed(self.h, _p(fr), kind, fr.shape[0], _p(out), _p(ph), _p(mk),
                                                           stride, int(unwrap), None))
        return (out, ph) if phases else out

    def last_auto(self):
        """(frames flagged for a second look, indices of the frames redone reliability-guided) of the last unwrap=3 call."""
        flagged, count = ctypes.c_longlong(0), ctypes.c_int(0)
        idx = (ctypes.c_int * 4096)()
        _native.check(self.lib, self.lib.fcd_last_auto(self.h, ctypes.byref(flagged), ctypes.byref(count), idx, 4096))
        return int(flagged.value), [idx[i] for i in range(count.value)]

    def unwrap_phase(self, wrapped):
        w = np.ascontiguousarray(wrapped, dtype=np.float32)
        n = w.size // (self.shape[0] * self.shape[1])
        out = np.empty_like(w)
        _native.check(self.lib, self.lib.fcd_unwrap_phase(self.h, _p(w), n, _p(out), None))
        return out

    def temporal_mean_spectrum(self, maps, first, zero, bpr, n1=0):
        maps = np.ascontiguousarray(maps, dtype=np.float32)
        n, rows, cols = maps.shape
        first = None if first is None else np.ascontiguousarray(first, dtype=np.float32)
        npos = n // 2 if n % 2 == 0 else (n + 1) // 2
        mean = np.zeros((bpr * bpr, npos), np.float64)
        valid = np.zeros(bpr * bpr, np.int32)
        mp_, vp_ = mean.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), valid.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
        if n1:
            _native.check(self.lib, self.lib.fcd_temporal_mean_spectrum_split(
                self.h, _p(maps), n, rows, cols, _p(first), float(zero), rows // bpr, bpr, bpr, int(n1), mp_, vp_, None))
        else:
            _native.check(self.lib, self.lib.fcd_temporal_mean_spectrum(
                self.h, _p(maps), n, rows, cols, _p(first), float(zero), rows // bpr, bpr, bpr, mp_, vp_, None))
        return mean, valid

    def temporal_harmonics(self, maps, bins, n_total=None, t0=0, zero=0.0, bpr=2, first=None, chunks=(None,)):
        maps = np.ascontiguousarray(maps, dtype=np.float32)
        n, rows, cols = maps.shape
        n_total = n if n_total is None else n_total
        bins = np.ascontiguousarray(bins, dtype=np.int32)
        nb = bins.shape[1]
        acc = np.zeros((nb, 2, rows * cols), np.float64)
        edges = [0] + [c for c in chunks if c is not None] + [n]
        for a, b in zip(edges[:-1], edges[1:]):
            _native.check(self.lib, self.lib.fcd_temporal_accumulate(
                self.h, _p(maps[a:b]), b - a, t0 + a, n_total, rows, cols, float(zero), rows // bpr, bpr, bpr,
                bins.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), nb, _p(acc), int(a == 0), None))
        return acc

    def temporal_finalize(self, acc, n_total, shape, first=None):
        nb = acc.shape[0]
        first = None if first is None else np.ascontiguousarray(first, dtype=np.float32)
        amps = np.zeros(tuple(shape) + (nb + 1,), np.float64)
        phases = np.zeros_like(amps)
        _native.check(self.lib, self.lib.fcd_temporal_finalize(self.h, _p(acc), nb, n_total, shape[0], shape[1], _p(first),
                                                               _p(amps), _p(phases), None))
        return amps, phases

    def set_height(self, height):
        _native.check(self.lib, self.lib.fcd_set_height(self.h, float(height)))

    def count_residues(self, phases):
        ph = np.ascontiguousarray(phases, dtype=np.float32)
        n = ph.size // (self.shape[0] * self.shape[1])
        out = (ctypes.c_int * max(n, 1))()
        _native.check(self.lib, self.lib.fcd_count_residues(self.h, _p(ph), n, out, None))
        return [out[i] for i in range(n)]

    def structure_mask(self, frames, smoothed=14):
        fr = np.ascontiguousarray(frames, dtype=np.float32)
        if fr.ndim == 2:
            fr = fr[None]
        out = np.zeros(fr.shape, np.uint8)
        _native.check(self.lib, self.lib.fcd_structure_mask(self.h, _p(fr), fr.shape[0], int(smoothed), _p(out), None))
        return out.astype(bool)

    def mask_center(self, masks):
        mk = np.ascontiguousarray(masks, dtype=np.uint8)
        if mk.ndim == 2:
            mk = mk[None]
        out = (ctypes.c_int * (2 * mk.shape[0]))()
        _native.check(self.lib, self.lib.fcd_mask_center(self.h, _p(mk), mk.shape[0], out, None))
        return [(out[2 * i], out[2 * i + 1]) for i in range(mk.shape[0])]

    def mask(self, i):
        m = np.zeros(self.shape, np.uint8)
        _native.check(self.lib, self.lib.fcd_get_carrier_mask(self.h, i, _p(m), None))
        return m.astype(bool)

    def ccsgn(self, i, c128=True):
        c = np.zeros(self.shape, np.complex128 if c128 else np.complex64)
        _native.check(self.lib, self.lib.fcd_get_carrier_ccsgn(self.h, i, _p(c), int(c128), None))
        return c

    def fft2(self, x, inverse=False):
        a = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(a)
        _native.check(self.lib, self.lib.fcd_fft2_c128(self.h, _p(a), _p(out), 1 if inverse else -1, None))
        return out
