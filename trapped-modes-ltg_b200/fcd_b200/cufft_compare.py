"""Per-stage cuFFT comparator (BENCHMARK UTILITY ONLY; csrc/cufft_compare.cu).

`north_star` asks that each hand-written kernel be compared with a cuFFT-based version of the same stage.  This
times exactly the cuFFT calls a library user would issue for K1..K5 (same shapes, same batch, CUDA events) and
nothing else, so the sum is a floor for any cuFFT-based pipeline.  Never imported by the product path."""
from __future__ import annotations

import ctypes
import os

import torch

from . import build as _build

STAGES = ("row_fwd", "col_band", "row_demod", "col_integrate", "row_inv")


def time_cufft_stages(shape, band_columns: int, frames: int = 32, reps: int = 5) -> dict:
    """{stage: us per frame} for the cuFFT transforms of each stage, plus the strided variant of K4."""
    if not os.path.exists(_build.CMP_LIB):
        raise ImportError(f"{_build.CMP_LIB} is missing: build it with `python -m fcd_b200.build`")
    lib = ctypes.CDLL(_build.CMP_LIB)
    lib.fcdcmp_time_stages.restype = ctypes.c_int
    lib.fcdcmp_time_stages.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
    lib.fcdcmp_last_error.restype = ctypes.c_char_p
    us = (ctypes.c_double * 6)()
    rc = lib.fcdcmp_time_stages(int(shape[0]), int(shape[1]), int(band_columns), int(frames), int(reps), us,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("cufft comparator: " + (lib.fcdcmp_last_error() or b"").decode())
    out = {"row_fwd": us[0], "col_band": us[1], "row_demod": us[2], "col_integrate": min(us[3], us[5]), "row_inv": us[4]}
    out["col_integrate_strided"] = us[3]
    out["col_integrate_contiguous"] = us[5]
    return out
