"""ctypes binding of the C ABI declared in include/fcd_b200.h.

There is no CPU fallback: if the CUDA library is missing or no CUDA device is present the
import of the compute path fails loudly."""
from __future__ import annotations

import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("FCD_B200_LIB", os.path.join(_PKG_DIR, "libfcd_b200.so"))

FCD_OK, FCD_ERR_INVALID, FCD_ERR_RUNTIME, FCD_ERR_STATE, FCD_ERR_NOPEAKS = 0, -1, -2, -3, -4

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); kept in one table so that tests can check that the shared
# library exports exactly what the header declares.
PROTOTYPES = {
    "fcd_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "fcd_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "fcd_last_error": (ctypes.c_char_p, []),
    "fcd_highpass_spectrum": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             c_double_p, ctypes.c_void_p]),
    "fcd_peak_locations": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int, c_int_p,
                                          c_int_p, ctypes.c_void_p]),
    "fcd_find_peaks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.c_void_p]),
    "fcd_bind_reference": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_double, ctypes.c_void_p]),
    "fcd_execute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]),
    "fcd_execute_typed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]),
    "fcd_set_height": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    "fcd_count_residues": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.c_void_p]),
    "fcd_unwrap_phase": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "fcd_temporal_mean_spectrum": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  c_double_p, c_int_p, ctypes.c_void_p]),
    "fcd_temporal_frames_supported": (ctypes.c_int, [ctypes.c_int]),
    "fcd_temporal_mean_spectrum_split": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_int_p,
                                                        ctypes.c_void_p]),
    "fcd_temporal_accumulate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, c_int_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_void_p]),
    "fcd_temporal_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "fcd_structure_mask": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "fcd_mask_center": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.c_void_p]),
    "fcd_get_carrier_mask": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "fcd_get_carrier_ccsgn": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "fcd_fft2_c128": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "fcd_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "fcd_stage_times": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.POINTER(ctypes.c_longlong),
                                       ctypes.POINTER(ctypes.c_longlong)]),
    "fcd_last_auto": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), c_int_p, c_int_p, ctypes.c_int]),
    "fcd_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "fcd_band_columns": (ctypes.c_int, [ctypes.c_void_p]),
    "fcd_plan_is_fused": (ctypes.c_int, [ctypes.c_void_p]),
}


def declare(lib: ctypes.CDLL) -> ctypes.CDLL:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libfcd_b200.so (built in-tree by fcd_b200.build); no device needed for this."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m fcd_b200.build` "
                              "(there is no CPU fallback for the FCD path)")
        _lib = declare(ctypes.CDLL(LIB_PATH))
    return _lib


class FcdError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fcd_b200 error {code}: {message}")
        self.code = code


def check(lib: ctypes.CDLL, rc: int) -> None:
    if rc == FCD_OK:
        return
    msg = (lib.fcd_last_error() or b"").decode("utf-8", "replace")
    if rc == FCD_ERR_NOPEAKS:
        # the reference raises ValueError from min() of an empty list (pyfcd/fourier.py:38)
        raise ValueError(msg or "min() arg is an empty sequence")
    raise FcdError(rc, msg)
