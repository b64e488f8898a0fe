"""Build libfcd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "libfcd_b200.so")
SOURCES = ["fcd_b200.cu"]
HEADERS = ["fft_core.cuh", "fcd_kernels.cuh", "fcd_generic.cuh", "fcd_mask.cuh", "fcd_unwrap.cuh", "fcd_temporal.cuh", "fcd_launch.cuh", "fcd_plan.inl"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(REPO, "include", "fcd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(REPO, "include"), "-I", CSRC,
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


CMP_LIB = os.path.join(PKG_DIR, "libfcd_cufft_compare.so")


def build_cufft_compare(force: bool = False) -> str:
    """Benchmark-only comparator (csrc/cufft_compare.cu, links cuFFT); the product library does not."""
    src = os.path.join(CSRC, "cufft_compare.cu")
    if not force and os.path.exists(CMP_LIB) and os.path.getmtime(CMP_LIB) >= os.path.getmtime(src):
        return CMP_LIB
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                           "-shared", "-Xcompiler", "-fPIC", "-o", CMP_LIB, src, "-lcufft"])
    return CMP_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_cufft_compare(force=True))
