"""Synthetic FCD workload: an exactly periodic rotated checkerboard, Gaussian-bump
displacement fields and the deformed frames I(r) = I0(r - u) evaluated analytically (the template is the
reference's own validator, pyval/val.py:79-108).  Input generation only -- used by bench.py, the scripts,
the tests and (re-exported) the oracle; no part of the height-map computation lives here."""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


def rotated_board(n: int, a: float | None = None, b: float | None = None, eps: float = 0.1,
                  uy=None, ux=None, dtype=np.float32) -> np.ndarray:
    """Exactly periodic rotated checkerboard I0(r - u).  With uy=ux=None returns I0.
    A = 2pi(a*y + b*x)/n, B = 2pi(-b*y + a*x)/n,
    I0 = 0.5 + 0.25*((1+eps)cos(A-B) - cos(A+B))/(1+eps/2)."""
    a = 60.0 * n / 1024.0 if a is None else a
    b = 3.0 * n / 1024.0 if b is None else b
    y = np.arange(n, dtype=np.float64)[:, None]
    x = np.arange(n, dtype=np.float64)[None, :]
    if uy is not None:
        y = y - uy
        x = x - ux
    A = TWO_PI * (a * y + b * x) / n
    B = TWO_PI * (-b * y + a * x) / n
    img = 0.5 + 0.25 * ((1.0 + eps) * np.cos(A - B) - np.cos(A + B)) / (1.0 + eps / 2.0)
    return img.astype(dtype)


def board_square_size(n: int, a: float | None = None) -> float:
    a = 60.0 * n / 1024.0 if a is None else a
    return n / (2.0 * a)


def gaussian_bump_displacement(n: int, center, sigma: float, peak_disp: float, H: float = 1.0):
    """h = A*exp(-r^2/2sigma^2); u = -H*grad(h), scaled so max|u| = peak_disp pixels.
    Returns (h, u_row, u_col)."""
    y = np.arange(n, dtype=np.float64)[:, None] - center[0]
    x = np.arange(n, dtype=np.float64)[None, :] - center[1]
    g = np.exp(-(y * y + x * x) / (2.0 * sigma * sigma))
    # |grad g| peaks at r = sigma with value exp(-1/2)/sigma
    amp = peak_disp * sigma * np.exp(0.5) / H
    h = amp * g
    hy = -amp * y / (sigma * sigma) * g
    hx = -amp * x / (sigma * sigma) * g
    return h, -H * hy, -H * hx


def synthetic_frames(n: int, count: int, seed: int = 20251018, peak_range=(0.2, 0.8), dtype=np.float32):
    """Reference + ``count`` deformed frames + ground-truth heights."""
    rng = np.random.default_rng(seed)
    ref = rotated_board(n, dtype=dtype)
    frames = np.empty((count, n, n), dtype=dtype)
    truth = np.empty((count, n, n), dtype=np.float64)
    for i in range(count):
        cy, cx = rng.uniform(0.35 * n, 0.65 * n, size=2)
        sigma = rng.uniform(n / 12.0, n / 6.0)
        peak = rng.uniform(*peak_range)
        h, uy, ux = gaussian_bump_displacement(n, (cy, cx), sigma, peak)
        frames[i] = rotated_board(n, uy=uy, ux=ux, dtype=dtype)
        truth[i] = h
    return ref, frames, truth
