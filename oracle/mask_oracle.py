"""CPU oracle for the floating-structure mask and its cavity centre -- TEST INFRASTRUCTURE ONLY.

Restates ``analyze.mask`` (pydata/analyze.py:43-100) and ``analyze.center``
(pydata/analyze.py:104-140).  Pinned: tests/golden/golden_mask.npz is produced by the
unmodified reference class through oracle/ref_shims.py (scikit-image's label/regionprops
replaced by the 8-connected raster-order stand-ins, as for the FCD oracle).  Third-party
arithmetic on the path that IS installed and therefore called directly:
``scipy.ndimage.uniform_filter`` (scipy 1.18.1) and ``numpy.mean`` on float32 (numpy 2.3);
the CUDA kernels replicate both bit-for-bit (running double sum ``tmp += new - old`` per line,
axis 0 then axis 1, reflect boundary; pairwise float32 summation in 128-element blocks).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter

from oracle.fcd_oracle import label8


def mask(image: np.ndarray, smoothed: int = 14) -> np.ndarray:
    """analyze.py:66-75: box filter, threshold at the mean, largest 8-connected region
    (ties: the first in label order)."""
    smooth = uniform_filter(image, size=smoothed)
    threshold = np.mean(smooth)
    below = smooth < threshold
    lab, n = label8(below)
    if n == 0:
        raise IndexError("list index out of range")      # regions_sorted[0] on an empty list
    areas = np.bincount(lab.ravel(), minlength=n + 1)[1:]
    best = int(np.argmax(areas)) + 1                      # first maximum == stable sort, reverse=True
    return lab == best


def center(mask_: np.ndarray):
    """analyze.py:121-140: largest 8-connected region of ~mask whose bounding box does not
    touch the image border; (int(row centroid), int(col centroid))."""
    lab, n = label8(~mask_)
    n_rows, n_cols = mask_.shape
    best_area, best = -1, None
    for l in range(1, n + 1):
        coords = np.argwhere(lab == l)
        r0, c0 = coords.min(axis=0)
        r1, c1 = coords.max(axis=0) + 1
        if r0 > 0 and c0 > 0 and r1 < n_rows and c1 < n_cols and len(coords) > best_area:
            best_area, best = len(coords), coords
    if best is None:
        raise UnboundLocalError("cannot access local variable 'center' where it is not associated with a value")
    cy, cx = best.mean(axis=0)
    return int(cy), int(cx)


def synthetic_structure(shape, seed, ring=(0.22, 0.33), offset=(0.05, -0.08), noise=0.06, blobs=6):
    """Camera-like frame: bright checker texture, a dark floating ring (annulus) with a bright
    cavity, a few dark specks and bright pinholes; float32 in 16-bit counts."""
    rng = np.random.default_rng(seed)
    h, w = shape
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    cy, cx = h * (0.5 + offset[0]), w * (0.5 + offset[1])
    r = np.hypot((y - cy) / min(h, w), (x - cx) / min(h, w))
    img = 30000.0 + 9000.0 * np.sin(2 * np.pi * x / 11.0) * np.sin(2 * np.pi * y / 11.0)
    img *= 1.0 - 0.25 * ((x - w / 2) ** 2 + (y - h / 2) ** 2) / (h * w)
    inside = (r > ring[0]) & (r < ring[1])
    img[inside] *= 0.18
    for _ in range(blobs):                                   # dark specks outside, bright pinholes in the ring
        by, bx = rng.uniform(0.05, 0.95) * h, rng.uniform(0.05, 0.95) * w
        rad = rng.uniform(3, 9)
        d = np.hypot(y - by, x - bx) < rad
        img[d] = np.where(inside[d], 32000.0, 4000.0)
    img += noise * 30000.0 * rng.standard_normal(shape)
    return np.clip(img, 0, 65535).astype(np.float32)
