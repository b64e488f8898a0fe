"""TEST INFRASTRUCTURE ONLY: import the *actual* reference (``/root/reference/pyfcd``) in the
build container.

The reference needs scikit-image and matplotlib at import time
(pyfcd/fcd.py:3-4, pyfcd/fourier.py:3, pyfcd/carriers.py:6); neither is installed and
there is no network.  This module registers minimal stand-in modules for exactly those
imports -- ``skimage.restoration.unwrap_phase``, ``skimage.measure.label`` /
``regionprops``, ``skimage.draw.disk`` and an inert ``matplotlib.pyplot`` -- and then
imports the reference unchanged from its read-only location.  Every other line that runs
is the reference's own code, so vectors produced through this path pin the oracle
restatement to the reference (see ``oracle/make_golden.py``).

``/root/reference`` does not exist on the GPU box; nothing that runs there imports this.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyfcd"))


class _Region:
    """The regionprops attributes the reference reads (label, coords, area, bbox, centroid)."""

    def __init__(self, lab, l):
        self.label = l
        self.coords = np.argwhere(lab == l)
        self.area = len(self.coords)
        r, c = self.coords[:, 0], self.coords[:, 1]
        self.bbox = (int(r.min()), int(c.min()), int(r.max()) + 1, int(c.max()) + 1)
        self.centroid = tuple(self.coords.mean(axis=0))


def _install_shims():
    from oracle import fcd_oracle as o

    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        restoration = types.ModuleType("skimage.restoration")
        restoration.unwrap_phase = o.unwrap_phase
        measure = types.ModuleType("skimage.measure")

        def label(img, *a, **k):
            return o.label8(np.asarray(img) != 0)[0]

        def regionprops(lab, *a, **k):
            return [_Region(lab, l) for l in range(1, int(lab.max()) + 1)]

        measure.label = label
        measure.regionprops = regionprops
        measure.find_contours = lambda *a, **k: []
        sk_io = types.ModuleType("skimage.io")
        sk_transform = types.ModuleType("skimage.transform")
        sk_transform.warp = sk_transform.rotate = lambda *a, **k: None
        sk.io, sk.transform = sk_io, sk_transform
        sys.modules.update({"skimage.io": sk_io, "skimage.transform": sk_transform})
        draw = types.ModuleType("skimage.draw")

        def disk(center, radius, *, shape=None):
            m = o.disk_mask(shape, center, radius)
            return np.nonzero(m)

        draw.disk = disk
        sk.restoration, sk.measure, sk.draw = restoration, measure, draw
        sys.modules.update({"skimage": sk, "skimage.restoration": restoration,
                            "skimage.measure": measure, "skimage.draw": draw})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            raise RuntimeError("matplotlib is not installed (oracle shim)")

        plt.subplots = _noop
        plt.show = _noop
        mpl.pyplot = plt
        anim = types.ModuleType("matplotlib.animation")
        mpl.animation = anim
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.animation": anim})


class _reference_path:
    """sys.path with the reference root first and WITHOUT any directory that holds this repo's
    drop-in packages: the reference's pyfcd/ and pydata/ have no __init__.py (namespace packages),
    so a regular package of the same name anywhere on sys.path would win over them."""

    def __enter__(self):
        self.saved = list(sys.path)
        keep = [p for p in sys.path if not os.path.exists(os.path.join(p or ".", "pyfcd", "__init__.py"))]
        sys.path[:] = [REFERENCE_ROOT] + keep
        importlib.invalidate_caches()

    def __exit__(self, *exc):
        sys.path[:] = self.saved
        importlib.invalidate_caches()


def import_reference():
    """Returns (fcd, fourier, Carrier) classes of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _install_shims()
    # make sure 'pyfcd' resolves to the reference, not to the product's drop-in package
    saved = {k: v for k, v in sys.modules.items() if k == "pyfcd" or k.startswith("pyfcd.")}
    for k in saved:
        del sys.modules[k]
    try:
        with _reference_path():
            m_fcd = importlib.import_module("pyfcd.fcd")
            m_four = importlib.import_module("pyfcd.fourier")
            m_car = importlib.import_module("pyfcd.carriers")
        assert m_fcd.__file__.startswith(REFERENCE_ROOT), m_fcd.__file__
        ref = (m_fcd.fcd, m_four.fourier, m_car.Carrier)
    finally:
        for k in [k for k in sys.modules if k == "pyfcd" or k.startswith("pyfcd.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref


def import_reference_analyze():
    """The unmodified reference class ``pydata.analyze.analyze`` (for analyze.mask / analyze.center)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _install_shims()
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("pyfcd", "pydata")}
    for k in saved:
        del sys.modules[k]
    try:
        with _reference_path():
            mod = importlib.import_module("pydata.analyze")
        assert mod.__file__.startswith(REFERENCE_ROOT), mod.__file__
        cls = mod.analyze
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("pyfcd", "pydata")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return cls
