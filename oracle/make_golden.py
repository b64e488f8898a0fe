"""TEST INFRASTRUCTURE ONLY: generate tests/golden/*.npz by running the UNMODIFIED reference
(``/root/reference/pyfcd`` and ``pyval/val.py``) in the build container through
``oracle/ref_shims.py``.  Run from the repo root:

    python -m oracle.make_golden

The vectors pin (a) the oracle restatement (tests/test_oracle_golden.py, CPU) and (b) the
CUDA path (tests/test_gpu_*.py) to the reference's own arithmetic.  The only code that is
not the reference's on this path are the scikit-image stand-ins documented in
``ref_shims.py`` (label/regionprops, disk, unwrap_phase) and an inert matplotlib.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from oracle import fcd_oracle as o
from oracle.ref_shims import REFERENCE_ROOT, import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_case(fcd, n, a, b, center, sigma, peak, seed_noise=None):
    ref = o.rotated_board(n, a=a, b=b)
    h, uy, ux = o.gaussian_bump_displacement(n, center, sigma, peak)
    frame = o.rotated_board(n, a=a, b=b, uy=uy, ux=ux)
    sq = o.board_square_size(n, a)
    hm, ph, cal = fcd.compute_height_map(ref.astype(np.float64), frame.astype(np.float64), sq, height=1.0)
    carriers, _ = fcd.compute_carriers(ref.astype(np.float64), sq)
    raw = fcd.compute_phases(np.fft.fft2(frame.astype(np.float64)), carriers, unwrap=False)
    return dict(ref=ref, frame=frame, square_size=np.float64(sq), height_map=hm,
                phase_minmax=np.array([ph.min(), ph.max()]), wrapped_minmax=np.array([raw.min(), raw.max()]),
                phases_sub=ph[:, ::4, ::4].copy(), cal=np.float64(cal),
                pixels=np.array([c.pixels for c in carriers]),
                freqs=np.array([c.frequencies for c in carriers]),
                radius=np.float64(carriers[0].radius),
                residues=np.array([o.count_residues(p) for p in raw]),
                bump=np.array([center[0], center[1], sigma, peak]))


def noisy_reference(n, period_px, angle_deg, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:n, 0:n].astype(np.float64)
    th = np.deg2rad(angle_deg)
    k = 2 * np.pi / period_px
    A = k * (np.cos(th) * x + np.sin(th) * y)
    B = k * (-np.sin(th) * x + np.cos(th) * y)
    img = 0.5 + 0.45 * np.sign(np.sin(A) * np.sin(B)) * 0.5 + 0.25 * np.sin(A) * np.sin(B)
    img *= 1.0 - 0.3 * ((x - n / 2) ** 2 + (y - n / 2) ** 2) / (n * n)  # vignetting
    img += 0.02 * rng.standard_normal((n, n))
    return np.clip(np.round(img * 200.0), 0, 255).astype(np.float32)


def main():
    fcd, fourier, Carrier = import_reference()
    os.makedirs(OUT, exist_ok=True)
    g = {}

    # 1/2: rotated synthetic board, 256^2, no-wrap and wrapping frames
    for name, peak, center, sigma in (("small", 0.6, (120.0, 140.0), 30.0), ("wrap", 14.0, (131.0, 122.0), 45.0)):
        c = synth_case(fcd, 256, 15.0, 1.0, center, sigma, peak)
        for k, v in c.items():
            g[f"synth256_{name}.{k}"] = v
        print(name, "pixels", c["pixels"].tolist(), "R", float(c["radius"]), "cal", float(c["cal"]),
              "residues", c["residues"].tolist(), "phase range", c["phase_minmax"].tolist(), "wrapped range", c["wrapped_minmax"].tolist())

    # 3: the reference's own validator (pyval/val.py) at N=256, n=15, val_example-like surface
    sys.path.insert(0, REFERENCE_ROOT)
    saved = {k: v for k, v in sys.modules.items() if k == "pyfcd" or k.startswith("pyfcd.")}
    try:
        import importlib
        for k in list(saved):
            del sys.modules[k]
        val = importlib.import_module("pyval.val").val
    finally:
        sys.path.remove(REFERENCE_ROOT)

    def step(X, a=0.08, w=100):
        x0 = len(X) // 2
        return 1 / (1 + np.exp(-a * (X - x0 + w / 2))) * 1 / (1 + np.exp(a * (X - x0 - w / 2)))

    def gauss_sin(X, Y, A=100, w=0.05):
        return step(X) * step(Y) * A * np.sin(w * (X + Y))

    X, Y, h, I, hm, I0, cal = val(0, func=gauss_sin, N=256, n=15)
    for k in [k for k in sys.modules if k == "pyfcd" or k.startswith("pyfcd.") or k.startswith("pyval")]:
        del sys.modules[k]
    sys.modules.update(saved)
    carriers, _ = fcd.compute_carriers(I0, 256 / 30)
    raw = fcd.compute_phases(np.fft.fft2(I), carriers, unwrap=False)
    g["val256.I0"], g["val256.I"], g["val256.h"], g["val256.height_map"] = I0, I, h.astype(np.float32), hm
    g["val256.cal"] = np.float64(cal)
    g["val256.pixels"] = np.array([c.pixels for c in carriers])
    g["val256.radius"] = np.float64(carriers[0].radius)
    g["val256.residues"] = np.array([o.count_residues(p) for p in raw])
    print("val256 pixels", g["val256.pixels"].tolist(), "cal", cal, "residues", g["val256.residues"].tolist(),
          "max err %", np.max(np.abs(hm - h)) * 100 / np.max(np.abs(hm)), "wrapped range", raw.min(), raw.max())

    # 4: carrier search on camera-like references (binary squares, vignetting, noise, 8-bit)
    for i, (n, period, ang, seed) in enumerate(((256, 17.3, 8.0, 1), (512, 19.1, -12.5, 2), (256, 14.2, 40.0, 3))):
        img = noisy_reference(n, period, ang, seed)
        cal, peaks = fcd.compute_calibration_factor(0.0022, img.astype(np.float64))
        g[f"noisy{i}.image"] = img.astype(np.uint8)
        g[f"noisy{i}.peaks"] = np.array(peaks)
        g[f"noisy{i}.cal"] = np.float64(cal)
        print("noisy", i, np.array(peaks).tolist(), cal)

    # 5: layers -> effective height (examples/fcd_example.py:17)
    layers = [[5.7e-2, 1.0003], [1.2e-2, 1.48899], [4.3e-2, 1.34], [80e-2, 1.0003]]
    g["layers.example"] = np.array(layers)
    g["layers.height"] = np.float64(fcd.height_from_layers(layers))

    # 6: Fourier integration with the N//2+1 quirk on broadband fields, square and not
    rng = np.random.default_rng(7)
    for shape in ((64, 64), (64, 128)):
        gx, gy = rng.standard_normal(shape), rng.standard_normal(shape)
        key = f"integrate{shape[0]}x{shape[1]}"
        g[key + ".gx"], g[key + ".gy"] = gx, gy
        g[key + ".h"] = fourier.integrate_in_fourier(gx, gy, 0.37)

    np.savez_compressed(os.path.join(OUT, "golden_fcd.npz"), **g)
    print("wrote", os.path.join(OUT, "golden_fcd.npz"), os.path.getsize(os.path.join(OUT, "golden_fcd.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
