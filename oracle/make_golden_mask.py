"""TEST INFRASTRUCTURE ONLY: golden vectors for analyze.mask / analyze.center from the
UNMODIFIED reference (pydata/analyze.py) run through oracle/ref_shims.py.

    python -m oracle.make_golden_mask
"""
import os

import numpy as np

from oracle import mask_oracle as mo
from oracle.ref_shims import import_reference_analyze

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    analyze = import_reference_analyze()
    g = {}
    cases = [((256, 256), 1, 15), ((256, 256), 2, 14), ((128, 256), 3, 15), ((512, 512), 4, 14)]
    for i, (shape, seed, smoothed) in enumerate(cases):
        img = mo.synthetic_structure(shape, seed)
        m = analyze.mask(img, smoothed=smoothed)
        c = analyze.center(m)
        assert np.array_equal(m, mo.mask(img, smoothed)) and tuple(c) == mo.center(m)
        g[f"case{i}.seed"] = np.int64(seed)
        g[f"case{i}.shape"] = np.array(shape)
        g[f"case{i}.smoothed"] = np.int64(smoothed)
        g[f"case{i}.mask"] = np.packbits(m)
        g[f"case{i}.center"] = np.array(c)
        print(i, shape, smoothed, "mask area", int(m.sum()), "center", c)
    np.savez_compressed(os.path.join(OUT, "golden_mask.npz"), **g)
    print("wrote golden_mask.npz", os.path.getsize(os.path.join(OUT, "golden_mask.npz")), "bytes")


if __name__ == "__main__":
    main()
