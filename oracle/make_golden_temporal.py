"""TEST INFRASTRUCTURE ONLY: golden vectors for analyze.block_amplitude / block_split from the
UNMODIFIED reference (pydata/analyze.py) run through oracle/ref_shims.py on a temporary folder
of ``*_map.npy`` files.

    python -m oracle.make_golden_temporal
"""
import os
import tempfile

import numpy as np

from oracle import temporal_oracle as to
from oracle.ref_shims import import_reference_analyze

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (frames, shape, num_blocks, mode, zero, f0 given?)
CASES = [(64, (64, 64), 4, 3, 0, False), (50, (64, 64), 4, 2, 0.01, False), (96, (64, 64), 16, 3, 0, True)]


def main():
    analyze = import_reference_analyze()
    g = {}
    for ci, (n, shape, nb, mode, zero, given) in enumerate(CASES):
        maps = to.synthetic_maps(n, shape, nb, seed=ci + 1)
        with tempfile.TemporaryDirectory() as d:
            for t in range(n):
                np.save(os.path.join(d, f"img_{t:05d}_map.npy"), maps[t])
            np.save(os.path.join(d, "calibration_factor.npy"), np.array([1.0]))
            for b in range(nb):
                f0 = 37.5 if given else None
                ref = analyze.block_amplitude(d, f0=f0, tasa=500, mode=mode, num_blocks=nb, block_index=b, zero=zero)
                mine = to.block_amplitude(maps, f0=f0, tasa=500, mode=mode, num_blocks=nb, block_index=b, zero=zero)
                assert np.array_equal(np.array(ref[0]), np.array(mine[0])) and ref[3] == mine[3]
                assert np.array_equal(ref[1], mine[1], equal_nan=True) and np.array_equal(ref[2], mine[2], equal_nan=True)
                g[f"case{ci}.block{b}.harmonics"] = np.array(ref[0], dtype=np.float64)
                g[f"case{ci}.block{b}.amps"] = ref[1].astype(np.float32)
                g[f"case{ci}.block{b}.phases"] = ref[2].astype(np.float32)
                g[f"case{ci}.block{b}.f0"] = np.float64(ref[3])
            sp = analyze.block_split(d, num_blocks=nb, block_index=nb - 1)
            assert np.array_equal(sp, to.block_split(maps, None, nb, nb - 1), equal_nan=True)
        g[f"case{ci}.params"] = np.array([n, shape[0], shape[1], nb, mode, zero, 37.5 if given else np.nan])
        print("case", ci, "frames", n, "blocks", nb, "f0", [float(g[f'case{ci}.block{b}.f0']) for b in range(nb)][:4])
    np.savez_compressed(os.path.join(OUT, "golden_temporal.npz"), **g)
    print("wrote golden_temporal.npz", os.path.getsize(os.path.join(OUT, "golden_temporal.npz")), "bytes")


if __name__ == "__main__":
    main()
