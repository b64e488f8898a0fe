"""TEST INFRASTRUCTURE ONLY: golden vector from the reference's OWN example data
(examples/fcd_example.py: Pictures/reference_2.png + Pictures/202406_1457001661.bmp, camera frames
whose wrapped phases contain residues), computed by the UNMODIFIED reference through
oracle/ref_shims.py on the central 512 x 512 crop (kept small for the repository).

    python -m oracle.make_golden_real
"""
import os

import cv2
import numpy as np

from oracle import fcd_oracle as o
from oracle.ref_shims import REFERENCE_ROOT, import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
LAYERS = [[5.7e-2, 1.0003], [1.2e-2, 1.48899], [4.3e-2, 1.34], [80e-2, 1.0003]]     # examples/fcd_example.py:17
SQUARE = 0.0022                                                                      # examples/fcd_example.py:19


def main():
    fcd = import_reference()[0]
    pics = os.path.join(REFERENCE_ROOT, "examples", "Pictures")
    ref = cv2.imread(os.path.join(pics, "reference_2.png"), cv2.IMREAD_UNCHANGED)[256:768, 256:768]
    frm = cv2.imread(os.path.join(pics, "202406_1457001661.bmp"), cv2.IMREAD_UNCHANGED)[256:768, 256:768]
    assert ref.dtype == np.uint8 and frm.dtype == np.uint8 and ref.shape == (512, 512)
    r32, f32 = ref.astype(np.float32), frm.astype(np.float32)                         # analyze.load_image
    hm, ph, cal = fcd.compute_height_map(r32, f32, SQUARE, LAYERS)
    hmo, pho, calo = o.compute_height_map(r32, f32, SQUARE, LAYERS)
    # the reference transforms float32 frames in complex64 (scipy keeps single precision); the oracle upcasts first
    assert np.linalg.norm(hm - hmo) / np.linalg.norm(hm) < 1e-5 and cal == calo
    _, wrapped, _ = o.compute_height_map(r32, f32, SQUARE, LAYERS, unwrap=False)
    res = [o.count_residues(wrapped[i]) for i in range(2)]
    print("crop residues", res, "cal", cal, "height range", float(hm.min()), float(hm.max()))
    np.savez_compressed(os.path.join(OUT, "golden_real.npz"), ref=ref, frame=frm, height_map=hm.astype(np.float32),
                        cal=np.float64(cal), residues=np.array(res), square_size=np.float64(SQUARE),
                        layers=np.array(LAYERS))
    print("wrote golden_real.npz", os.path.getsize(os.path.join(OUT, "golden_real.npz")), "bytes")


if __name__ == "__main__":
    main()
