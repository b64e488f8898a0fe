"""Mirror of ``analyze.mask``, ``analyze.center`` and ``analyze.load_image`` of the reference's
``pydata/analyze.py`` (class ``analyze``), the two per-frame steps that surround
``fcd.compute_height_map`` in the masked workflow (pydata/analyze.py:225-234, examples/mask_example.py).
Everything else in that 1300-line module (folder driver, statistics, plots) is out of scope."""
import numpy as np
import torch

from fcd_b200 import engine as _eng


class analyze:

    @classmethod
    def load_image(cls, path):
        """Grayscale image as float32.  Reference: pydata/analyze.py:26-40 (skimage.io.imread(as_gray=True)).
        Host-side file decode (OpenCV); single-channel files are returned unchanged in value."""
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3:
            raise ValueError("colour images are not supported by this mirror (the reference fixtures are grayscale)")
        return img.astype(np.float32)

    @classmethod
    def mask(cls, image, smoothed=14, show_mask=False, find_center=False):
        """Boolean mask of the floating structure; optionally (mask, center).
        Reference: pydata/analyze.py:43-100.  Bit-exact (box filter, mean threshold and
        connected components are reproduced exactly on the device)."""
        image = np.asarray(image, dtype=np.float32)
        plan = _eng.get_plan(image.shape, 1)
        m = plan.structure_mask(image, smoothed)
        mask = m.cpu().numpy()
        if show_mask:
            import matplotlib.pyplot as plt
            fig, ax = plt.subplots(1, 3, figsize=(10, 4))
            ax[0].imshow(image, cmap="gray"); ax[0].set_title("Original image")
            ax[1].imshow(mask, cmap="gray"); ax[1].set_title("Mask")
            ax[2].imshow(image * mask, cmap="gray"); ax[2].set_title("Masked image")
            for a in ax:
                a.axis("off")
            plt.tight_layout()
            plt.show()
        if find_center:
            return mask, cls.center(mask)
        return mask

    @classmethod
    def center(cls, mask):
        """(cy, cx) of the cavity inside the structure.  Reference: pydata/analyze.py:104-140."""
        mask = np.asarray(mask).astype(bool)
        plan = _eng.get_plan(mask.shape, 1)
        cy, cx = plan.mask_center(torch.from_numpy(mask))[0]
        if cy < 0:
            # the reference falls through to `return center` with no enclosed region found
            raise UnboundLocalError("cannot access local variable 'center' where it is not associated with a value")
        return (cy, cx)
