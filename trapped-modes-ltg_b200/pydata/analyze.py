"""Mirror of the parts of the reference's ``pydata/analyze.py`` (class ``analyze``) that sit on
either side of the FCD path: ``load_image``, ``mask``, ``center`` (the per-frame
steps around ``fcd.compute_height_map`` in the masked workflow, analyze.py:225-234),
``folder`` (the batch driver, analyze.py:141-286, non-polar output) and the temporal analysis of
the resulting maps, ``block_split`` / ``block_amplitude`` (analyze.py:365-417, 542-641).
Statistics, fits and plots of that 1300-line module are out of scope."""
import os

import numpy as np
import torch

from fcd_b200 import engine as _eng
from fcd_b200 import temporal as _temporal

_stack_cache: dict = {}


def _map_files(map_folder, t_limit=None):
    files = sorted(f for f in os.listdir(map_folder) if f.endswith('_map.npy') and 'calibration_factor' not in f)
    return files[:t_limit]


def _load_stack(map_folder, t_limit=None) -> torch.Tensor:
    """All ``*_map.npy`` of a folder as one CUDA float32 stack, read ONCE and kept while the files
    are unchanged (the reference re-reads every file for every block, analyze.py:582-592)."""
    files = _map_files(map_folder, t_limit)
    key = (os.path.abspath(map_folder), tuple((f, os.path.getmtime(os.path.join(map_folder, f))) for f in files))
    hit = _stack_cache.get("stack")
    if hit is not None and hit[0] == key:
        return hit[1]
    first = np.load(os.path.join(map_folder, files[0]))
    host = torch.empty((len(files),) + first.shape, dtype=torch.float32).pin_memory()
    for t, f in enumerate(files):
        host[t] = torch.from_numpy(np.load(os.path.join(map_folder, f)).astype(np.float32, copy=False))
    stack = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    _stack_cache.clear()
    _stack_cache["stack"] = (key, stack)
    return stack


class analyze:

    @classmethod
    def load_image(cls, path):
        """Grayscale image as float32.  Reference: pydata/analyze.py:26-40 (skimage.io.imread(as_gray=True)).
        Host-side file decode (OpenCV).  Single-channel files are returned unchanged in value (what as_gray does to
        a 2-D image: nothing).  Colour files go through scikit-image's documented conversion, restated here because
        scikit-image is not installed (unpinned, like the other scikit-image boundaries): integers scaled to [0, 1]
        (img_as_float), RGBA blended over white (rgba2rgb), luminance 0.2125 R + 0.7154 G + 0.0721 B (rgb2gray)."""
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3:
            if img.dtype == np.uint8:
                f = img.astype(np.float64) / 255.0
            elif img.dtype == np.uint16:
                f = img.astype(np.float64) / 65535.0
            else:
                f = img.astype(np.float64)
            if f.shape[2] == 4:                                   # OpenCV order B, G, R, A
                a = f[..., 3:4]
                f = np.clip((1.0 - a) + a * f[..., :3], 0.0, 1.0)
            elif f.shape[2] != 3:
                raise ValueError(f"unsupported channel count {f.shape[2]}")
            return (0.2125 * f[..., 2] + 0.7154 * f[..., 1] + 0.0721 * f[..., 0]).astype(np.float32)
        return img.astype(np.float32)

    @classmethod
    def folder(cls, reference_path, displaced_dir, layers, square_size, smoothed=None, polar=False,
               show_mask=False, batch=16, **kwargs):
        """Height maps of every ``.tif`` of a folder -> ``<displaced_dir>/maps/<name>_map.npy`` (float32),
        ``calibration_factor.npy`` and, with ``smoothed``, ``centers.txt``.  Reference:
        pydata/analyze.py:141-286, same file selection, resume rule and output formats.  Frames are
        decoded on the host, then masked (analyze.mask / center), demodulated and integrated on
        the device ``batch`` at a time with the per-reference work done once; integer camera
        frames travel as integers and are widened in the first kernel.  ``show_mask`` runs the reference's
        interactive preview first (``_preview_mask``); the polar output (cv2.fitEllipse + warp) is not part of
        this mirror."""
        if polar:
            raise NotImplementedError("polar maps (analyze.py:236-243,264-275) are outside the mirrored path")
        if show_mask and not smoothed:
            raise ValueError("If show_mask == True, expect smoothed too")          # analyze.py:216-217
        import cv2
        reference = cls.load_image(reference_path)
        file_list = sorted(os.listdir(displaced_dir))
        tif_list = [f for f in file_list if f.endswith('.tif') and 'reference' not in f]
        output_dir = os.path.join(displaced_dir, 'maps')
        os.makedirs(output_dir, exist_ok=True)
        existing_maps = sorted(f for f in os.listdir(output_dir) if f.endswith('_map.npy'))
        start_index = len(existing_maps)
        centers_path = os.path.join(output_dir, 'centers.txt')
        if smoothed:
            if not os.path.exists(centers_path):
                open(centers_path, "w").close()
            with open(centers_path, "r") as f:
                lines = f.readlines()
            start_index = max(start_index, len(lines))
            if show_mask:
                smoothed = cls._preview_mask(displaced_dir, tif_list, smoothed)
        todo = list(enumerate(tif_list))[start_index:]
        if not todo:
            return
        plan = _eng.get_plan(reference.shape, max(1, int(batch)))
        calibration_factor = plan.bind(reference, square_size=square_size, layers=layers)
        calibration_saved = False
        for b0 in range(0, len(todo), max(1, int(batch))):
            chunk = todo[b0:b0 + max(1, int(batch))]
            raw = [cv2.imread(os.path.join(displaced_dir, fname), cv2.IMREAD_UNCHANGED) for _, fname in chunk]
            for (_, fname), img in zip(chunk, raw):
                if img is None or img.ndim != 2 or img.shape != reference.shape:
                    raise ValueError(f"{fname}: not a grayscale image of the reference's shape")
            same = len({im.dtype for im in raw}) == 1 and raw[0].dtype in (np.uint8, np.uint16)
            host = np.stack(raw) if same else np.stack([im.astype(np.float32) for im in raw])
            frames = torch.from_numpy(host).to(plan.device)
            masks = centers = None
            if smoothed:
                masks = plan.structure_mask(frames.to(torch.float32), smoothed)          # analyze.py:227
                centers = plan.mask_center(masks)                                         # analyze.py:230
                if any(c[0] < 0 for c in centers):
                    raise UnboundLocalError("cannot access local variable 'center' where it is not associated with a value")
            # np.where(mask, reference, frame) before and `height_map *= ~mask` after are fused into the kernels
            height = plan.execute(frames, mask=masks, unwrap="auto").cpu().numpy()
            for k, (i, fname) in enumerate(chunk):
                base_name = fname.replace('.tif', '')
                np.save(os.path.join(output_dir, f"{base_name}_map.npy"), height[k])
                if smoothed:
                    with open(centers_path, "a") as f:
                        f.write(f"{i}\t{(int(centers[k][0]), int(centers[k][1]))}\n")
                if not calibration_saved:
                    np.save(os.path.join(output_dir, 'calibration_factor.npy'), np.array([calibration_factor]))
                    calibration_saved = True

    @classmethod
    def _preview_mask(cls, displaced_dir, tif_list, smoothed, ask=input):
        """The interactive preview of ``folder(show_mask=True)``: show the mask of the eleventh frame (or the last
        one), ask, and let the user try another ``smoothed`` until the answer is "Y".  Returns the accepted value.
        Reference: pydata/analyze.py:193-215 (same prompts, same 8 s pause)."""
        import matplotlib.pyplot as plt
        displaced_image = cls.load_image(os.path.join(displaced_dir, tif_list[min(10, len(tif_list) - 1)]))
        while True:
            cls.mask(displaced_image, smoothed=smoothed, show_mask=True)
            plt.pause(8)
            plt.close("all")
            message = ask("Continue with this mask? [Y,n]: ")
            if message == "Y":
                return smoothed
            if message == "n":
                try:
                    smoothed = int(ask("Enter new smoothed value (int): "))
                except ValueError:
                    print("Invalid input, keeping previous smoothed.")

    @classmethod
    def block_split(cls, map_folder, t_limit=None, num_blocks=64, block_index=0):
        """Temporal stack (block_size, block_size, N) of one spatial block, NaN where the first map
        is zero.  Reference: pydata/analyze.py:365-417."""
        stack = _load_stack(map_folder, t_limit)
        H = stack.shape[1]
        bpr = int(np.sqrt(num_blocks))
        bs = H // bpr
        i, j = block_index // bpr, block_index % bpr
        blk = stack[:, i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
        valid = (stack[0] != 0)[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
        out = torch.where(valid[None], blk, torch.full((), float("nan"), device=blk.device))
        return out.permute(1, 2, 0).cpu().numpy()

    @classmethod
    def block_amplitude(cls, map_folder, f0=None, tasa=500, mode=1, num_blocks=64, block_index=0, zero=0):
        """(harmonics, amps, phases, f0) of one spatial block.  Reference: pydata/analyze.py:542-641.
        All blocks are analysed on the device in one pass over the stack the first time a folder /
        parameter set is seen (csrc/fcd_temporal.cuh); later block indices are slices of that result."""
        stack = _load_stack(map_folder)
        key = (id(stack), f0, tasa, mode, num_blocks, zero)
        hit = _stack_cache.get("amps")
        if hit is None or hit[0] != key:
            res = _temporal.block_amplitudes(stack, f0=f0, tasa=tasa, mode=mode, num_blocks=num_blocks, zero=zero)
            _stack_cache["amps"] = (key, res)
        return _stack_cache["amps"][1].block(block_index)

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
