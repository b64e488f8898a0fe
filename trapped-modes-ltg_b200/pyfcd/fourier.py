"""Mirror of the reference's ``pyfcd/fourier.py`` (class ``fourier``), computed on the GPU.

Scalar helpers (wavenumber vectors) stay on the host exactly as the reference computes
them; everything that touches an image goes through the C ABI."""
import numpy as np
import torch

from fcd_b200 import engine as _eng


class fourier:

    @classmethod
    def find_peaks(cls, image):
        """(rightmost_peak, perpendicular_peak) in fftshift-ed pixel coordinates.
        Reference: pyfcd/fourier.py:8-41."""
        image = np.asarray(image) if not isinstance(image, torch.Tensor) else image
        plan = _eng.get_plan(tuple(image.shape))
        return plan.find_peaks(image)

    @classmethod
    def wavenumber(cls, size, calibration_factor=1, shifted=False):
        """Reference: pyfcd/fourier.py:44-57."""
        return _eng.wavenumber(size, calibration_factor, shifted)

    @classmethod
    def wavenumber_meshgrid(cls, shape, calibration_factor=1, shifted=False):
        """'ij' meshgrid (rows-mesh, cols-mesh).  Reference: pyfcd/fourier.py:59-73."""
        k_rows = cls.wavenumber(shape[0], calibration_factor, shifted)
        k_cols = cls.wavenumber(shape[1], calibration_factor, shifted)
        return np.meshgrid(k_rows, k_cols, indexing='ij')

    @classmethod
    def remove_degeneracy(cls, kx, ky, shape):
        """In place; zeroes index N//2+1 (not N//2), like the reference: pyfcd/fourier.py:76-92."""
        if shape[1] % 2 == 0:
            kx[:, shape[1] // 2 + 1] = 0
        if shape[0] % 2 == 0:
            ky[shape[0] // 2 + 1, :] = 0

    @classmethod
    def pixel_to_wavenumber(cls, image_shape, locations, calibration_factor=1):
        """Reference: pyfcd/fourier.py:95-113."""
        return _eng.pixel_to_wavenumber(image_shape, locations, calibration_factor)

    @classmethod
    def integrate_in_fourier(cls, gradient_x, gradient_y, calibration_factor=1):
        """Inverse-gradient integration of a user-supplied gradient pair in float64.
        Reference: pyfcd/fourier.py:116-137.  (The fused float32 version of this step is part
        of fcd.compute_height_map; this stage-level entry point runs the hand-written float64
        FFT kernels with torch CUDA elementwise glue.)"""
        gx = np.asarray(gradient_x, dtype=np.float64)
        gy = np.asarray(gradient_y, dtype=np.float64)
        plan = _eng.get_plan(gx.shape)
        dev = plan.device
        ky, kx = cls.wavenumber_meshgrid(gx.shape, calibration_factor)
        k2 = kx ** 2 + ky ** 2
        k2[0, 0] = 1
        cls.remove_degeneracy(kx, ky, gx.shape)
        kx_d, ky_d, k2_d = (torch.from_numpy(a).to(dev) for a in (kx, ky, k2))
        gxh = plan.fft2_c128(torch.from_numpy(gx).to(dev).to(torch.complex128))
        gyh = plan.fft2_c128(torch.from_numpy(gy).to(dev).to(torch.complex128))
        integrated_hat = (-1.0j * kx_d * gxh + -1.0j * ky_d * gyh) / k2_d
        return plan.fft2_c128(integrated_hat, inverse=True).real.contiguous().cpu().numpy()

    @classmethod
    def find_peak_locations(cls, image, threshold, no_peaks):
        """Reference: pyfcd/fourier.py:140-168."""
        image = np.asarray(image) if not isinstance(image, torch.Tensor) else image
        plan = _eng.get_plan(tuple(image.shape))
        return plan.peak_locations(image, threshold, no_peaks)
