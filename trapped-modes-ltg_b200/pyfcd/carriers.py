"""Mirror of the reference's ``pyfcd/carriers.py`` (class ``Carrier``)."""
import numpy as np
import torch

from fcd_b200 import engine as _eng
from pyfcd.fourier import fourier


class Carrier:
    """Per-carrier state: ``pixels``, ``frequencies``, ``radius``, ``mask``, ``ccsgn``.
    Reference: pyfcd/carriers.py:9-24.  ``mask`` (bool, un-shifted layout) and ``ccsgn``
    (complex128) are produced on the GPU the first time they are read."""

    def __init__(self, reference_image, calibration_factor, peak, peak_radius):
        self.pixels = peak
        self.frequencies = fourier.pixel_to_wavenumber(np.shape(reference_image), peak, calibration_factor)
        self.radius = peak_radius
        self._reference = reference_image
        self._cal = calibration_factor
        self._mask = None
        self._ccsgn = None

    def _materialise(self):
        plan = _eng.HeightMapPlan(np.shape(self._reference), 1)
        try:
            # both slots carry this carrier; nothing is executed, so the singular pair is harmless
            plan.bind(self._reference, calibration_factor=self._cal, height=1.0,
                      peaks=(np.asarray(self.pixels), np.asarray(self.pixels)), radius=self.radius,
                      allow_collinear=True)
            self._mask = plan.carrier_mask(0).cpu().numpy()
            self._ccsgn = plan.carrier_ccsgn(0, complex128=True).cpu().numpy()
        finally:
            plan.close()

    @property
    def mask(self):
        if self._mask is None:
            self._materialise()
        return self._mask

    @property
    def ccsgn(self):
        if self._ccsgn is None:
            self._materialise()
        return self._ccsgn

    def peak_mask(self, shape, pos, r):
        """Reference: pyfcd/carriers.py:17-20."""
        other = Carrier(np.zeros(shape, dtype=np.float32), self._cal, pos, r)
        return other.mask

    def _ccsgn_of(self, reference_image):
        """Reference: pyfcd/carriers.py:22-24."""
        return Carrier(reference_image, self._cal, self.pixels, self.radius).ccsgn
