"""cuFFT-based version of the same FCD pipeline (torch.fft = cuFFT, torch elementwise kernels).

BENCHMARK / CROSS-CHECK UTILITY ONLY -- the comparison `north_star` asks for ("each is also
compared with a cuFFT-based version of the same pipeline").  The product path
(`HeightMapPlan.execute`, the `pyfcd` drop-in) never calls this.  It restates the reference
steps (pyfcd/fcd.py:28-33, 104-138; pyfcd/fourier.py:116-137) with library FFTs:

    F = fft2(frame); g_i = ifft2(F * mask_i); phi_i = -angle(g_i * ccsgn_i); scan unwrap;
    Z = fft2(phi_0 + i phi_1) -> Phi_0, Phi_1 by Hermitian symmetry; hhat = cA*Phi_0 + cB*Phi_1;
    h = real(ifft2(hhat))

using the same algebraic savings as the hand-written kernels where a library user would get
them for free (per-reference work hoisted, one packed complex transform for the two phase
fields, folded 2x2-solve / -1/height / -i k / k^2 coefficients).  float32 / complex64.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import HeightMapPlan, wavenumber

TWO_PI = 2.0 * np.pi


class CufftPipeline:
    def __init__(self, plan: HeightMapPlan):
        """Takes its per-reference state (masks, ccsgn, carriers, calibration) from a bound plan."""
        if plan.peaks is None:
            raise RuntimeError("bind the plan to a reference first")
        self.shape = plan.shape
        dev = plan.device
        H, W = self.shape
        self.masks = torch.stack([plan.carrier_mask(i) for i in range(2)]).to(torch.complex64)      # [2,H,W]
        self.ccsgn = torch.stack([plan.carrier_ccsgn(i, complex128=False) for i in range(2)])       # [2,H,W]
        f0, f1 = plan.carrier_frequencies()
        cal, height = plan.calibration_factor, plan.height
        det = f0[1] * f1[0] - f0[0] * f1[1]
        # integrate_in_fourier's meshes with its quirks (fourier.py:128-132)
        ky = np.repeat(wavenumber(H, cal)[:, None], W, axis=1)
        kx = np.repeat(wavenumber(W, cal)[None, :], H, axis=0)
        k2 = kx ** 2 + ky ** 2
        k2[0, 0] = 1
        kx[:, W // 2 + 1] = 0
        ky[H // 2 + 1, :] = 0
        # hhat = i/(height*det*k2) * [(kx f1r - ky f1c) Phi0 + (ky f0c - kx f0r) Phi1]; np.real(ifft2(.)) keeps
        # the Hermitian part: c_H(k) = (c(k) + conj(c(-k)))/2 applied to Phi (Phi(-k) = conj Phi(k))
        cA = 1j * (kx * f1[0] - ky * f1[1]) / (height * det * k2)
        cB = 1j * (ky * f0[1] - kx * f0[0]) / (height * det * k2)

        def herm(c):
            cm = np.conj(np.roll(np.flip(c, (0, 1)), (1, 1), (0, 1)))
            return 0.5 * (c + cm)

        self.cA = torch.from_numpy(herm(cA)).to(dev).to(torch.complex64)
        self.cB = torch.from_numpy(herm(cB)).to(dev).to(torch.complex64)

    @staticmethod
    def _unwrap_scan(w: torch.Tensor) -> torch.Tensor:
        """Same path as the fused kernels: rows from the centre column, rows linked along it."""
        n0, n1 = w.shape[-2:]
        jr = torch.zeros_like(w)
        jr[..., :, 1:] = torch.round((w[..., :, 1:] - w[..., :, :-1]) / TWO_PI)
        c = torch.cumsum(jr, dim=-1)
        c = c - c[..., :, n1 // 2:n1 // 2 + 1]
        col = w[..., :, n1 // 2]
        jc = torch.zeros_like(col)
        jc[..., 1:] = torch.round((col[..., 1:] - col[..., :-1]) / TWO_PI)
        m = torch.cumsum(jc, dim=-1)
        m = m - m[..., n0 // 2:n0 // 2 + 1]
        return w - TWO_PI * (c + m[..., :, None])

    def execute(self, frames: torch.Tensor, unwrap: bool = True) -> torch.Tensor:
        """frames [n,H,W] float32 CUDA -> height maps [n,H,W] float32."""
        F = torch.fft.fft2(frames)                                    # [n,H,W]
        g = torch.fft.ifft2(F[:, None] * self.masks[None])           # [n,2,H,W]
        ph = -torch.angle(g * self.ccsgn[None])
        if unwrap:
            ph = self._unwrap_scan(ph)
        Z = torch.fft.fft2(torch.complex(ph[:, 0], ph[:, 1]))         # packed phi0 + i phi1
        Zm = torch.conj(torch.roll(torch.flip(Z, (-2, -1)), (1, 1), (-2, -1)))
        P0 = 0.5 * (Z + Zm)
        P1 = -0.5j * (Z - Zm)
        return torch.fft.ifft2(self.cA * P0 + self.cB * P1).real.contiguous()
