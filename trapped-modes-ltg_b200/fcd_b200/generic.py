"""Per-frame path for frame shapes that are not powers of two (GENERIC plans, include/fcd_b200.h).

The reference takes any image shape (pyfcd/fcd.py:14; scipy's fft2 does).  The fused float32 kernels K1..K5 exist for
powers of two -- every fixture and every BASELINE.json config -- so for other shapes the path is composed, in float64,
from the stage-level pieces of the C ABI: `fcd_fft2_c128` (Bluestein's chirp convolution on the hand-written
power-of-two float64 transforms), the carrier masks / ccsgn of `fcd_bind_reference`, `fcd_count_residues` and
`fcd_unwrap_phase`; the elementwise steps between them (mask product, angle, 2x2 solve, k-space coefficients) are torch
CUDA ops.  Same steps and quirks as pyfcd/fcd.py:28-33, 104-138 and pyfcd/fourier.py:116-137.  This is a compatibility
path, one frame at a time and an order of magnitude slower per pixel than the fused one; there is still no CPU
fallback."""
from __future__ import annotations

import numpy as np
import torch

TWO_PI = 2.0 * np.pi


def unwrap_scan(w: torch.Tensor) -> torch.Tensor:
    """Row/column path unwrap on the device, anchored at the centre pixel (same path as the fused kernels: RowDemod +
    RowLink in csrc/fcd_kernels.cuh).  Exact where the wrapped phase has no residues."""
    n0, n1 = w.shape
    jr = torch.zeros_like(w, dtype=torch.int64)
    jr[:, 1:] = torch.round((w[:, 1:] - w[:, :-1]) / TWO_PI).to(torch.int64)
    c = torch.cumsum(jr, dim=1)
    c = c - c[:, n1 // 2:n1 // 2 + 1]
    col = w[:, n1 // 2]
    jc = torch.zeros(n0, dtype=torch.int64, device=w.device)
    jc[1:] = torch.round((col[1:] - col[:-1]) / TWO_PI).to(torch.int64)
    m = torch.cumsum(jc, dim=0)
    m = m - m[n0 // 2]
    return w - TWO_PI * (c + m[:, None]).to(w.dtype)


class GenericPipeline:
    """Per-reference state of a bound GENERIC plan plus the per-frame float64 path."""

    def __init__(self, plan):
        from .engine import wavenumber
        self.plan = plan
        dev = plan.device
        H, W = plan.shape
        self.masks = [plan.carrier_mask(i) for i in range(2)]                       # bool, un-shifted layout
        self.ccsgn = [plan.carrier_ccsgn(i, complex128=True) for i in range(2)]
        self.reference = plan._reference.to(torch.float64)
        cal = plan.calibration_factor
        # integrate_in_fourier's meshes with its quirks (fourier.py:128-132): k2 before the zeroing, k2[0,0] = 1,
        # index N//2 + 1 zeroed for even sizes only
        ky = np.repeat(wavenumber(H, cal)[:, None], W, axis=1)
        kx = np.repeat(wavenumber(W, cal)[None, :], H, axis=0)
        k2 = kx ** 2 + ky ** 2
        k2[0, 0] = 1
        if W % 2 == 0:
            kx[:, W // 2 + 1] = 0
        if H % 2 == 0:
            ky[H // 2 + 1, :] = 0
        self.kx, self.ky, self.k2 = (torch.from_numpy(a).to(dev) for a in (kx, ky, k2))

    def phases_of(self, frame64: torch.Tensor, mode: int):
        plan = self.plan
        F = plan.fft2_c128(frame64.to(torch.complex128))                             # fcd.py:28
        out, guided = [], False
        for i in range(2):
            g = plan.fft2_c128(F * self.masks[i], inverse=True)
            ang = -torch.angle(g * self.ccsgn[i])                                    # fcd.py:118
            if mode in (1, 2, 3):
                use_guided = mode == 2
                w32 = ang.to(torch.float32)
                if mode == 3:
                    use_guided = plan.count_residues(w32)[0] != 0
                if use_guided:
                    # the integer field of the device unwrap (float32 input), applied to the float64 angles
                    k = torch.round((plan.unwrap_phase(w32) - w32) / TWO_PI).to(torch.float64)
                    ang = ang + TWO_PI * k
                    guided = True
                else:
                    ang = unwrap_scan(ang)
            out.append(ang)
        return torch.stack(out), guided

    def height_of(self, phases: torch.Tensor) -> torch.Tensor:
        plan = self.plan
        f0, f1 = plan.carrier_frequencies()
        det = f0[1] * f1[0] - f0[0] * f1[1]                                          # fcd.py:134-137
        if det == 0:
            raise ValueError("carriers are collinear (singular 2x2 system)")
        u = (f1[0] * phases[0] - f0[0] * phases[1]) / det
        v = (f0[1] * phases[1] - f1[1] * phases[0]) / det
        gx, gy = -u / plan.height, -v / plan.height                                  # fcd.py:32
        gxh = plan.fft2_c128(gx.to(torch.complex128))
        gyh = plan.fft2_c128(gy.to(torch.complex128))
        hhat = (-1.0j * self.kx * gxh + -1.0j * self.ky * gyh) / self.k2             # fourier.py:135
        return plan.fft2_c128(hhat, inverse=True).real.contiguous()

    def execute(self, frames: torch.Tensor, out, phases, mask, mode: int):
        n = frames.shape[0]
        guided_frames = []
        for f in range(n):
            fr = frames[f].to(torch.float64)
            mk = None
            if mask is not None:
                mk = (mask[f] if mask.dim() == 3 else mask) != 0
                fr = torch.where(mk, self.reference, fr)                             # analyze.py:231
            ph, guided = self.phases_of(fr, mode)
            h = self.height_of(ph)
            if mk is not None:
                h = h * (~mk)                                                        # analyze.py:255
            out[f].copy_(h)
            if phases is not None:
                phases[f].copy_(ph)
            if guided:
                guided_frames.append(f)
        return guided_frames
