"""CPU: analyze.mask / analyze.center -- oracle vs reference-derived goldens, and the CUDA kernels
in CPU emulation vs both (bit-exact: these are byte / integer results)."""
import os

import numpy as np
import pytest

from oracle import mask_oracle as mo
from tests.emul_lib import EmulPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PICS = "/root/reference/examples/Pictures"


@pytest.fixture(scope="module")
def gmask():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_mask.npz"))


def case(g, i):
    shape = tuple(int(v) for v in g[f"case{i}.shape"])
    img = mo.synthetic_structure(shape, int(g[f"case{i}.seed"]))
    m = np.unpackbits(g[f"case{i}.mask"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
    return img, int(g[f"case{i}.smoothed"]), m, tuple(int(v) for v in g[f"case{i}.center"])


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_oracle_matches_reference_golden(gmask, i):
    img, smoothed, m, c = case(gmask, i)
    assert np.array_equal(mo.mask(img, smoothed), m)
    assert mo.center(m) == c


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_emulated_kernels_bit_exact(gmask, i):
    img, smoothed, m, c = case(gmask, i)
    plan = EmulPlan(img.shape)
    flipped = np.ascontiguousarray(img[::-1, ::-1])
    got = plan.structure_mask(np.stack([img, flipped, img, img, flipped]), smoothed)     # two chunks of the kernel batch
    assert np.array_equal(got[0], m) and np.array_equal(got[2], m) and np.array_equal(got[3], m)
    assert np.array_equal(got[1], mo.mask(flipped, smoothed)) and np.array_equal(got[4], got[1])
    centers = plan.mask_center(got)
    assert centers[0] == c and centers[2] == c and centers[1] == mo.center(got[1])
    plan.close()


def test_degenerate_masks():
    plan = EmulPlan((64, 64))
    full = np.ones((64, 64), bool)
    ring = np.zeros((64, 64), bool); ring[10:50, 10:50] = True; ring[20:40, 22:44] = False
    touching = np.zeros((64, 64), bool); touching[0:30, 5:40] = True      # ~mask is one region that touches the border
    two = ring.copy(); two[12:16, 12:16] = False                           # a second, smaller hole
    got = plan.mask_center(np.stack([full, ring, touching, two]))
    assert got[0] == (-1, -1) and got[2] == (-1, -1)
    assert got[1] == mo.center(ring) == (29, 32)
    assert got[3] == mo.center(two) == (29, 32)
    with pytest.raises(UnboundLocalError):
        mo.center(full)
    img = np.full((64, 64), 7.0, np.float32)                               # nothing below the mean -> empty mask
    assert not plan.structure_mask(img, 5).any()
    plan.close()


@pytest.mark.skipif(not os.path.isdir(REF_PICS), reason="reference fixtures not on this machine")
def test_real_fixture_against_the_reference_itself():
    """examples/mask_example.py: frame 6 of the mask series, smoothed=15, through the unmodified
    reference class, the oracle and the emulated kernels."""
    import cv2
    from oracle.ref_shims import import_reference_analyze
    analyze = import_reference_analyze()
    path = os.path.join(REF_PICS, "mask", "0_5mm_circular_100ms_20250529_141304_C1S0001000006.tif")
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED).astype(np.float32)
    m_ref = analyze.mask(img, smoothed=15)
    c_ref = analyze.center(m_ref)
    assert np.array_equal(mo.mask(img, 15), m_ref) and mo.center(m_ref) == tuple(c_ref)
    plan = EmulPlan(img.shape)
    m = plan.structure_mask(img, 15)[0]
    assert np.array_equal(m, m_ref)
    assert plan.mask_center(m)[0] == tuple(c_ref)
    plan.close()


def test_random_masks_stress_connected_components():
    """Speckle at several densities: many small regions, diagonal-only links, ties in area."""
    rng = np.random.default_rng(42)
    plan = EmulPlan((64, 64), 1)
    batch, want = [], []
    for trial in range(120):
        dens = rng.choice([0.35, 0.5, 0.6, 0.75, 0.9])
        m = rng.random((64, 64)) < dens
        if trial % 3 == 0:
            m[20:44, 20:44] = True; m[26:38, 27:39] = rng.random((12, 12)) < 0.3     # a cavity with debris
        batch.append(m)
        try:
            want.append(mo.center(m))
        except UnboundLocalError:
            want.append((-1, -1))
    got = plan.mask_center(np.stack(batch))
    assert got == want
    # largest-region selection incl. ties: feed images whose smooth field is the mask itself (size-1 filter)
    for trial in range(40):
        img = (rng.random((64, 64)) < rng.choice([0.4, 0.55, 0.7])).astype(np.float32)
        assert np.array_equal(plan.structure_mask(img, 1)[0], mo.mask(img, 1))
    plan.close()


def test_division_by_the_window_size_is_exact():
    """The box filters replace tmp / size by a multiply and two FMAs (csrc/fcd_mask.cuh::div_by_const); the result
    has to be the division's, bit for bit, for every window size: sums of camera-like float32 values, doubles with
    random significands over 80 binades, and neighbours of exact multiples of the size."""
    import ctypes
    from tests.emul_lib import lib
    L = lib()
    L.fcd_emul_div_mismatches.restype = ctypes.c_longlong
    L.fcd_emul_div_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_double]
    rng = np.random.default_rng(7)
    n = 200_000
    for size in list(range(1, 65)) + [97, 127, 255, 1000]:
        sums8 = rng.integers(0, 256, (n, min(size, 64))).astype(np.float32).astype(np.float64).sum(1)
        sums16 = (rng.integers(0, 65536, (n, min(size, 64))).astype(np.float32) + rng.random((n, 1), np.float32)).astype(np.float64).sum(1)
        unit = rng.random((n, min(size, 64)), np.float32).astype(np.float64).sum(1)
        wide = np.ldexp(1.0 + rng.random(n), rng.integers(-40, 40, n)) * rng.choice([-1.0, 1.0], n)
        mult = np.nextafter(rng.integers(0, 10**6, n).astype(np.float64) * size, rng.choice([-np.inf, np.inf], n))
        a = np.ascontiguousarray(np.concatenate([sums8, sums16, unit, wide, mult, [0.0]]))
        assert L.fcd_emul_div_mismatches(a.ctypes.data, a.size, float(size)) == 0, size


@pytest.mark.parametrize("size", [1, 2, 3, 8, 14, 15, 16, 31, 32, 33, 47])
def test_emulated_box_filter_is_scipys_uniform_filter(size):
    """smooth = uniform_filter(image, size) (pydata/analyze.py:70): the emulated kernels reproduce scipy's float32
    output exactly -- the field the threshold and np.mean then work on."""
    import ctypes
    from scipy.ndimage import uniform_filter
    from tests.emul_lib import lib
    L = lib()
    L.fcd_emul_box_filter.restype = None
    L.fcd_emul_box_filter.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4
    rng = np.random.default_rng(size)
    H, W = 64, 128
    imgs = np.stack([rng.integers(0, 1024, (H, W)).astype(np.float32),                  # 10-bit camera counts
                     rng.random((H, W), np.float32),
                     (rng.random((H, W), np.float32) < 0.5).astype(np.float32) * 255.0])
    tmp, out = np.empty_like(imgs), np.empty_like(imgs)
    L.fcd_emul_box_filter(imgs.ctypes.data, tmp.ctypes.data, out.ctypes.data, len(imgs), H, W, size)
    for k in range(len(imgs)):
        want = uniform_filter(imgs[k], size)
        assert want.dtype == np.float32
        assert np.array_equal(out[k].view(np.uint32), want.view(np.uint32)), (size, k)


# ---- warp-level model of LabelInit's device path (csrc/fcd_mask.cuh) --------------------------------------------
# The CPU emulation runs LabelInit's sequential branch; its device branch (ballots, 16 mask bytes per lane gathered
# into bitmap words by a multiply and one shuffle, ragged trips for narrow images) is mirrored here lane by lane so
# that its bit logic is checked for every supported width, including the 64- and 128-pixel rows no GPU test covers.
def _ballot(pred):  # list of 32 bools -> word
    w=0
    for l,p in enumerate(pred):
        if p: w|=1<<l
    return w
def _vcmpeq4_zero(w):
    r=0
    for k in range(4):
        if ((w>>(8*k))&0xff)==0: r|=0xff<<(8*k)
    return r
def _nib(w): return ((((_vcmpeq4_zero(w)&0x01010101)*0x01020408)&0xffffffff)>>24)&15
def _model(mask_row=None, smooth_row=None, thr=0.0, W=64, r=3):
    """returns (bits words, dict of L writes {col: value})"""
    L={}; bits=[None]*(W//32); carry=[0]
    def emit(b,c0):
        starts=b & ~(((b<<1)&0xffffffff)|carry[0])
        for lane in range(32):
            if (starts>>lane)&1: L[c0+lane]=r*W+c0+lane
        carry[0]=b>>31
    if smooth_row is not None:
        for c0 in range(0,W,128):
            v=[[ (smooth_row[c0+32*q+lane] if c0+32*q<W else thr) for lane in range(32)] for q in range(4)]
            mine=[0]*32
            for q in range(4):
                if c0+32*q<W:
                    b=_ballot([v[q][lane]<thr for lane in range(32)])
                    emit(b,c0+32*q)
                    mine[q]=b
            for lane in range(4):
                if c0+32*lane<W: bits[(c0>>5)+lane]=mine[lane]
    else:
        for c0 in range(0,W,512):
            vs=[]
            for lane in range(32):
                inn=c0+16*lane<W
                if inn:
                    by=mask_row[c0+16*lane:c0+16*lane+16]
                    m=[int.from_bytes(bytes(by[4*k:4*k+4]),'little') for k in range(4)]
                else: m=[0x01010101]*4
                v=_nib(m[0])|(_nib(m[1])<<4)|(_nib(m[2])<<8)|(_nib(m[3])<<12)
                v=(v<<(16*(lane&1)))&0xffffffff
                vs.append(v)
            vs2=[vs[l]|vs[l^1] for l in range(32)]
            for q in range(16):
                b=vs2[2*q]
                if c0+32*q<W: emit(b,c0+32*q)
            for lane in range(32):
                if not (lane&1) and c0+16*lane<W: bits[(c0>>5)+(lane>>1)]=vs2[lane]
    return bits,L

@pytest.mark.parametrize("W", [64, 128, 256, 512, 1024, 2048])
def test_label_init_device_path_bit_logic(W):
    rng = np.random.default_rng(W)
    for trial in range(20):
        fg = rng.random(W) < rng.choice([0.0, 0.1, 0.5, 0.9, 1.0])
        want_bits = [sum(int(fg[32 * w + i]) << i for i in range(32)) for w in range(W // 32)]
        want_starts = {c: 3 * W + c for c in range(W) if fg[c] and not (c > 0 and fg[c - 1])}
        mask = np.where(fg, 0, rng.integers(1, 256, W)).astype(np.uint8)                   # mode 1: foreground = !mask
        assert _model(mask_row=mask, W=W) == (want_bits, want_starts)
        smooth = np.where(fg, 0.2, 0.8).astype(np.float32)                                  # mode 0: smooth < threshold
        assert _model(smooth_row=smooth, thr=0.5, W=W) == (want_bits, want_starts)


def _box_rows_warp_model(a, size):
    """One lane's row through BoxRowsWarp's device path (csrc/fcd_mask.cuh): ring of three 32-column tiles, tile k + 2
    parked while tile k is filtered, interior tiles without the reflection lookup.  The emulation runs the kernel's
    sequential branch; this mirrors the tile bookkeeping of the device branch for every window the kernel takes."""
    n = len(a); s1 = size // 2; s2 = size - s1 - 1; ntiles = n // 32
    T = [None, None, None]

    def at(j):                                   # 'reflect', then the ring
        if j < 0: j = -j - 1
        if j >= n: j = 2 * n - 1 - j
        return float(T[(j >> 5) % 3][j & 31])

    T[0] = a[0:32].copy()
    if ntiles > 1: T[1] = a[32:64].copy()
    tmp = 0.0
    for l in range(size): tmp += at(l - s1)
    out, slot, d = np.empty(n, np.float32), 0, float(size)
    for k in range(ntiles):
        more = k + 2 < ntiles
        parked = a[32 * (k + 2):32 * (k + 3)].copy() if more else None
        if 0 < k < ntiles - 1:
            tc, tn, tp = T[slot], T[0 if slot == 2 else slot + 1], T[2 if slot == 0 else slot - 1]
            for cc in range(32):
                jn, jo = cc + s2, cc - 1 - s1
                tmp += (float(tc[jn]) if jn < 32 else float(tn[jn - 32])) - (float(tc[jo]) if jo >= 0 else float(tp[jo + 32]))
                out[32 * k + cc] = np.float32(tmp / d)
        else:
            for cc in range(32):
                l = 32 * k + cc
                if l > 0: tmp += at(l + s2) - at(l - 1 - s1)
                out[l] = np.float32(tmp / d)
        if more: T[2 if slot == 0 else slot - 1] = parked
        slot = 0 if slot == 2 else slot + 1
    return out


@pytest.mark.parametrize("n", [64, 128, 256])
def test_box_rows_warp_tile_bookkeeping(n):
    from scipy.ndimage import uniform_filter1d
    rng = np.random.default_rng(n)
    for size in range(1, 33):
        a = rng.integers(0, 4096, n).astype(np.float32) + rng.random(n).astype(np.float32)
        want = uniform_filter1d(a, size)
        assert np.array_equal(_box_rows_warp_model(a, size).view(np.uint32), want.view(np.uint32)), size
