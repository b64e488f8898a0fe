// fcd_mask.cuh -- floating-structure mask and cavity centre on the device.
//
// Reference: analyze.mask (pydata/analyze.py:43-100) and analyze.center (pydata/analyze.py:104-140):
//   smooth = scipy.ndimage.uniform_filter(image, size); Mask = smooth < np.mean(smooth);
//   mask = largest 8-connected region of Mask;  centre = int(centroid) of the largest 8-connected
//   region of ~mask whose bounding box does not touch the border.
// These are byte / integer results, so the arithmetic in front of the threshold is reproduced
// bit for bit:
//   * uniform_filter (scipy 1.18): per line a float64 running sum `tmp += new - old`, output
//     (float)(tmp / size), axis 0 first then axis 1, 'reflect' boundary, window [-(size/2), size-size/2-1];
//   * np.mean of a float32 array: pairwise summation in float32 -- 128-element blocks with eight
//     interleaved accumulators, then a binary tree (exact for the power-of-two sizes supported).
// Connected components: lock-free union-find on pixel indices; the root of a component is its
// smallest raster index, so "first label" ties resolve like skimage.measure.label.
#pragma once
#include "fcd_generic.cuh"

namespace fcd {

// H, W (and so n = H*W) are powers of two (plan constraint): pixel indices decompose with masks and
// shifts; a 64-bit division by a runtime value per pixel costs more than the rest of these kernels
FCD_HD int ilog2_pow2(int v) {
#if defined(__CUDA_ARCH__)
    return __ffs(v) - 1;
#else
    return __builtin_ctz((unsigned)v);
#endif
}

// tmp / size, bit for bit, without the division routine (a reciprocal estimate, Newton steps and a slow path per
// pixel -- half the instructions of the box filters in round 1).  With r = RN(1 / d) computed once per thread:
//     q = RN(a * r);  rem = a - q * d  (exact in one FMA);  RN(q + rem * r)
// is the correctly rounded quotient (Markstein's theorem: r correctly rounded, q within an ulp of a / d; it is the
// last step of the hardware routine itself).  d is a small integer and a a finite sum of float32 values here, so
// neither the exceptional divisors of the theorem (significand all ones) nor underflow of the remainder can occur.
// tests/test_mask_oracle_and_emulation.py compares it with the division on 10^7 sums for every window size.
FCD_HD double div_by_const(double a, double d, double r) {
    const double q = a * r;
    const double rem = fma(-q, d, a);
    return fma(rem, r, q);
}

struct ElemBase : NoPrologue {
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
};

// ---- box filter along one axis: one thread per line ---------------------------------------
struct BoxLinesParams {
    const float* in;
    float* out;
    int H, W, size, axis;
    long long n_lines;        // frames * (axis == 0 ? W : H)
};
struct BoxLines : ElemBase {
    using Params = BoxLinesParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long line = (long long)bx * THREADS + tid;
        if (line >= p.n_lines) return;
        const int per = p.axis == 0 ? p.W : p.H;
        const long long frame = line / per;
        const int pos = (int)(line % per);
        const int n = p.axis == 0 ? p.H : p.W;
        const long long stride = p.axis == 0 ? p.W : 1;
        const long long base = frame * p.H * p.W + (p.axis == 0 ? pos : (long long)pos * p.W);
        const float* __restrict__ a = p.in + base;
        float* __restrict__ o = p.out + base;
        const int s1 = p.size / 2, s2 = p.size - s1 - 1;
        const double dsize = (double)p.size, rsize = 1.0 / dsize;
        auto at = [&](int j) -> double {           // 'reflect': d c b a | a b c d | d c b a
            if (j < 0) j = -j - 1;
            if (j >= n) j = 2 * n - 1 - j;
            return (double)a[(long long)j * stride];
        };
        double tmp = 0.0;
        for (int l = 0; l < p.size; ++l) tmp += at(l - s1);
        o[0] = (float)div_by_const(tmp, dsize, rsize);
        // interior: no boundary handling, loads independent of the running sum -> unrolled
        int l = 1;
        for (; l < n && l - 1 - s1 < 0; ++l) {
            const double d = at(l + s2) - at(l - 1 - s1);
            tmp += d;
            o[(long long)l * stride] = (float)div_by_const(tmp, dsize, rsize);
        }
        const int l_hi = n - s2;      // l + s2 < n
        for (; l + 8 <= l_hi; l += 8) {
            double nv[8], ov[8];
            FCD_UNROLL
            for (int q = 0; q < 8; ++q) {
                nv[q] = (double)a[(long long)(l + q + s2) * stride];
                ov[q] = (double)a[(long long)(l + q - 1 - s1) * stride];
            }
            FCD_UNROLL
            for (int q = 0; q < 8; ++q) {
                const double d = nv[q] - ov[q];
                tmp += d;
                o[(long long)(l + q) * stride] = (float)div_by_const(tmp, dsize, rsize);
            }
        }
        for (; l < n; ++l) {
            const double d = at(l + s2) - at(l - 1 - s1);
            tmp += d;
            o[(long long)l * stride] = (float)div_by_const(tmp, dsize, rsize);
        }
    }
};

// ---- box filter along the rows (axis 1), window <= 32: one warp per 32 rows --------------------------------
// A thread per row (BoxLines) walks 32 different cache lines per warp step.  Here the warp moves 32 x 32
// tiles through shared memory instead: rows are loaded and stored 128 contiguous bytes at a time, lane l then
// runs row l's sequential float64 recurrence (exactly BoxLines' operation order) out of the tile ring.
struct BoxRowsWarp : NoPrologue {
    using Params = BoxLinesParams;      // axis is 1; n_lines = frames * H (a multiple of 32)
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 128, PHASES = 1;
    static constexpr int TILE = 32 * 33;                                   // padded 32 x 32 floats
    static constexpr int SMEM_BYTES = (THREADS / 32) * 4 * TILE * (int)sizeof(float);   // ring of 3 + output tile per warp
    struct State { int dummy; };

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State&) {
        const long long wid = (long long)bx * (THREADS / 32) + tid / 32;  // global warp = block of 32 rows
        const int lane = tid % 32;
        if (wid * 32 >= p.n_lines) return;
        const int n = p.W;
        const int s1 = p.size / 2, s2 = p.size - s1 - 1;
        const double dsize = (double)p.size, rsize = 1.0 / dsize;
#if defined(__CUDA_ARCH__)
        const float* __restrict__ a = p.in + wid * 32 * n;                // 32 consecutive rows (same frame: H % 32 == 0)
        float* __restrict__ o = p.out + wid * 32 * n;
        float* T = reinterpret_cast<float*>(smem) + (tid / 32) * 4 * TILE;
        float* O = T + 3 * TILE;
        const int ntiles = n / 32;
        auto load_tile = [&](int k) {
            float* t = T + (k % 3) * TILE;
            FCD_UNROLL
            for (int rr = 0; rr < 32; ++rr) t[rr * 33 + lane] = a[(long long)rr * n + 32 * k + lane];
        };
        auto at = [&](int j) -> double {           // 'reflect', then this lane's row out of the ring
            if (j < 0) j = -j - 1;
            if (j >= n) j = 2 * n - 1 - j;
            return (double)T[((j >> 5) % 3) * TILE + lane * 33 + (j & 31)];
        };
        load_tile(0);
        if (ntiles > 1) load_tile(1);
        __syncwarp();
        double tmp = 0.0;
        for (int l = 0; l < p.size; ++l) tmp += at(l - s1);
        int slot = 0;                                                     // k % 3
        for (int k = 0; k < ntiles; ++k) {
            // tile k + 2 is requested now and parked in registers: its DRAM round trip runs under this tile's recurrence
            float nx[32];
            const bool more = k + 2 < ntiles;
            if (more) {
                FCD_UNROLL
                for (int rr = 0; rr < 32; ++rr) nx[rr] = a[(long long)rr * n + 32 * (k + 2) + lane];
            }
            if (k > 0 && k + 1 < ntiles) {
                // interior tile (size <= 32): the entering sample sits in tile k or k + 1, the leaving one in tile
                // k - 1 or k, no reflection -- two address selects per pixel instead of the general lookup
                const float* tc = T + slot * TILE + lane * 33;                              // tile k
                const float* tn = T + (slot == 2 ? 0 : slot + 1) * TILE + lane * 33 - 32;   // tile k + 1, column - 32
                const float* tp = T + (slot == 0 ? 2 : slot - 1) * TILE + lane * 33 + 32;   // tile k - 1, column + 32
                FCD_UNROLL
                for (int cc = 0; cc < 32; ++cc) {
                    const int jn = cc + s2, jo = cc - 1 - s1;
                    const double nv = (double)(jn < 32 ? tc : tn)[jn];
                    const double ov = (double)(jo >= 0 ? tc : tp)[jo];
                    const double d = nv - ov;
                    tmp += d;
                    O[lane * 33 + cc] = (float)div_by_const(tmp, dsize, rsize);
                }
            } else {
                for (int cc = 0; cc < 32; ++cc) {
                    const int l = 32 * k + cc;
                    if (l > 0) {
                        const double d = at(l + s2) - at(l - 1 - s1);
                        tmp += d;
                    }
                    O[lane * 33 + cc] = (float)div_by_const(tmp, dsize, rsize);
                }
            }
            __syncwarp();
            FCD_UNROLL
            for (int rr = 0; rr < 32; ++rr) o[(long long)rr * n + 32 * k + lane] = O[rr * 33 + lane];
            if (more) {                                    // replaces tile k - 1, which no later column reads
                float* t = T + (slot == 0 ? 2 : slot - 1) * TILE;
                FCD_UNROLL
                for (int rr = 0; rr < 32; ++rr) t[rr * 33 + lane] = nx[rr];
            }
            __syncwarp();
            slot = slot == 2 ? 0 : slot + 1;
        }
#else
        // sequential emulation: the same recurrence, one row per emulated thread
        const float* __restrict__ a = p.in + (wid * 32 + lane) * n;
        float* __restrict__ o = p.out + (wid * 32 + lane) * n;
        auto at = [&](int j) -> double {
            if (j < 0) j = -j - 1;
            if (j >= n) j = 2 * n - 1 - j;
            return (double)a[j];
        };
        double tmp = 0.0;
        for (int l = 0; l < p.size; ++l) tmp += at(l - s1);
        for (int l = 0; l < n; ++l) {
            if (l > 0) {
                const double d = at(l + s2) - at(l - 1 - s1);
                tmp += d;
            }
            o[l] = (float)div_by_const(tmp, dsize, rsize);
        }
        (void)smem;
#endif
    }
};

// ---- np.mean(float32): block sums of 128 with eight accumulators, then a binary tree ------
FCD_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
struct PairBlockParams {
    const float* in;     // [frames][n]
    float* out;          // [frames][n / 128]
    long long n_blocks;  // frames * n / 128
};
// Eight lanes per 128-element block, lane k owning numpy's accumulator r[k] (elements k, k+8, ...): a warp
// step reads four full 32-byte sectors instead of 32 scattered words, and the final
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is three butterfly shuffles (float addition is commutative, so
// every lane of a pair holds the same partial sum).  n_threads = 8 * n_blocks.
struct PairBlockSum : ElemBase {
    using Params = PairBlockParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long item = (long long)bx * THREADS + tid;
        const long long b = item >> 3;
        const int k = (int)(item & 7);
        if (b >= p.n_blocks) return;                 // n_blocks is a multiple of 4: whole warps leave together
        const float* __restrict__ a = p.in + b * 128;
#if defined(__CUDA_ARCH__)
        float r = a[k];
        FCD_UNROLL
        for (int i = 8; i < 128; i += 8) r = fadd_rn(r, a[i + k]);
        r = fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (k == 0) p.out[b] = r;
#else
        if (k != 0) return;                          // sequential emulation: one thread does the block
        float r[8];
        for (int q = 0; q < 8; ++q) r[q] = a[q];
        for (int i = 8; i < 128; i += 8)
            for (int q = 0; q < 8; ++q) r[q] = fadd_rn(r[q], a[i + q]);
        p.out[b] = fadd_rn(fadd_rn(fadd_rn(r[0], r[1]), fadd_rn(r[2], r[3])),
                           fadd_rn(fadd_rn(r[4], r[5]), fadd_rn(r[6], r[7])));
#endif
    }
};
struct PairTreeParams {
    const float* in;     // [frames][2 * n_out]
    float* out;          // [frames][n_out]
    long long total;     // frames * n_out
};
struct PairTree : ElemBase {
    using Params = PairTreeParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i < p.total) p.out[i] = fadd_rn(p.in[2 * i], p.in[2 * i + 1]);
    }
};

// ---- connected components -------------------------------------------------------------------
FCD_HD int uf_find(const int* L, int i) {
    int r = L[i];
    while (r != i) { i = r; r = L[i]; }
    return r;
}
FCD_HD void uf_unite(int* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }      // a < b: hang b under a
#if defined(__CUDA_ARCH__)
        const int old = atomicMin(&L[b], a);
#else
        const int old = L[b];
        if (a < old) L[b] = a;
#endif
        if (old == b) return;
        b = old;
    }
}

// Labels live at RUN STARTS only (round 2).  A foreground pixel belongs to the horizontal run it sits in, the run is
// named by its first pixel, and which pixels are foreground is a bitmap (1 bit per pixel): the run start of any pixel
// is found from the bitmap alone.  The union-find array L is therefore only ever written and read at run-start
// pixels (everything else in it is never touched), and every later pass works on bitmap words and runs instead of
// 4-byte labels per pixel.
FCD_HD int count_lz(unsigned v) {               // leading zeros of a non-zero word
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}
FCD_HD int count_tz(unsigned v) {               // trailing zeros of a non-zero word
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// first column of the run that contains the foreground pixel at column c of the line whose bitmap words are B
FCD_HD int run_start_col(const unsigned* B, int c) {
    int wi = c >> 5;
    const int q = c & 31;
    const unsigned below = q ? (~B[wi] & ((1u << q) - 1u)) : 0u;      // background pixels left of c inside its word
    if (below) return (wi << 5) + 32 - count_lz(below);
    while (wi > 0) {                                                    // the run reaches back into earlier words
        const unsigned z = ~B[wi - 1];
        if (z) return z >> 31 ? (wi << 5) : ((wi - 1) << 5) + 32 - count_lz(z);
        --wi;
    }
    return 0;
}

struct LabelInitParams {
    const float* smooth;      // mode 0: foreground = smooth < sum[frame] / n
    const float* sums;        // [frames] pairwise float32 sums
    const uint8_t* mask;      // mode 1: foreground = !mask
    int* L;                   // [frames][n]: union-find parent, defined at run-start pixels only (initially itself)
    unsigned* bits;           // [frames][H][W/32]: foreground bit per pixel (bit i = column 32*word + i)
    long long n_rows;         // frames * H lines, one warp each (32 * n_rows threads)
    int H, W;
    int mode;
};
// One warp per line: 32 consecutive pixels per step (coalesced), the foreground word from a ballot; the pixels
// whose left neighbour is background start a run and become their own union-find roots.
struct LabelInit : ElemBase {
    using Params = LabelInitParams;
    FCD_HD static bool fg_at(const Params& p, long long o, int c, float thr) {
        return p.mode == 0 ? (p.smooth[o + c] < thr) : (p.mask[o + c] == 0);
    }
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long item = (long long)bx * THREADS + tid;
        const long long row = item >> 5;
        const int lane = (int)(item & 31);
        if (row >= p.n_rows) return;               // whole warps: THREADS is a multiple of 32
        const long long f = row >> ilog2_pow2(p.H);
        const int r = (int)(row & (p.H - 1));
        const int n = p.H * p.W;
        const long long o = row * p.W;
        const float thr = p.mode == 0 ? p.sums[f] / (float)n : 0.f;   // np.mean: float32 sum / count
#if defined(__CUDA_ARCH__)
        // Several loads per lane are in flight before the first ballot: one 128-byte request per warp and trip left the
        // kernel waiting on DRAM latency (round 2, ncu launch list: 630 us for 268 MB of floats, 400 us for 67 MB of bytes).
        unsigned carry = 0u;                        // last pixel of the previous word is foreground
        auto emit = [&](unsigned bits, int c0) {    // warp-uniform word of 32 pixels starting at column c0
            const unsigned starts = bits & ~((bits << 1) | carry);
            if ((starts >> lane) & 1u) p.L[o + c0 + lane] = r * p.W + c0 + lane;
            carry = bits >> 31;
        };
        if (p.mode == 0) {
            const float* __restrict__ a = p.smooth + o;
            for (int c0 = 0; c0 < p.W; c0 += 128) {                 // four words per trip (W is a multiple of 64)
                float v[4];
                FCD_UNROLL
                for (int q = 0; q < 4; ++q) v[q] = (c0 + 32 * q < p.W) ? a[c0 + 32 * q + lane] : thr;
                unsigned mine = 0u;
                FCD_UNROLL
                for (int q = 0; q < 4; ++q) {
                    if (c0 + 32 * q < p.W) {
                        const unsigned bits = __ballot_sync(0xffffffffu, v[q] < thr);
                        emit(bits, c0 + 32 * q);
                        if (lane == q) mine = bits;
                    }
                }
                if (lane < 4 && c0 + 32 * lane < p.W) p.bits[((o + c0) >> 5) + lane] = mine;
            }
        } else if ((reinterpret_cast<uintptr_t>(p.mask) & 15u) == 0) {
            // 16 mask bytes per lane: 512 pixels per trip.  Lane l covers columns c0 + 16 l .. + 15, i.e. one half of
            // bitmap word l / 2; the two halves meet in one shuffle.
            const uint4* __restrict__ m16 = reinterpret_cast<const uint4*>(p.mask + o);
            for (int c0 = 0; c0 < p.W; c0 += 512) {
                const bool in = c0 + 16 * lane < p.W;
                uint4 m = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);      // beyond the line: background
                if (in) m = m16[(c0 >> 4) + lane];
                // byte == 0 -> foreground bit: per 32-bit word a 0 / 1 flag per byte, gathered into a nibble by a multiply
                auto nib = [](unsigned w) { return (((__vcmpeq4(w, 0u) & 0x01010101u) * 0x01020408u) >> 24) & 15u; };
                unsigned v = nib(m.x) | (nib(m.y) << 4) | (nib(m.z) << 8) | (nib(m.w) << 12);
                v <<= 16 * (lane & 1);
                v |= __shfl_xor_sync(0xffffffffu, v, 1);             // both lanes of a pair hold word lane / 2
                FCD_UNROLL
                for (int q = 0; q < 16; ++q) {
                    const unsigned bits = __shfl_sync(0xffffffffu, v, 2 * q);
                    if (c0 + 32 * q < p.W) emit(bits, c0 + 32 * q);
                }
                if (!(lane & 1) && in) p.bits[((o + c0) >> 5) + (lane >> 1)] = v;
            }
        } else {                                    // a caller's mask that is not 16-byte aligned: byte loads
            for (int c0 = 0; c0 < p.W; c0 += 32) {
                const unsigned bits = __ballot_sync(0xffffffffu, p.mask[o + c0 + lane] == 0);
                emit(bits, c0);
                if (lane == 0) p.bits[(o + c0) >> 5] = bits;
            }
        }
#else
        for (int c = lane; c < p.W; c += 32)        // sequential emulation: every emulated thread looks left on its own
            if (fg_at(p, o, c, thr) && !(c > 0 && fg_at(p, o, c - 1, thr))) p.L[o + c] = r * p.W + c;
        if (lane == 0)
            for (int c0 = 0; c0 < p.W; c0 += 32) {
                unsigned bits = 0;
                for (int q = 0; q < 32; ++q) bits |= (fg_at(p, o, c0 + q, thr) ? 1u : 0u) << q;
                p.bits[(o + c0) >> 5] = bits;
            }
#endif
    }
};
struct LabelMergeParams {
    int* L;
    const unsigned* bits;     // [frames][H][W/32] from LabelInit
    long long total;          // frames * H * (W / 32): one thread per 32-pixel word
    int H, W;
};
// 8-connectivity links to the previous row, once per place where two runs first touch.  Works on the
// foreground bit words: the three "first touch" conditions are bit expressions over the word, the word
// above and the edge bits of the four neighbouring words, and only their set bits reach the union-find.
struct LabelMerge : ElemBase {
    using Params = LabelMergeParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int wpr = p.W >> 5;                                   // words per row (a power of two)
        const int wc = (int)(i & (wpr - 1));
        const long long row = i >> ilog2_pow2(wpr);
        const int r = (int)(row & (p.H - 1));
        if (r == 0) return;
        const unsigned C = p.bits[i];
        if (C == 0u) return;
        const unsigned N = p.bits[i - wpr];
        const unsigned Cl = wc > 0 ? p.bits[i - 1] : 0u, Cr = wc + 1 < wpr ? p.bits[i + 1] : 0u;
        const unsigned Nl = wc > 0 ? p.bits[i - wpr - 1] : 0u, Nr = wc + 1 < wpr ? p.bits[i - wpr + 1] : 0u;
        const unsigned Wm = (C << 1) | (Cl >> 31), Em = (C >> 1) | (Cr << 31);       // west / east neighbour is foreground
        const unsigned NWm = (N << 1) | (Nl >> 31), NEm = (N >> 1) | (Nr << 31);
        unsigned up = C & N & ~(Wm & NWm);          // first pixel of an overlap with the run above
        unsigned dl = C & ~N & NWm & ~Wm;           // diagonal touch on the left
        unsigned dr = C & ~N & NEm & ~Em;           // diagonal touch on the right
        const int n = p.H * p.W;
        int* L = p.L + (row >> ilog2_pow2(p.H)) * n;
        const unsigned* Bc = p.bits + row * wpr;                      // this line's bitmap and the one above
        const unsigned* Bn = Bc - wpr;
        const int c0 = wc << 5;
        // the union-find lives on run starts: unite the runs the two touching pixels belong to
        auto link = [&](int c, int cn) {
            uf_unite(L, r * p.W + run_start_col(Bc, c), (r - 1) * p.W + run_start_col(Bn, cn));
        };
        while (up) { const int q = count_tz(up); up &= up - 1u; link(c0 + q, c0 + q); }
        while (dl) { const int q = count_tz(dl); dl &= dl - 1u; link(c0 + q, c0 + q - 1); }
        while (dr) { const int q = count_tz(dr); dr &= dr - 1u; link(c0 + q, c0 + q + 1); }
    }
};

FCD_HD void atomic_add_i32(int* a, int v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(a, v);
#else
    *a += v;
#endif
}
FCD_HD void atomic_min_i32(int* a, int v) {
#if defined(__CUDA_ARCH__)
    atomicMin(a, v);
#else
    if (v < *a) *a = v;
#endif
}
FCD_HD void atomic_max_i32(int* a, int v) {
#if defined(__CUDA_ARCH__)
    atomicMax(a, v);
#else
    if (v > *a) *a = v;
#endif
}
FCD_HD void atomic_add_u64(unsigned long long* a, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(a, v);
#else
    *a += v;
#endif
}

// per-root statistics (arrays indexed by the root's pixel index, per frame; only roots are ever touched, so only
// those entries are initialised)
struct RegionStats {
    int* area;                      // zero-initialised
    int* minr; int* maxr; int* minc; int* maxc;     // initialised to +big / -1 (only with bbox)
    unsigned long long* sumr; unsigned long long* sumc;
};
struct RootStatsInitParams {
    const int* L;
    const unsigned* bits;
    RegionStats st;
    long long total;          // frames * H * (W / 32): one thread per 32-pixel word
    int H, W;
    int with_bbox;
};
struct RootStatsInit : ElemBase {
    using Params = RootStatsInitParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const unsigned w = p.bits[i];
        if (w == 0u) return;
        const int wpr = p.W >> 5;
        const int wc = (int)(i & (wpr - 1));
        const long long row = i >> ilog2_pow2(wpr);
        const int n = p.H * p.W;
        const long long fo = (row >> ilog2_pow2(p.H)) * n;
        const int px0 = (int)(row & (p.H - 1)) * p.W + (wc << 5);
        const unsigned carry = wc > 0 ? (p.bits[i - 1] >> 31) : 0u;
        unsigned starts = w & ~((w << 1) | carry);
        while (starts) {
            const int px = px0 + count_tz(starts);
            starts &= starts - 1u;
            if (p.L[fo + px] != px) continue;                        // not a root
            p.st.area[fo + px] = 0;
            if (p.with_bbox) {
                p.st.minr[fo + px] = 0x7fffffff; p.st.minc[fo + px] = 0x7fffffff;
                p.st.maxr[fo + px] = -1; p.st.maxc[fo + px] = -1;
                p.st.sumr[fo + px] = 0ull; p.st.sumc[fo + px] = 0ull;
            }
        }
    }
};

struct LabelFlattenParams {
    int* L;
    const unsigned* bits;   // [frames][H][W/32] foreground bit per pixel (LabelInit)
    RegionStats st;
    long long n_rows;       // frames * H lines, one warp each (32 * n_rows threads)
    int H, W;
    int with_bbox;
};
FCD_HD int trailing_ones(unsigned v) {          // number of consecutive set bits from bit 0
#if defined(__CUDA_ARCH__)
    return v == 0xffffffffu ? 32 : __ffs((int)~v) - 1;
#else
    int n = 0;
    while (n < 32 && ((v >> n) & 1u)) ++n;
    return n;
#endif
}
// Run-based flatten (round 2).  Every foreground pixel already points at the first pixel of its horizontal run
// (LabelInit) and the merges only ever re-point run starts, so a region's statistics are sums over RUNS and the only
// labels that need flattening are the run starts: pixel -> run start -> root is then two hops for whoever needs a
// pixel's region (MaskOut).  The kernel therefore never touches the label plane as a whole -- it reads the foreground
// bitmap (1 bit per pixel instead of 4 bytes), and the lane that finds a run's first bit in its word walks the
// following words for the run's length, finds the root once, points the run start at it and adds the run to the
// region's area / bounding box / coordinate sums: one set of atomics per run, no per-pixel traffic.  (Round 1 walked
// the 4-byte labels of every line word by word and rewrote them all: 2.3 ms per 64-frame wave of 2048^2, 88 % of it
// waiting on those loads.)  Plain code, no warp intrinsics: the CPU emulation runs the same lines.
struct LabelFlatten : ElemBase {
    using Params = LabelFlattenParams;
    FCD_HD static void flush(const Params& p, long long fo, int root, int r, int cnt, int c0, int c1, long long sc) {
        if (cnt == 0) return;
        atomic_add_i32(p.st.area + fo + root, cnt);
        if (p.with_bbox) {
            atomic_min_i32(p.st.minr + fo + root, r);
            atomic_max_i32(p.st.maxr + fo + root, r);
            atomic_min_i32(p.st.minc + fo + root, c0);
            atomic_max_i32(p.st.maxc + fo + root, c1);
            atomic_add_u64(p.st.sumr + fo + root, (unsigned long long)cnt * (unsigned long long)r);
            atomic_add_u64(p.st.sumc + fo + root, (unsigned long long)sc);
        }
    }
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long item = (long long)bx * THREADS + tid;
        const long long row = item >> 5;
        const int lane = (int)(item & 31);
        if (row >= p.n_rows) return;
        const int n = p.H * p.W;
        const long long fo = (row >> ilog2_pow2(p.H)) * n;
        const int r = (int)(row & (p.H - 1));
        int* L = p.L + fo;
        const int wpr = p.W >> 5;                                   // bitmap words per line
        const unsigned* B = p.bits + row * wpr;
        for (int wi = lane; wi < wpr; wi += 32) {
            const unsigned w = B[wi];
            if (w == 0u) continue;
            const unsigned carry = wi > 0 ? (B[wi - 1] >> 31) : 0u;
            unsigned starts = w & ~((w << 1) | carry);                // first pixels of runs inside this word
            while (starts) {
                const int q = count_tz(starts);
                starts &= starts - 1u;
                int len = trailing_ones(w >> q);
                if (q + len == 32) {                                  // the run goes on in the following words
                    for (int wj = wi + 1; wj < wpr; ++wj) {
                        const int t = trailing_ones(B[wj]);
                        len += t;
                        if (t < 32) break;
                    }
                }
                const int c0 = (wi << 5) + q;
                const int px = r * p.W + c0;
                const int root = uf_find(L, px);
                // other lines may be walking through L[px] to their roots: writing the root keeps every chain valid
                L[px] = root;
                flush(p, fo, root, r, len, c0, c0 + len - 1, (long long)len * c0 + (long long)len * (len - 1) / 2);
            }
        }
    }
};

// largest region per frame: key = area << 32 | ~root  (ties -> smallest root = first label)
struct LargestParams {
    const int* L;
    const unsigned* bits;
    RegionStats st;
    unsigned long long* best;     // [frames], zero-initialised
    long long total;              // frames * H * (W / 32): one thread per 32-pixel word
    int H, W;
    int holes_only;               // keep regions whose bounding box stays off the border
};
struct LargestRegion : ElemBase {
    using Params = LargestParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const unsigned w = p.bits[i];
        if (w == 0u) return;
        const int wpr = p.W >> 5;
        const int wc = (int)(i & (wpr - 1));
        const long long row = i >> ilog2_pow2(wpr);
        const int n = p.H * p.W;
        const long long f = row >> ilog2_pow2(p.H), fo = f * n;
        const int px0 = (int)(row & (p.H - 1)) * p.W + (wc << 5);
        const unsigned carry = wc > 0 ? (p.bits[i - 1] >> 31) : 0u;
        unsigned starts = w & ~((w << 1) | carry);
        while (starts) {
            const int px = px0 + count_tz(starts);
            starts &= starts - 1u;
            if (p.L[fo + px] != px) continue;                        // not a root
            if (p.holes_only) {
                const bool inside = p.st.minr[fo + px] > 0 && p.st.minc[fo + px] > 0 &&
                                    p.st.maxr[fo + px] + 1 < p.H && p.st.maxc[fo + px] + 1 < p.W;
                if (!inside) continue;
            }
            const unsigned long long key = ((unsigned long long)(unsigned)p.st.area[fo + px] << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (unsigned)px);
            atomic_max_u64(p.best + f, key);
        }
    }
};

struct MaskOutParams {
    const int* L;
    const unsigned* bits;
    const unsigned long long* best;
    uint8_t* mask;
    long long total;              // frames * H * (W / 32): one thread per 32-pixel word = 32 output bytes
    int H, W;
};
struct alignas(16) u32x4 { unsigned a, b, c, d; };
// One thread per bitmap word: the runs that cross the word are looked up once each (run start -> root, which
// LabelFlatten left there), the 32 mask bytes go out as two 16-byte stores.
struct MaskOut : ElemBase {
    using Params = MaskOutParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int wpr = p.W >> 5;
        const int wc = (int)(i & (wpr - 1));
        const long long row = i >> ilog2_pow2(wpr);
        const int n = p.H * p.W;
        const long long f = row >> ilog2_pow2(p.H), fo = f * n;
        const int r = (int)(row & (p.H - 1));
        const unsigned long long key = p.best[f];
        const int root = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        const unsigned w = p.bits[i];
        unsigned in = 0u;                                            // pixels of the word inside the chosen region
        if (w != 0u && key != 0ull) {
            const unsigned* B = p.bits + row * wpr;
            unsigned rem = w;
            while (rem) {
                const int q = count_tz(rem);
                const int len = trailing_ones(w >> q);
                const unsigned seg = (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << q;
                const int start = run_start_col(B, (wc << 5) + q);
                if (p.L[fo + r * p.W + start] == root) in |= seg;
                rem &= ~seg;
            }
        }
        unsigned ww[8];
        FCD_UNROLL
        for (int k = 0; k < 8; ++k) {
            const unsigned nib = (in >> (4 * k)) & 15u;               // four pixels -> four bytes 0 / 1
            ww[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        uint8_t* dst = p.mask + i * 32;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            u32x4 lo, hi;
            lo.a = ww[0]; lo.b = ww[1]; lo.c = ww[2]; lo.d = ww[3];
            hi.a = ww[4]; hi.b = ww[5]; hi.c = ww[6]; hi.d = ww[7];
            u32x4* out = reinterpret_cast<u32x4*>(dst);
            out[0] = lo;
            out[1] = hi;
        } else {                                                      // a caller's mask buffer need not be 16-byte aligned
            for (int k = 0; k < 32; ++k) dst[k] = (uint8_t)((in >> k) & 1u);
        }
    }
};

struct CenterOutParams {
    const unsigned long long* best;
    RegionStats st;
    int* centers;        // [frames][2] = (cy, cx) or (-1, -1)
    int frames;
    int n;
};
struct CenterOut : ElemBase {
    using Params = CenterOutParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const int f = bx * THREADS + tid;
        if (f >= p.frames) return;
        const unsigned long long key = p.best[f];
        if (key == 0ull) { p.centers[2 * f] = -1; p.centers[2 * f + 1] = -1; return; }
        const int root = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        const long long o = (long long)f * p.n + root;
        const double area = (double)p.st.area[o];
        // regionprops centroid = mean of the integer coordinates (exact sums), then int()
        p.centers[2 * f] = (int)((double)p.st.sumr[o] / area);
        p.centers[2 * f + 1] = (int)((double)p.st.sumc[o] / area);
    }
};

struct FillI32Params { int* a; int v; long long total; };
struct FillI32 : ElemBase {
    using Params = FillI32Params;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i < p.total) p.a[i] = p.v;
    }
};

}  // namespace fcd
