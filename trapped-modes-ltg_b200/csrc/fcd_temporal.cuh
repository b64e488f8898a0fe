// fcd_temporal.cuh -- temporal harmonic analysis of a stack of height maps.
//
// Reference: analyze.block_amplitude (pydata/analyze.py:542-641) with analyze.block_split
// (analyze.py:365-417).  For one spatial block the reference loads every map file again,
// stacks the block over time, masks the pixels that are zero in the first map (NaN),
//     fft_vals = np.fft.fft(maps, axis=-1)[..., fft_freqs >= 0]
// estimates f0 (when not given) from the peak of nanmean(|fft_vals|) over the block, and reads
// amplitude (|X|/N at the first index, 2|X|/N after) and phase (angle X) of `mode` bins per pixel.
//
// Here the stack stays where the height maps were produced ([N][H][W] float32 in HBM) and all
// blocks are handled in one pass:
//   TemporalSpectrum  per pixel a length-N transform along time (strided "column" FFT with the
//                     same Stockham core as the spatial kernels; any other N through Bluestein's
//                     chirp convolution on a power-of-two length L >= 2N-1), |X[k]| summed per
//                     spatial block in registers of persistent thread blocks -> mean spectra.
//                     The spectrum itself never goes to memory.
//   HarmonicAccumulate  X[k_j] for the few bins the caller asks for, as a streaming sum over
//                     frames (float64 accumulators): reads every map exactly once, works for any
//                     N, and is additive over frame shards (multi-GPU: all-reduce of the sums).
//   HarmonicFinalize  amplitude / phase planes in the reference's (ny, nx, mode+1) layout.
#pragma once
#include "fcd_mask.cuh"

namespace fcd {

constexpr double kPiT = 3.14159265358979323846;

FCD_HD double nan_f64() {
    union { unsigned long long u; double d; } c;
    c.u = 0x7ff8000000000000ull;
    return c.d;
}

// spatial block of pixel (r, c): blocks are bs x bs pixels in a grid of brows x bcols blocks (the
// reference uses bs = H // blocks_per_row for both axes of a square grid; a row band of a
// frame-sharded stack holds fewer block rows); -1 outside the blocked area
FCD_HD int temporal_block_of(int r, int c, int bs, int brows, int bcols) {
    const int bi = r / bs, bj = c / bs;
    return (bi < brows && bj < bcols) ? bi * bcols + bj : -1;
}

struct TemporalSpecParams {
    const float* maps;     // [N][H][W]
    const float* first;    // [H][W] first map: pixels equal to zero there are excluded (NaN in the reference); may be null
    const cf* tw;          // twiddle table of the transform length L
    const cf* chirp;       // [N] exp(+i pi t^2 / N)          (Bluestein only)
    const cf* bspec;       // [L] FFT_L(chirp filter) / L      (Bluestein only)
    double* mean;          // [blocks][npos] sum of |X[k]| over the block's valid pixels (zero-initialised)
    float zero;            // subtracted from every map (analyze.py:585)
    int N, npos, H, W, bs, brows, bcols;
};

template <int L, int G, bool BLUE>
struct TemporalSpectrum {
    using FF = Fft<L, -1, float>;
    using FI = Fft<L, +1, float>;
    using Params = TemporalSpecParams;
    static constexpr bool BLOCKED_TILES = true;    // contiguous tiles per thread block: few flushes
    static constexpr bool PIPELINED = true;        // tile link: next tile for staging, and when to flush
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = L / 16, THREADS = G * TPF, PHASES = BLUE ? 8 : 4;
    static constexpr int STRIDE = GroupLayout<L, G>::STRIDE;   // skewed: lanes of a warp sit in different groups
    using TW = SmemTwiddles<FF, THREADS>;
    // + thread-private staging slots for the next tile's 16 samples (cp.async one tile ahead, so
    // the strided DRAM reads overlap the previous tile's transform)
    static constexpr int STAGE_OFF = TW::TW_BYTES + G * STRIDE * (int)sizeof(cf);
    static constexpr int SMEM_BYTES = STAGE_OFF + 16 * THREADS * (int)sizeof(float);
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    template <int PH, class P> FCD_HD static bool enabled(const P&, const unsigned char*, int) { return true; }
    struct State { cf v[16]; double acc[16]; TileLink link; };

    FCD_HD static void stage(const Params& p, int bx, int by, int g, int t, float* slots) {
        long long px;
        pixel_of(p, bx, by, g, px);
        const long long plane = (long long)p.H * p.W;
        FCD_UNROLL
        for (int m = 0; m < 16; ++m) {
            const int tt = t + TPF * m;
            if (tt < p.N) async_copy4(slots + m * THREADS, p.maps + (long long)tt * plane + px);
        }
    }

    // tile (bx, by): by = spatial block, bx = (row in block) * (bs / G) + column tile
    FCD_HD static bool pixel_of(const Params& p, int bx, int by, int g, long long& px) {
        const int tiles_per_row = p.bs / G;
        const int r = (by / p.bcols) * p.bs + bx / tiles_per_row;
        const int c = (by % p.bcols) * p.bs + (bx % tiles_per_row) * G + g;
        px = (long long)r * p.W + c;
        return p.first == nullptr || p.first[px] != 0.0f;
    }
    FCD_HD static void accumulate(const Params& p, int bx, int by, int g, int t, State& st) {
        long long px;
        const bool valid = pixel_of(p, bx, by, g, px);
        FCD_UNROLL
        for (int m = 0; m < 16; ++m) {
            const int k = t + TPF * m;
            if (valid && k < p.npos) {
                cf x = st.v[m];
                if constexpr (BLUE) x = x * conj(p.chirp[k]);
                st.acc[m] += (double)sqrtf(x.x * x.x + x.y * x.y);
            }
        }
        // leaving this spatial block (or the last tile of this thread block): flush the sums
        if (!st.link.has_next || st.link.next_by != by) {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int k = t + TPF * m;
                if (k < p.npos && st.acc[m] != 0.0) atomic_add_f64(p.mean + (long long)by * p.npos + k, st.acc[m]);
                st.acc[m] = 0.0;
            }
        }
    }

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        const int g = tid % G, t = tid / G;      // group fastest: a warp reads G adjacent pixels of 32/G frames
        cf* s = reinterpret_cast<cf*>(smem_all + TW::TW_BYTES) + g * STRIDE;
        if constexpr (PH == 0) {
            if (st.link.first) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) st.acc[m] = 0.0;
            }
            float* slots = reinterpret_cast<float*>(smem_all + STAGE_OFF) + tid;
            if (st.link.first) stage(p, bx, by, g, t, slots);      // later tiles were staged in phase 1
            async_wait_all();
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int tt = t + TPF * m;
                cf a = mk<float>(0.f, 0.f);
                if (tt < p.N) {
                    const float x = slots[m * THREADS] - p.zero;
                    a = BLUE ? scale(conj(p.chirp[tt]), x) : mk<float>(x, 0.f);
                }
                st.v[m] = a;
            }
            FF::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            if (st.link.has_next)      // own slots were consumed before the barrier
                stage(p, st.link.next_bx, st.link.next_by, g, t, reinterpret_cast<float*>(smem_all + STAGE_OFF) + tid);
            FF::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 2) {
            FF::stepC(st.v, t, s);
        } else if constexpr (PH == 3) {
            FF::stepD(st.v, t, s, tw);
            if constexpr (BLUE) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) st.v[m] = st.v[m] * p.bspec[t + TPF * m];
            } else {
                accumulate(p, bx, by, g, t, st);
            }
        } else if constexpr (PH == 4) {
            FI::stepA(st.v, t, s);
        } else if constexpr (PH == 5) {
            FI::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 6) {
            FI::stepC(st.v, t, s);
        } else {
            FI::stepD(st.v, t, s, tw);
            accumulate(p, bx, by, g, t, st);
        }
    }
};

// ---- long series: N = N1 * N2 as two levels of chirp-convolution transforms ------------------------------------
// X[k1 + N1 k2] = sum_{n2} W_N^{n2 k1} [ sum_{n1} x[n1 N2 + n2] W_N1^{n1 k1} ] W_N2^{n2 k2}
// Stage 0: for every n2 a length-N1 transform over the frames n1*N2 + n2, times W_N^{n2 k1}, to a workspace
//          ws[(k1*N2 + n2)][pixel of the chunk];  stage 1: for every k1 a length-N2 transform over n2 whose
//          magnitudes go to mean[block][k1 + N1*k2].  A chunk is a range of spatial blocks whose N * pixels
//          complex workspace fits the budget.  Both levels use Bluestein on a power-of-two length >= 2*Ni - 1.
struct TemporalTwoLevelParams {
    const float* maps;     // [N][H][W]
    const float* first;    // [H][W] or null
    cf* ws;                // [N][pixels of the chunk]
    const cf* tw;          // engine twiddles of this stage's length L
    const cf* chirp;       // [Ni]  exp(+i pi m^2 / Ni)
    const cf* bspec;       // [L]
    const cf* twn;         // [N]   exp(-2 pi i m / N)
    double* mean;          // [blocks][npos]
    float zero;
    int N, N1, N2, npos, H, W, bs, brows, bcols;
    int b0;                // first spatial block of the chunk
    int chunk_pixels;      // pixels per frame in the workspace = blocks in chunk * bs * bs
};

template <int L, int G, int STAGE>
struct TemporalTwoLevel {
    using FF = Fft<L, -1, float>;
    using FI = Fft<L, +1, float>;
    using Params = TemporalTwoLevelParams;
    static constexpr bool BLOCKED_TILES = true;
    static constexpr bool PIPELINED = true;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = L / 16, THREADS = G * TPF, PHASES = 8;
    static constexpr int STRIDE = GroupLayout<L, G>::STRIDE;
    using TW = SmemTwiddles<FF, THREADS>;
    static constexpr int SMEM_BYTES = TW::TW_BYTES + G * STRIDE * (int)sizeof(cf);
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    template <int PH, class P> FCD_HD static bool enabled(const P&, const unsigned char*, int) { return true; }
    struct State { cf v[16]; double acc[16]; TileLink link; };

    // stage 0 tiles: bx = pixel tile + tiles_per_block * n2, by = block of the chunk
    // stage 1 tiles: bx = pixel tile,                         by = block of the chunk * N1 + k1
    FCD_HD static void decode(const Params& p, int bx, int by, int g, int& blk, int& other, long long& px, int& pc) {
        const int tpb = p.bs * (p.bs / G);                  // pixel tiles per spatial block
        const int ptile = STAGE == 0 ? bx % tpb : bx;
        other = STAGE == 0 ? bx / tpb : by % p.N1;          // n2 (stage 0) or k1 (stage 1)
        const int brel = STAGE == 0 ? by : by / p.N1;
        blk = p.b0 + brel;
        const int tiles_per_row = p.bs / G;
        const int r = (blk / p.bcols) * p.bs + ptile / tiles_per_row;
        const int c = (blk % p.bcols) * p.bs + (ptile % tiles_per_row) * G + g;
        px = (long long)r * p.W + c;
        pc = brel * p.bs * p.bs + ptile * G + g;
    }

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        const int g = tid % G, t = tid / G;
        cf* s = reinterpret_cast<cf*>(smem_all + TW::TW_BYTES) + g * STRIDE;
        const int M = STAGE == 0 ? p.N1 : p.N2;             // this stage's transform length
        int blk, other; long long px; int pc;
        decode(p, bx, by, g, blk, other, px, pc);
        if constexpr (PH == 0) {
            if (st.link.first) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) st.acc[m] = 0.0;
            }
            const long long plane = (long long)p.H * p.W;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int sidx = t + TPF * m;
                cf a = mk<float>(0.f, 0.f);
                if (sidx < M) {
                    if constexpr (STAGE == 0) {
                        const float x = p.maps[((long long)sidx * p.N2 + other) * plane + px] - p.zero;
                        a = scale(conj(p.chirp[sidx]), x);
                    } else {
                        a = p.ws[((long long)other * p.N2 + sidx) * p.chunk_pixels + pc] * conj(p.chirp[sidx]);
                    }
                }
                st.v[m] = a;
            }
            FF::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            FF::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 2) {
            FF::stepC(st.v, t, s);
        } else if constexpr (PH == 3) {
            FF::stepD(st.v, t, s, tw);
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) st.v[m] = st.v[m] * p.bspec[t + TPF * m];
        } else if constexpr (PH == 4) {
            FI::stepA(st.v, t, s);
        } else if constexpr (PH == 5) {
            FI::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 6) {
            FI::stepC(st.v, t, s);
        } else {
            FI::stepD(st.v, t, s, tw);
            if constexpr (STAGE == 0) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    const int k1 = t + TPF * m;
                    if (k1 < p.N1) {
                        const cf y = st.v[m] * conj(p.chirp[k1]);
                        const long long q = ((long long)other * k1) % p.N;
                        p.ws[((long long)k1 * p.N2 + other) * p.chunk_pixels + pc] = y * p.twn[q];
                    }
                }
            } else {
                const bool valid = p.first == nullptr || p.first[px] != 0.0f;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    const int k2 = t + TPF * m;
                    const long long k = (long long)other + (long long)p.N1 * k2;
                    if (valid && k2 < p.N2 && k < p.npos) {
                        const cf x = st.v[m] * conj(p.chirp[k2]);
                        st.acc[m] += (double)sqrtf(x.x * x.x + x.y * x.y);
                    }
                }
                if (!st.link.has_next || st.link.next_by != by) {      // leaving this (block, k1): flush
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) {
                        const long long k = (long long)other + (long long)p.N1 * (t + TPF * m);
                        if (t + TPF * m < p.N2 && k < p.npos && st.acc[m] != 0.0)
                            atomic_add_f64(p.mean + (long long)blk * p.npos + k, st.acc[m]);
                        st.acc[m] = 0.0;
                    }
                }
            }
        }
    }
};

// number of valid pixels (first map != 0) per spatial block
struct BlockValidParams {
    const float* first;   // may be null: every pixel valid
    int* counts;          // [blocks], zero-initialised
    int H, W, bs, brows, bcols;
    long long total;      // (bs * brows) * bcols: one thread per row segment inside a block
};
struct BlockValidCount : ElemBase {
    using Params = BlockValidParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int r = (int)(i / p.bcols), bj = (int)(i % p.bcols);
        int n = p.bs;
        if (p.first) {
            n = 0;
            const float* row = p.first + (long long)r * p.W + (long long)bj * p.bs;
            for (int c = 0; c < p.bs; ++c) n += row[c] != 0.0f;
        }
        atomic_add_i32(p.counts + (r / p.bs) * p.bcols + bj, n);
    }
};

constexpr int kMaxHarmonicBins = 8;

struct HarmonicAccParams {
    const float* maps;    // [nf][H][W]
    const double* tw;     // [nf][blocks][nb][2]: cos, -sin of 2 pi k t / N for the frame's global time index
    double* acc;          // [nb][2][H*W]: real and imaginary planes
    float zero;
    int nf, nb, H, W, bs, brows, bcols, init;
    long long total;      // H * W
};
// cos / -sin of 2 pi k t / N for every (frame of the chunk, block, bin), phase reduced exactly in integers
struct HarmonicTwParams {
    const int* bins;      // [blocks][nb] (device)
    double* tw;           // [nf][blocks][nb][2]
    int nf, nblk, nb, t0, n_total;
    long long total;      // nf * blocks * nb
};
struct HarmonicTwiddles : ElemBase {
    using Params = HarmonicTwParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int per = p.nblk * p.nb;
        const int f = (int)(i / per), bj = (int)(i % per);
        const long long q = ((long long)p.bins[bj] * (p.t0 + f)) % p.n_total;
        const double ang = 2.0 * kPiT * (double)q / (double)p.n_total;
        p.tw[2 * i] = cos(ang);
        p.tw[2 * i + 1] = -sin(ang);
    }
};

struct alignas(16) dbl2 { double c, s; };
template <int NB>
struct HarmonicAccumulate : ElemBase {
    using Params = HarmonicAccParams;
    static constexpr int MIN_BLOCKS = 4;     // 64 registers: the twiddle loads hit L1, only the map loads need depth
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int b = temporal_block_of((int)(i / p.W), (int)(i % p.W), p.bs, p.brows, p.bcols);
        if (b < 0) {          // outside the block grid: nothing accumulates, but `init` still defines the sums
            if (p.init) {
                FCD_UNROLL
                for (int j = 0; j < 2 * NB; ++j) p.acc[(long long)j * p.total + i] = 0.0;
            }
            return;
        }
        double re[NB], im[NB];
        FCD_UNROLL
        for (int j = 0; j < NB; ++j) {
            re[j] = p.init ? 0.0 : p.acc[(long long)(2 * j) * p.total + i];
            im[j] = p.init ? 0.0 : p.acc[(long long)(2 * j + 1) * p.total + i];
        }
        const int nblk = p.brows * p.bcols;
        const dbl2* __restrict__ tw = reinterpret_cast<const dbl2*>(p.tw) + (long long)b * NB;
        const float* __restrict__ src = p.maps + i;
        constexpr int U = 8;           // frames per batch: the batch's loads are issued before its arithmetic
        int f0 = 0;
        for (; f0 + U <= p.nf; f0 += U) {
            float x[U];
            FCD_UNROLL
            for (int u = 0; u < U; ++u) x[u] = src[(long long)(f0 + u) * p.total];
            FCD_UNROLL
            for (int u = 0; u < U; ++u) {
                const double xd = (double)(x[u] - p.zero);
                const dbl2* w = tw + (long long)(f0 + u) * nblk * NB;
                FCD_UNROLL
                for (int j = 0; j < NB; ++j) { const dbl2 ww = w[j]; re[j] += xd * ww.c; im[j] += xd * ww.s; }
            }
        }
        for (; f0 < p.nf; ++f0) {
            const double xd = (double)(src[(long long)f0 * p.total] - p.zero);
            const dbl2* w = tw + (long long)f0 * nblk * NB;
            FCD_UNROLL
            for (int j = 0; j < NB; ++j) { const dbl2 ww = w[j]; re[j] += xd * ww.c; im[j] += xd * ww.s; }
        }
        FCD_UNROLL
        for (int j = 0; j < NB; ++j) {
            p.acc[(long long)(2 * j) * p.total + i] = re[j];
            p.acc[(long long)(2 * j + 1) * p.total + i] = im[j];
        }
    }
};

struct HarmonicFinParams {
    const double* acc;     // [nb][2][H*W]
    const float* first;    // may be null
    double* amps;          // [H][W][nb+1]
    double* phases;        // [H][W][nb+1]
    int nb, n_total;
    long long total;
};
struct HarmonicFinalize : ElemBase {
    using Params = HarmonicFinParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const bool valid = p.first == nullptr || p.first[i] != 0.0f;
        const double nan = nan_f64();
        double* a = p.amps + i * (p.nb + 1);
        double* ph = p.phases + i * (p.nb + 1);
        for (int j = 0; j < p.nb; ++j) {
            const double re = p.acc[(long long)(2 * j) * p.total + i], im = p.acc[(long long)(2 * j + 1) * p.total + i];
            const double mag = hypot(re, im) / (double)p.n_total;
            a[j] = valid ? (j == 0 ? mag : 2.0 * mag) : nan;       // analyze.py:634-637
            ph[j] = valid ? atan2(im, re) : nan;                    // analyze.py:638
        }
        a[p.nb] = 0.0;      // the reference allocates mode+1 planes and fills mode of them
        ph[p.nb] = 0.0;
    }
};

}  // namespace fcd
