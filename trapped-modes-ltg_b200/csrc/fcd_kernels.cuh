// fcd_kernels.cuh -- the fused FCD height-map pipeline as phase-structured kernels.
//
// Reference path being replaced (file:line in /root/reference):
//   pyfcd/fcd.py:28        displaced_fft = fft2(displaced)                       -> K1 + K2
//   pyfcd/fcd.py:118       -angle(ifft2(displaced_fft*mask) * ccsgn)             -> K2 + K3
//   pyfcd/fcd.py:119       unwrap_phase                                          -> K3 + K3b (+ K4 col 0)
//   pyfcd/fcd.py:123-138   2x2 carrier solve                                     -> K4 (folded coefficients)
//   pyfcd/fcd.py:32        height_gradient = -displacement/height                -> K4 (folded)
//   pyfcd/fourier.py:116-137 integrate_in_fourier (+ remove_degeneracy 76-92)    -> K3 (row fwd) + K4 + K5
//
// A "kernel" is a struct with compile-time THREADS / PHASES / SMEM_BYTES, a Params struct,
// a per-thread State (registers that live across barriers) and
//     template<int PH> static void phase(params, block x, block y, tid, smem, state)
// Consecutive phases are separated by a block-wide barrier.  On the GPU the phases are
// inlined into one __global__ function (fcd_launch.cuh); tests/emul runs the very same
// phase bodies thread-by-thread on the CPU.
//
// Thread organisation: a block holds G groups of TPF = L/16 threads; each group owns one
// transform at a time (see fft_core.cuh for the natural strided ownership t + TPF*m).
#pragma once
#include <cmath>
#include <cstring>
#include "fft_core.cuh"

namespace fcd {

constexpr float kTwoPiF = 6.28318530717958647692f;
constexpr float kInvTwoPiF = 0.15915494309189533577f;

FCD_HD int imin(int a, int b) { return a < b ? a : b; }

// 1/x as one MUFU.RCP.  Without .ftz the compiler wraps every reciprocal in a denormal rescue
// (FMUL by 2^24, FSETP, FSEL, predicated FMUL: SASS of profiles r01_m3); callers keep x >= 1e-30.
FCD_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// Branch-free atan2 (degree-7 minimax in a^2; max abs error 3e-7 rad, about one ulp of pi in float32).  The
// library atan2f compiles to calls and divergent slow paths, which stops the compiler from
// batching the loads around it (ncu profiles/r01: one exposed DRAM round trip per element).
// np.angle(0) = 0 is preserved.
FCD_HD float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);   // (0, 0) -> a = 0 -> angle 0
    const float a = mn * fast_rcp(mx);
    const float s = a * a;
    float r = -0.004054565913975239f;
    r = r * s + 0.021862953901290894f;
    r = r * s + -0.0559123195707798f;
    r = r * s + 0.0964219719171524f;
    r = r * s + -0.1390862911939621f;
    r = r * s + 0.19946566224098206f;
    r = r * s + -0.33329859375953674f;
    r = r * s + 0.9999993443489075f;
    r = r * a;
    r = (ay > ax) ? 1.57079632679489661923f - r : r;
    r = (x < 0.f) ? 3.14159265358979323846f - r : r;
    return copysignf(r, y);
}

// Both carriers of a pixel at once: the polynomial and the wrap run as packed FMUL2 / FFMA2 chains
// (.x = carrier 0, .y = carrier 1), the range reduction and the quadrant fix-ups stay scalar (FMNMX,
// MUFU, predicated FADD have no packed form).  Same arithmetic per lane as fast_atan2f + the scalar
// wrap of RowDemod::demod (what tests/emul executes).   z0, z1: the two complex samples;  th: the
// two reference angles;  returns (phi0, phi1) = -wrap(angle(z_i) + th_i)      (fcd.py:118)
#if defined(FCD_PACKED_F32)
__device__ __forceinline__ cf demod_pair(cf z0, cf z1, float th0, float th1) {
    const float ax0 = fabsf(z0.x), ay0 = fabsf(z0.y), ax1 = fabsf(z1.x), ay1 = fabsf(z1.y);
    const float mx0 = fmaxf(fmaxf(ax0, ay0), 1e-30f), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(fmaxf(ax1, ay1), 1e-30f), mn1 = fminf(ax1, ay1);
    const float2 a = __fmul2_rn(make_float2(mn0, mn1), make_float2(fast_rcp(mx0), fast_rcp(mx1)));
    const float2 s = __fmul2_rn(a, a);
    float2 r = make_float2(-0.004054565913975239f, -0.004054565913975239f);
    r = __ffma2_rn(r, s, make_float2(0.021862953901290894f, 0.021862953901290894f));
    r = __ffma2_rn(r, s, make_float2(-0.0559123195707798f, -0.0559123195707798f));
    r = __ffma2_rn(r, s, make_float2(0.0964219719171524f, 0.0964219719171524f));
    r = __ffma2_rn(r, s, make_float2(-0.1390862911939621f, -0.1390862911939621f));
    r = __ffma2_rn(r, s, make_float2(0.19946566224098206f, 0.19946566224098206f));
    r = __ffma2_rn(r, s, make_float2(-0.33329859375953674f, -0.33329859375953674f));
    r = __ffma2_rn(r, s, make_float2(0.9999993443489075f, 0.9999993443489075f));
    r = __fmul2_rn(r, a);
    float r0 = r.x, r1 = r.y;
    r0 = (ay0 > ax0) ? 1.57079632679489661923f - r0 : r0;
    r1 = (ay1 > ax1) ? 1.57079632679489661923f - r1 : r1;
    r0 = (z0.x < 0.f) ? 3.14159265358979323846f - r0 : r0;
    r1 = (z1.x < 0.f) ? 3.14159265358979323846f - r1 : r1;
    const float2 ang = __fadd2_rn(make_float2(copysignf(r0, z0.y), copysignf(r1, z1.y)), make_float2(th0, th1));
    const float2 q = __fmul2_rn(ang, make_float2(kInvTwoPiF, kInvTwoPiF));
    // 2 pi * rint(q) - ang as one FMA per lane, which is what the compiler contracts the scalar expression to
    return cf_of(__ffma2_rn(make_float2(rintf(q.x), rintf(q.y)), make_float2(kTwoPiF, kTwoPiF),
                            make_float2(-ang.x, -ang.y)));
}
#endif

// common helpers -------------------------------------------------------------------------
template <int L, int G, int BUFS = 1>
struct GroupLayout {
    static constexpr int TPF = L / 16;
    static constexpr int THREADS = G * TPF;
    // per-group exchange buffer(s); the skew spreads "lane = group" accesses over the banks
    static constexpr int SKEW = (G * BUFS > 1 && G * BUFS <= 16) ? 16 / (G * BUFS) : (G * BUFS > 16 ? 1 : 0);
    static constexpr int STRIDE = L + L / 16 + SKEW;  // elements per buffer
    static constexpr int GROUP_STRIDE = BUFS * STRIDE;
};

struct alignas(16) cf2 {
    cf a, b;
};

// w3 (row-transformed z = phi0 + i*phi1) is stored column-blocked so that the column kernel
// reads one contiguous tile:  w3[f][slot/4][y][slot%4],  slot(kc) = kc for kc <= W/2 and
// kc + 3 above (then the mirrored columns W-kc..W-kc-3 of a 4-column tile share one block).
FCD_HD int w3_slot(int kc, int W) { return kc + (kc > W / 2 ? 3 : 0); }
FCD_HD int w3_blocks(int W) { return W / 4 + 1; }
FCD_HD long long w3_index(int f, int kc, int y, int H, int W) {
    const int slot = w3_slot(kc, W);
    return (((long long)f * w3_blocks(W) + (slot >> 2)) * H + y) * 4 + (slot & 3);
}

// Blocks are persistent: they copy the twiddle table of their transform length into shared
// memory once (prologue) and then loop over tiles, so twiddle reads are LDS, not LDG.
template <class F, int THREADS>
struct SmemTwiddles {
    static constexpr int TW_BYTES = F::TW_ELEMS * (int)sizeof(cf);
    static_assert(TW_BYTES % 16 == 0, "twiddle block must keep 16-byte alignment");
    FCD_HD static void load(const cf* __restrict__ g, int tid, unsigned char* smem) {
        cf* s = reinterpret_cast<cf*>(smem);
        for (int i = tid; i < F::TW_ELEMS; i += THREADS) s[i] = g[i];
    }
};
// asynchronous 8-byte global -> shared copy (LDGSTS); the CPU emulation copies immediately
FCD_HD void async_copy8(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#else
    *reinterpret_cast<cf*>(smem_dst) = *reinterpret_cast<const cf*>(gsrc);
#endif
}
FCD_HD void async_copy4(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
#else
    *reinterpret_cast<float*>(smem_dst) = *reinterpret_cast<const float*>(gsrc);
#endif
}
FCD_HD void async_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
// ---- TMA (bulk async copy engine): one thread hands a contiguous global -> shared copy to the copy
// engine, completion is signalled on an mbarrier that the consumers poll.  In the CPU emulation the
// copy happens immediately and the barrier operations are no-ops.
using mbar_t = unsigned long long;
FCD_HD void mbar_init(mbar_t* bar, int count) {
#if defined(__CUDA_ARCH__)
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    *bar = 0;
    (void)count;
#endif
}
FCD_HD void mbar_expect_tx(mbar_t* bar, unsigned bytes) {
#if defined(__CUDA_ARCH__)
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#else
    (void)bar; (void)bytes;
#endif
}
// size and both addresses must be multiples of 16 bytes
FCD_HD void bulk_copy_g2s(void* smem_dst, const void* gsrc, unsigned bytes, mbar_t* bar) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
#else
    std::memcpy(smem_dst, gsrc, bytes);
    (void)bar;
#endif
}
FCD_HD void mbar_wait(mbar_t* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
#else
    (void)bar; (void)parity;
#endif
}
// orders this thread's earlier generic-proxy shared-memory accesses before later async-proxy (TMA) writes
FCD_HD void fence_proxy_async() {
#if defined(__CUDA_ARCH__)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// state the launcher maintains for kernels that prefetch the next tile (K::PIPELINED)
struct TileLink {
    int next_bx, next_by;
    bool has_next, first;
};

// phases may be skipped (together with their trailing barrier) when a kernel says so; the
// answer must be uniform over the barrier domain (block, or group for named barriers)
struct AllPhases {
    template <int PH, class P> FCD_HD static bool enabled(const P&, const unsigned char*, int) { return true; }
};
struct NoPrologue : AllPhases {
    template <class P> FCD_HD static void prologue(const P&, int, unsigned char*) {}
};

// =========================================================================================
// K1  row forward: two real rows packed into one complex FFT; emits only the columns that
//     fall inside the two carrier disks, transposed:  w1[f][i][c'][y]  (y contiguous)
// =========================================================================================
struct RowFwdParams {
    const void* frames;       // [F][H][W] float32 / uint8 / uint16 (frame_kind 0 / 1 / 2)
    int frame_kind;
    const float* reference;   // [H][W]   (mask substitution, may be null)
    const uint8_t* mask;      // [F][H][W] or [H][W] (mask_stride = 0), may be null
    long long mask_stride;    // elements between consecutive frames' masks
    cf* w1;                   // [F][2][ncp][H]
    const cf* tw;             // W_L table
    int H, ncp;
    int nc[2];                // columns per carrier
    int kc0[2];               // signed column frequency of c' = 0
};

template <int L, int G>
struct RowFwd : AllPhases {
    using F = Fft<L, -1, float, RegenRowPlan<L>>;     // same radix order and table as RowPlan<L> (fft_core.cuh, TwRegen)
    using GL = GroupLayout<L, G>;
    using Params = RowFwdParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = true;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = ((G * L / 16) <= 128 ? 6 : ((G * L / 16) <= 256 ? 3 : 1));
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 6;
    using TW = SmemTwiddles<F, THREADS>;
    static constexpr int SMEM_BYTES = TW::TW_BYTES + G * GL::STRIDE * (int)sizeof(cf);
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    struct State { cf v[16]; TileLink link; };

    // two image rows -> one complex sequence (mask substitution analyze.py:231 fused); camera
    // frames may be passed as uint8 / uint16 and are widened on load (analyze.load_image's astype)
    template <class PT>
    FCD_HD static void load_rows_t(const Params& p, int bx, int by, int g, int t, cf* v) {
        const int W = L;
        const int ya = (bx * G + g) * 2;
        const long long base = ((long long)by * p.H + ya) * W;
        const PT* __restrict__ fa = reinterpret_cast<const PT*>(p.frames) + base;
        const PT* __restrict__ fb = fa + W;
        if (p.mask) {
            const uint8_t* ma = p.mask + (long long)by * p.mask_stride + (long long)ya * W;
            const uint8_t* mb = ma + W;
            const float* ra = p.reference + (long long)ya * W;
            const float* rb = ra + W;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int x = t + TPF * m;
                v[m] = mk<float>(ma[x] ? ra[x] : (float)fa[x], mb[x] ? rb[x] : (float)fb[x]);
            }
        } else {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int x = t + TPF * m;
                v[m] = mk<float>((float)fa[x], (float)fb[x]);
            }
        }
    }
    FCD_HD static void load_rows(const Params& p, int bx, int by, int g, int t, cf* v) {
        if (p.frame_kind == 0) load_rows_t<float>(p, bx, by, g, t, v);
        else if (p.frame_kind == 1) load_rows_t<uint8_t>(p, bx, by, g, t, v);
        else load_rows_t<uint16_t>(p, bx, by, g, t, v);
    }

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        unsigned char* smem = smem_all + TW::TW_BYTES;
        const int g = tid / TPF, t = tid % TPF;
        cf* s = reinterpret_cast<cf*>(smem) + g * GL::STRIDE;
        const int W = L;
        if constexpr (PH == 0) {
            if (st.link.first) load_rows(p, bx, by, g, t, st.v);   // later tiles were prefetched in phase 5
            F::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            F::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 2) {
            F::stepC(st.v, t, s);
        } else if constexpr (PH == 3) {
            F::stepD(st.v, t, s, tw);
        } else if constexpr (PH == 4) {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) s[fft_nat<TPF>(t, m)] = st.v[m];
        } else {
            // the data registers are dead from here on: fetch the next tile's rows into them
            // while this phase and the loop barrier run
            if (st.link.has_next) load_rows(p, st.link.next_bx, st.link.next_by, g, t, st.v);
            // band extraction: lanes = G adjacent row pairs (one 16-byte store each) x consecutive band columns
            const cf* sb = reinterpret_cast<const cf*>(smem);
            const int gg = tid % G, c0 = tid / G;
            const int ya = (bx * G + gg) * 2;
            const cf* sg = sb + gg * GL::STRIDE;
            FCD_UNROLL
            for (int i = 0; i < 2; ++i) {
                cf* __restrict__ dst = p.w1 + ((long long)by * 2 + i) * p.ncp * p.H + ya;
                const int nc = p.nc[i], kc0 = p.kc0[i];
                for (int c = c0; c < nc; c += THREADS / G) {
                    const int kc = kc0 + c;
                    const int k = kc < 0 ? -kc : kc;
                    const cf x1 = sg[fft_pos(k)];
                    const cf x2 = conj(sg[fft_pos((W - k) & (W - 1))]);
                    cf a = x1 + x2;              // 2 * rowA spectrum at k
                    cf b = mul_mi(x1 - x2);      // 2 * rowB spectrum at k
                    if (kc < 0) { a = conj(a); b = conj(b); }
                    cf2 o; o.a = a; o.b = b;
                    *reinterpret_cast<cf2*>(dst + (long long)c * p.H) = o;
                }
            }
        }
    }
};

// =========================================================================================
// K2  column band-pass: forward FFT along y, keep the disk chord of this column, inverse
//     FFT along y.   w1[f][i][c'][y]  ->  w2[f][i][y][c']  (c' contiguous)
// =========================================================================================
struct ColBandParams {
    const cf* w1;
    cf* w2;
    const cf* tw;
    const int* chord_lo;   // [2][ncp] first kept shifted row (inclusive)
    const int* chord_hi;   // [2][ncp] last kept shifted row (inclusive); lo > hi: empty
    int ncp;
    int nc[2];
    float scale;
};

template <int L, int G>
struct ColBand : AllPhases {
    using FF = Fft<L, -1, float, ColPlan<L>>;
    using FI = Fft<L, +1, float, ColPlan<L>>;
    using GL = GroupLayout<L, G>;
    using Params = ColBandParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = true;   // the next tile's columns arrive by TMA while this tile finishes
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = ((G * L / 16) <= 256 ? 2 : 1);
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 10;
    using TW = SmemTwiddles<FF, THREADS>;
    static_assert(L + 1 <= GL::STRIDE, "a column (16-byte aligned) must fit in the exchange buffer");
    static constexpr int BAR_OFF = TW::TW_BYTES + G * GL::STRIDE * (int)sizeof(cf);
    static constexpr int SMEM_BYTES = BAR_OFF + 16;
    FCD_HD static mbar_t* bar_of(unsigned char* smem_all) { return reinterpret_cast<mbar_t*>(smem_all + BAR_OFF); }
    FCD_HD static cf* landing(unsigned char* smem_all, int g) {
        return reinterpret_cast<cf*>(smem_all + TW::TW_BYTES) + g * GL::STRIDE + ((g * GL::STRIDE) & 1);
    }
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) {
        TW::load(p.tw, tid, smem);
        if (tid == 0) mbar_init(bar_of(smem), 1);
    }
    struct State { cf v[16]; TileLink link; unsigned parity; };

    // one thread hands the tile's (up to G) contiguous w1 columns to the copy engine; the exchange buffers
    // are the landing zones (linear order)
    FCD_HD static void stage_cols(const Params& p, int bx, int by, unsigned char* smem_all) {
        const int i = by & 1;
        constexpr unsigned BYTES = L * (unsigned)sizeof(cf);
        int n = 0;
        for (int g = 0; g < G; ++g) n += (bx * G + g < p.nc[i]) ? 1 : 0;
        mbar_t* bar = bar_of(smem_all);
        mbar_expect_tx(bar, n * BYTES);
        for (int g = 0; g < G; ++g) {
            const int c = bx * G + g;
            if (c < p.nc[i]) bulk_copy_g2s(landing(smem_all, g), p.w1 + ((long long)by * p.ncp + c) * L, BYTES, bar);
        }
    }

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        unsigned char* smem = smem_all + TW::TW_BYTES;
        const int g = tid % G, t = tid / G;   // group fastest: a warp covers 32/G rows x G columns
        cf* s = reinterpret_cast<cf*>(smem) + g * GL::STRIDE;
        const int H = L;
        const int i = by & 1;
        const int c = bx * G + g;
        if constexpr (PH == 0) {
            if (st.link.first) {
                st.parity = 0;
#if defined(FCD_EMULATE)
                stage_cols(p, bx, by, smem_all);      // sequential emulation: every thread copies for itself (idempotent)
#else
                if (tid == 0) stage_cols(p, bx, by, smem_all);
#endif
            }
            mbar_wait(bar_of(smem_all), st.parity);
            st.parity ^= 1u;
            const cf* col = landing(smem_all, g);
            const bool valid = c < p.nc[i];
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) st.v[m] = valid ? col[t + TPF * m] : mk<float>(0.f, 0.f);
        } else if constexpr (PH == 1) {
            FF::stepA(st.v, t, s);
        } else if constexpr (PH == 2) {
            FF::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 3) {
            FF::stepC(st.v, t, s);
        } else if constexpr (PH == 4) {
            FF::stepD(st.v, t, s, tw);
            int lo = 1, hi = 0;
            if (c < p.nc[i]) { lo = p.chord_lo[i * p.ncp + c]; hi = p.chord_hi[i * p.ncp + c]; }
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int r = (t + TPF * m + H / 2) & (H - 1);  // shifted row of frequency index
                const bool keep = (r >= lo) && (r <= hi);
                st.v[m] = keep ? scale(st.v[m], p.scale) : mk<float>(0.f, 0.f);
            }
        } else if constexpr (PH == 5) {
            FI::stepA(st.v, t, s);
        } else if constexpr (PH == 6) {
            FI::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 7) {
            FI::stepC(st.v, t, s);
        } else if constexpr (PH == 8) {
            FI::template gather<FI::R3, FI::R1 * FI::R2>(st.v, t, s, tw);
        } else {
            // the exchange buffers are idle from here on: hand them to the copy engine for the next tile
            if (st.link.has_next && tid == 0) {
                fence_proxy_async();
                stage_cols(p, st.link.next_bx, st.link.next_by, smem_all);
            }
            FI::template butterflies<FI::R3>(st.v);
            if (c < p.nc[i]) {   // lanes: G adjacent columns x 32/G adjacent rows -> full sectors
                cf* o = p.w2 + ((long long)by * H + t) * p.ncp + c;      // row t; the owned rows are TPF rows apart
                const long long stride = (long long)TPF * p.ncp;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) { *o = st.v[m]; o += stride; }
            }
        }
    }
};

// =========================================================================================
// K3  row demodulation: for both carriers inverse row FFT of the band, multiply by ccsgn,
//     -atan2 -> wrapped phases; unwrap along the row (integer prefix sum of 2pi jumps,
//     anchored at x_ref); forward row FFT of z = phi0 + i*phi1.
//     w2[f][i][y][c'] -> w3[f][y][kc] (+ colphase[f][i][y] = wrapped phase at x_ref,
//     + optional phases[f][i][y][x])
// =========================================================================================
struct RowDemodParams {
    const cf* w2;
    const float* theta;  // [H][W][2]  angle(ccsgn) of the bound reference, the two carriers interleaved
    cf* w3;              // [F][W/4+1][H][4]  (column-blocked, see w3_index)
    float* colphase;     // [F][2][H]
    float* phases;       // [F][2][H][W] or null
    const cf* tw;
    int H, ncp;
    int nc[2];
    int kc0[2];
    int x_ref;
    int unwrap;
    int* frameflag;      // [F] or null: set to 1 for frames with |phi| > pi/2 somewhere (the only frames that can
                         // hold a 2*pi jump, let alone a residue: unwrap "auto" looks no further at the others)
};

struct int2s { int a, b; };

template <int L, int G, bool PRUNED = false>
struct RowDemod {
    using FF = Fft<L, -1, float, RowPlan<L>>;
    using FI = Fft<L, +1, float, RowPlan<L>>;
    using GL = GroupLayout<L, G, 2>;
    using Params = RowDemodParams;
    static constexpr bool BLOCKED_TILES = true;
    static constexpr bool PIPELINED = PRUNED;   // pruned path: next tile's band values arrive by cp.async
    static constexpr int SYNC_THREADS = (L / 16 >= 32 && G > 1 && G <= 15) ? L / 16 : 0;   // per-group named barriers
    static constexpr int MIN_BLOCKS = ((G * L / 16) <= 128 ? 4 : ((G * L / 16) <= 256 ? 2 : 1));
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 13;
    using TW = SmemTwiddles<FF, THREADS>;
    // per group: two exchange buffers (one per carrier; the second doubles as the jump-scan
    // array), chunk totals, chunk offsets, flag, and (pruned path) the staging slots of the
    // four band values each thread feeds into the first pass: band[q][t], q = carrier*2 + butterfly
    static constexpr int AUX_INTS = 4 * TPF + 4;
    static constexpr int BAND_ELEMS = PRUNED ? 4 * TPF : 0;
    static constexpr int GROUP_BYTES = GL::GROUP_STRIDE * (int)sizeof(cf) + AUX_INTS * (int)sizeof(int) +
                                       BAND_ELEMS * (int)sizeof(cf);
    static constexpr int SMEM_BYTES = TW::TW_BYTES + G * GROUP_BYTES;
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    struct State {
        cf v[16];   // carrier 0, later (phi0, phi1)
        cf w[16];   // carrier 1
        TileLink link;
    };

    // Pruned path: first-pass butterfly ii of carrier i has a single non-zero input, band column
    // c(t, ii, i) of the row -- independent of the tile.  Each thread stages its own four values
    // (thread-private slots: cp.async + wait, no barrier) one tile ahead, so the DRAM round trip
    // that used to open every tile (ncu r01_m2: 32 % of the kernel's samples, 65 % long
    // scoreboard) overlaps the previous tile's arithmetic.
    FCD_HD static int band_col(const Params& p, int i, int t, int ii) {
        constexpr int M1 = L / 8;
        return (t + TPF * ii - (p.kc0[i] & (L - 1))) & (M1 - 1);
    }
    FCD_HD static cf* band_slots(unsigned char* gbase) {
        return reinterpret_cast<cf*>(gbase + GL::GROUP_STRIDE * sizeof(cf) + AUX_INTS * sizeof(int));
    }
    FCD_HD static void stage_band(const Params& p, int f, int y, int t, cf* band) {
        FCD_UNROLL
        for (int q = 0; q < 4; ++q) {
            const int i = q >> 1;
            const int c = band_col(p, i, t, q & 1);
            if (c < p.nc[i])
                async_copy8(band + q * TPF + t, p.w2 + (((long long)f * 2 + i) * p.H + y) * p.ncp + c);
        }
    }

    // the skip decision must be uniform over the barrier domain: one flag per group with
    // per-group named barriers, one flag per block with block-wide barriers
    FCD_HD static int* group_flag(unsigned char* smem_all, int tid) {
        const int fg = (SYNC_THREADS != 0) ? tid / TPF : 0;
        unsigned char* gbase = smem_all + TW::TW_BYTES + (size_t)fg * GROUP_BYTES;
        return reinterpret_cast<int*>(gbase + GL::GROUP_STRIDE * sizeof(cf)) + 4 * TPF;
    }
    // Phases 4..8 are the row unwrap.  A 2*pi jump between neighbours needs |phi| > pi/2
    // somewhere in the row, which phase 3 detects thread-locally; rows that cannot wrap skip
    // the whole block (group-uniform: the flag is per row).
    template <int PH>
    FCD_HD static bool enabled(const Params& p, const unsigned char* smem_all, int tid) {
        if constexpr (PH >= 4 && PH <= 8)
            return p.unwrap && *group_flag(const_cast<unsigned char*>(smem_all), tid) != 0;
        else
            return true;
    }

    FCD_HD static void load_band(const Params& p, int f, int i, int y, int t, cf* v) {
        const int W = L;
        const cf* __restrict__ row = p.w2 + (((long long)f * 2 + i) * p.H + y) * p.ncp;
        const int c0 = (t - p.kc0[i]) & (W - 1);     // band column of spectrum position t
        const int nc = p.nc[i];
        FCD_UNROLL
        for (int m = 0; m < 16; ++m) {
            const int c = (c0 + TPF * m) & (W - 1);
            v[m] = (c < nc) ? row[c] : mk<float>(0.f, 0.f);
        }
    }
    // -angle(g * ccsgn) = -wrap(angle(g) + angle(ccsgn))           (fcd.py:118)
    // both carriers' reference angles of pixels m0 .. m0 + 7 of this thread: one 8-byte load per pixel
    FCD_HD static void load_theta8(const Params& p, int y, int t, int m0, float* c0, float* c1) {
        const cf* __restrict__ th = reinterpret_cast<const cf*>(p.theta) + (long long)y * L;
        FCD_UNROLL
        for (int m = m0; m < m0 + 8; ++m) {
            const cf a = th[t + TPF * m];
            c0[m] = a.x; c1[m] = a.y;
        }
    }
    FCD_HD static bool demod(const float* c, const cf* v, float* ph) {
        float big = 0.f;
        FCD_UNROLL
        for (int m = 0; m < 16; ++m) {
            const float a = fast_atan2f(v[m].y, v[m].x) + c[m];
            ph[m] = kTwoPiF * rintf(a * kInvTwoPiF) - a;
            big = fmaxf(big, fabsf(ph[m]));
        }
        return big > 1.57079632679489661923f;
    }

    // Both carriers of a row are transformed together (shared twiddle loads, half the barriers).
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        unsigned char* smem = smem_all + TW::TW_BYTES;
        const int g = tid / TPF, t = tid % TPF;
        unsigned char* gbase = smem + (size_t)g * GROUP_BYTES;
        cf* s0 = reinterpret_cast<cf*>(gbase);
        cf* s1 = s0 + GL::STRIDE;
        int2s* sj = reinterpret_cast<int2s*>(s1);               // jump scan array lives in buffer 1
        int* aux = reinterpret_cast<int*>(gbase + GL::GROUP_STRIDE * sizeof(cf));
        int2s* part = reinterpret_cast<int2s*>(aux);            // [TPF]
        int2s* off = reinterpret_cast<int2s*>(aux + 2 * TPF);   // [TPF]
        int* flag = group_flag(smem_all, tid);
        const int W = L;
        const int y = by * G + g;   // tiles are ordered frame-fastest so that consecutive tiles of
        const int f = bx;           // a block reuse the same theta rows out of L2
        if constexpr (PH == 0) {
            if (t == 0) *flag = 0;
            if constexpr (PRUNED) {
                // band no wider than W/8: every radix-8 butterfly of the first pass has at most
                // one non-zero input -> one staged value and a few rotations per butterfly
                constexpr int M1 = L / 8;
                cf* band = band_slots(gbase);
                if (st.link.first) stage_band(p, f, y, t, band);   // later tiles were staged in phase 1
                async_wait_all();
                cf nx[4];
                FCD_UNROLL
                for (int q = 0; q < 4; ++q)
                    nx[q] = (band_col(p, q >> 1, t, q & 1) < p.nc[q >> 1]) ? band[q * TPF + t] : mk<float>(0.f, 0.f);
                FCD_UNROLL
                for (int i = 0; i < 2; ++i) {
                    const int p0 = p.kc0[i] & (W - 1);
                    cf* sb = i == 0 ? s0 : s1;
                    FCD_UNROLL
                    for (int ii = 0; ii < 2; ++ii) {
                        const int c = (t + TPF * ii - p0) & (M1 - 1);
                        const int pp = (p0 + c) & (W - 1);
                        FI::stepA_single(nx[i * 2 + ii], pp / M1, ii, t, sb, tw);
                    }
                }
            } else {
                load_band(p, f, 0, y, t, st.v);
                load_band(p, f, 1, y, t, st.w);
                FI::stepA(st.v, t, s0);
                FI::stepA(st.w, t, s1);
            }
        } else if constexpr (PH == 1) {
            if constexpr (PRUNED) {   // own slots were consumed before the barrier: stage the next tile
                if (st.link.has_next)
                    stage_band(p, st.link.next_bx, st.link.next_by * G + g, t, band_slots(gbase));
            }
            FI::stepB2(st.v, st.w, t, s0, s1, tw);
        } else if constexpr (PH == 2) {
            FI::stepC(st.v, t, s0);
            FI::stepC(st.w, t, s1);
        } else if constexpr (PH == 3) {
            // the reference angles of the first eight pixels are requested before the last butterflies, those of the
            // other eight before the first arctangents: each batch of loads has a stretch of arithmetic to hide behind
            float c0[16], c1[16];
            load_theta8(p, y, t, 0, c0, c1);
            FI::stepD2(st.v, st.w, t, s0, s1, tw);
            load_theta8(p, y, t, 8, c0, c1);
#if defined(FCD_PACKED_F32)
            float big = 0.f;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                st.v[m] = demod_pair(st.v[m], st.w[m], c0[m], c1[m]);
                big = fmaxf(big, fmaxf(fabsf(st.v[m].x), fabsf(st.v[m].y)));
            }
            if (big > 1.57079632679489661923f) {               // this row may contain 2*pi jumps
                *flag = 1;
                if (p.frameflag) p.frameflag[f] = 1;
            }
#else
            float ph0[16], ph1[16];
            const bool big0 = demod(c0, st.v, ph0);
            const bool big1 = demod(c1, st.w, ph1);
            if (big0 || big1) {               // this row may contain 2*pi jumps
                *flag = 1;
                if (p.frameflag) p.frameflag[f] = 1;
            }
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) st.v[m] = mk<float>(ph0[m], ph1[m]);
#endif
            // wrapped phase at the anchor column links the rows (RowLink)
            if (t == (p.x_ref % TPF)) {
                const int mr = p.x_ref / TPF;
                cf a = mk<float>(0.f, 0.f);
                FCD_UNROLL
                for (int mm = 0; mm < 16; ++mm)
                    if (mm == mr) a = st.v[mm];
                p.colphase[((long long)f * 2 + 0) * p.H + y] = a.x;
                p.colphase[((long long)f * 2 + 1) * p.H + y] = a.y;
            }
        } else if constexpr (PH == 4) {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) s0[fft_nat<TPF>(t, m)] = st.v[m];
        } else if constexpr (PH == 5) {
            // 2*pi jumps to the left neighbour -> buffer 1 (as integers)
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int x = t + TPF * m;
                int2s j; j.a = 0; j.b = 0;
                if (x > 0) {
                    const cf prev = s0[fft_pos(x - 1)];
                    j.a = (int)rintf((st.v[m].x - prev.x) * kInvTwoPiF);
                    j.b = (int)rintf((st.v[m].y - prev.y) * kInvTwoPiF);
                }
                sj[fft_nat<TPF>(t, m)] = j;
            }
        } else if constexpr (PH == 6) {
            int a = 0, b = 0;              // inclusive scan of this thread's contiguous chunk
            FCD_UNROLL
            for (int q = 0; q < 16; ++q) {
                int2s j = sj[fft_pos(16 * t + q)];
                a += j.a; b += j.b;
                j.a = a; j.b = b;
                sj[fft_pos(16 * t + q)] = j;
            }
            int2s tot; tot.a = a; tot.b = b;
            part[t] = tot;
        } else if constexpr (PH == 7) {
            int a = 0, b = 0;
            for (int q = 0; q < t; ++q) { a += part[q].a; b += part[q].b; }
            int2s o; o.a = a; o.b = b;
            off[t] = o;
        } else if constexpr (PH == 8) {
            const int2s cr = sj[fft_pos(p.x_ref)];
            const int2s orf = off[p.x_ref >> 4];
            const int ra = cr.a + orf.a, rb = cr.b + orf.b;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int x = t + TPF * m;
                const int2s c = sj[fft_nat<TPF>(t, m)];
                const int2s o = off[x >> 4];
                st.v[m].x -= kTwoPiF * (float)(c.a + o.a - ra);
                st.v[m].y -= kTwoPiF * (float)(c.b + o.b - rb);
            }
        } else if constexpr (PH == 9) {
            if (p.phases) {
                float* o0 = p.phases + (((long long)f * 2 + 0) * p.H + y) * W;
                float* o1 = p.phases + (((long long)f * 2 + 1) * p.H + y) * W;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    o0[t + TPF * m] = st.v[m].x;
                    o1[t + TPF * m] = st.v[m].y;
                }
            }
            FF::stepA(st.v, t, s0);
        } else if constexpr (PH == 10) {
            FF::stepB(st.v, t, s0, tw);
        } else if constexpr (PH == 11) {
            FF::stepC(st.v, t, s0);
        } else {
            FF::stepD(st.v, t, s0, tw);
            // column-blocked store (see w3_index): two bases, constant strides
            const int H = p.H;
            cf* base = p.w3 + ((long long)f * w3_blocks(W) * H + y) * 4;
            const long long step = (long long)TPF * H;                     // TPF/4 blocks of H*4 elements
            const long long lo = (long long)(t >> 2) * H * 4 + (t & 3);     // slot = t + TPF*m, m < 8
            const int s8 = W / 2 + t + (t > 0 ? 3 : 0);                     // kc = W/2 + t (m = 8)
            const int sh = W / 2 + t + 3;                                   // kc = W/2 + t + TPF*(m-8), m > 8
            const long long o8 = (long long)(s8 >> 2) * H * 4 + (s8 & 3);
            const long long hi = (long long)(sh >> 2) * H * 4 + (sh & 3);
            FCD_UNROLL
            for (int m = 0; m < 8; ++m) base[lo + m * step] = st.v[m];
            base[o8] = st.v[8];
            FCD_UNROLL
            for (int m = 9; m < 16; ++m) base[hi + (m - 8) * step] = st.v[m];
        }
    }
};



// =========================================================================================
// K3b row linking: integer prefix sum along y of the 2pi jumps of the anchor column.
//     rowoff[f][i][y] = -2pi * (M(y) - M(y_ref));  also (optionally) finishes `phases`.
// =========================================================================================
struct RowLinkParams {
    const float* colphase;  // [F][2][H]
    float* rowoff;          // [F][2][H]
    int H, y_ref;
    int unwrap;
};

struct RowLink : NoPrologue {
    using Params = RowLinkParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 4;
    static constexpr int MAXH = 4096;
    static constexpr int SMEM_BYTES = (MAXH + 2 * THREADS + 4) * (int)sizeof(int);
    struct State { int dummy; };

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State&) {
        int* c = reinterpret_cast<int*>(smem);
        int* part = c + MAXH;
        int* off = part + THREADS;
        int* mref = off + THREADS;
        const int H = p.H;
        const int chunk = (H + THREADS - 1) / THREADS;
        const int y0 = tid * chunk;
        const float* cp = p.colphase + ((long long)by * 2 + bx) * H;
        float* out = p.rowoff + ((long long)by * 2 + bx) * H;
        if constexpr (PH == 0) {
            int acc = 0;
            for (int q = 0; q < chunk; ++q) {
                const int y = y0 + q;
                if (y < H) {
                    if (y > 0 && p.unwrap) acc += (int)rintf((cp[y] - cp[y - 1]) * kInvTwoPiF);
                    c[y] = acc;
                }
            }
            part[tid] = acc;
        } else if constexpr (PH == 1) {
            int a = 0;
            for (int q = 0; q < tid; ++q) a += part[q];
            off[tid] = a;
        } else if constexpr (PH == 2) {
            if (tid == 0) *mref = c[p.y_ref] + off[p.y_ref / chunk];
        } else {
            for (int q = 0; q < chunk; ++q) {
                const int y = y0 + q;
                if (y < H) out[y] = -kTwoPiF * (float)(c[y] + off[tid] - *mref);
            }
        }
    }
};

// adds the row offsets to a materialised phases array (only when phases are requested)
struct PhaseFixParams {
    float* phases;         // [F][2][H][W]
    const float* rowoff;   // [F][2][H]
    int H, W;
};
struct PhaseFix : NoPrologue {
    using Params = PhaseFixParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        // bx: row, by: f*2+i
        const float o = p.rowoff[(long long)by * p.H + bx];
        if (o != 0.f) {
            float* row = p.phases + ((long long)by * p.H + bx) * p.W;
            for (int x = tid; x < p.W; x += THREADS) row[x] += o;
        }
    }
};

// Residue guard: number of 2x2 loops of a phase map whose wrapped differences do
// not sum to zero.  Parity of the unwrapped phases with the reference's unwrapper is defined
// only when this is zero.
struct ResidueParams {
    const float* phases;   // [M][H][W]
    int* counts;           // [M], zero-initialised
    int H, W;
};
struct ResidueCount : NoPrologue {
    using Params = ResidueParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    FCD_HD static float wrapd(float d) { return d - kTwoPiF * rintf(d * kInvTwoPiF); }
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        // bx: row y (0..H-2), by: map
        const float* r0 = p.phases + ((long long)by * p.H + bx) * p.W;
        const float* r1 = r0 + p.W;
        int n = 0;
        for (int x = tid; x + 1 < p.W; x += THREADS) {
            const float s = wrapd(r0[x + 1] - r0[x]) + wrapd(r1[x + 1] - r0[x + 1]) + wrapd(r1[x] - r1[x + 1]) +
                            wrapd(r0[x] - r1[x]);
            if (rintf(s * kInvTwoPiF) != 0.f) ++n;
        }
        if (n) {
#if defined(__CUDA_ARCH__)
            atomicAdd(p.counts + by, n);
#else
            p.counts[by] += n;
#endif
        }
    }
};

// =========================================================================================
// K4  column integration: forward FFT along y of columns kc and W-kc of z_row, split into
//     the spectra of phi0 and phi1, apply the folded 2x2-solve / (-1/height) / (-i k / k^2)
//     coefficients (Hermitian part, = np.real(ifft2(.))), inverse FFT along y.
//     w3[f][y][kc] -> w4[f][y][kc], kc in [0, W/2]
// =========================================================================================
struct ColIntegrateParams {
    const cf* w3;
    const float* rowoff;   // [F][2][H]
    cf* w4;                // [F][H][w4p]
    const cf* tw;
    const float* kx;       // [W]  column wavenumbers (plain, used for k^2)
    const float* kxq;      // [W]  with index W/2+1 zeroed (fourier.py:89)
    float dky;             // row wavenumber step: ky[kr] = signed(kr) * dky; index H/2+1 zeroed (fourier.py:92)
    int W, w4p;
    float f0r, f0c, f1r, f1c;   // carrier wavevectors [k_row, k_col] (fcd.py:134-137)
    float scale;                // 1 / (2 * height * det * H * W)
    int unwrap;
};

template <int L, int G>
struct ColIntegrate : AllPhases {
    using FF = Fft<L, -1, float, ColPlan<L>>;
    using FI = Fft<L, +1, float, ColPlan<L>>;
    using GL = GroupLayout<L, G, 2>;
    using Params = ColIntegrateParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = true;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 9;
    using TW = SmemTwiddles<FF, THREADS>;
    static constexpr int SMEM_BYTES = TW::TW_BYTES + G * GL::GROUP_STRIDE * (int)sizeof(cf);
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    struct State { cf v[16]; cf va[16]; TileLink link; };

    FCD_HD static const cf* col_ptr(const Params& p, int f, int kc, int t) { return p.w3 + w3_index(f, kc, t, L, p.W); }

    // Column pair (kc, W-kc) of z_row = Phi0_row + i Phi1_row.  The two real fields' row spectra
    // are separated pointwise (Hermitian symmetry in kc) *before* the column transforms, so
    // the transforms of Phi0 and Phi1 are combined locally afterwards.  While the inverse
    // transform runs, the next tile's columns are prefetched: column kc into the idle
    // registers, column W-kc with cp.async into the idle exchange buffer.
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        unsigned char* smem = smem_all + TW::TW_BYTES;
        const int g = tid % G, t = tid / G;   // group fastest: a warp covers 32/G rows x G columns
        cf* s0 = reinterpret_cast<cf*>(smem) + g * GL::GROUP_STRIDE;
        cf* s1 = s0 + GL::STRIDE;
        const int H = L;
        const int f = by;
        const int kc = bx * G + g;
        const bool valid = kc <= p.W / 2;
        const int kcm = (p.W - kc) & (p.W - 1);
        if constexpr (PH == 0) {
            if (st.link.first) {
                if (valid) {
                    const cf* a = col_ptr(p, f, kc, t);
                    const cf* b = col_ptr(p, f, kcm, t);
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) st.va[m] = a[TPF * m * 4];
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) st.v[m] = b[TPF * m * 4];
                }
            } else {
                async_wait_all();   // each thread reads back only what it copied itself
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) st.v[m] = s1[fft_nat<TPF>(t, m)];
            }
        } else if constexpr (PH == 1) {
            if (valid) {
                if (kc == 0 && p.unwrap) {   // row offsets of the unwrap enter only the kc = 0 column
                    const float* r0 = p.rowoff + ((long long)f * 2 + 0) * H;
                    const float* r1 = p.rowoff + ((long long)f * 2 + 1) * H;
                    const float w = (float)p.W;
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) {
                        const cf add = mk<float>(w * r0[t + TPF * m], w * r1[t + TPF * m]);
                        st.va[m] = st.va[m] + add;
                        st.v[m] = st.v[m] + add;      // kcm == kc == 0
                    }
                }
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    const cf z = st.va[m], zm = conj(st.v[m]);
                    st.va[m] = z + zm;              // 2 * Phi0_row(y, kc)
                    st.v[m] = z - zm;               // 2i * Phi1_row(y, kc): the factor -i commutes exactly with the
                                                    // transform (swap / negate) and is applied in phase 4
                }
            } else {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) { st.va[m] = mk<float>(0.f, 0.f); st.v[m] = mk<float>(0.f, 0.f); }
            }
            FF::stepA(st.va, t, s0);
            FF::stepA(st.v, t, s1);
        } else if constexpr (PH == 2) {
            FF::stepB2(st.va, st.v, t, s0, s1, tw);
        } else if constexpr (PH == 3) {
            FF::stepC(st.va, t, s0);
            FF::stepC(st.v, t, s1);
        } else if constexpr (PH == 4) {
            FF::stepD2(st.va, st.v, t, s0, s1, tw);
            if (valid) {
                // Folded coefficients of Phi0, Phi1 in hhat / i, Hermitian part:
                //   aH = ((kxa - kxb) f1r - (kya - kyb) f1c) / (2 k^2) * scale,   bH = ((kya - kyb) f0c - (kxa - kxb) f0r) / (2 k^2) * scale
                // with kxa / kxb = quirk-zeroed kx[kc] / kx[-kc] (fourier.py:89) and kya / kyb = quirk-zeroed ky[kr] / ky[-kr]
                // (fourier.py:92).  kya - kyb = wy * ky[kr] with wy = 2 except on three rows: kr = H/2 (ky[-kr] = ky[kr]: 0),
                // kr = H/2 + 1 (kya zeroed: 1), kr = H/2 - 1 (kyb zeroed: 1).  Rows are kr = t + TPF*m with t < TPF, so the
                // exceptions are compile-time impossible except for m = 7, 8.
                const float kxv = p.kx[kc];
                const float hs = 0.5f * p.scale;
                const float dkx = (p.kxq[kc] - p.kxq[kcm]) * hs;
                const cf cx0 = mk<float>(dkx * p.f1r, -dkx * p.f0r);       // kx part of (aH, bH) * k^2
                const cf cy = mk<float>(-p.f1c * hs, p.f0c * hs);           // (aH, bH) * k^2 per unit of (kya - kyb)
                const float kx2 = kxv * kxv;
                const float tf = (float)t;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    const float kyv = (tf + (float)(TPF * m - (m >= 8 ? H : 0))) * p.dky;   // ky[kr], exact integer * dky
                    float wy = 2.f;
                    if (m == 8) wy = (t == 0) ? 0.f : ((t == 1) ? 1.f : 2.f);
                    if (m == 7) wy = (t == TPF - 1) ? 1.f : 2.f;
                    float k2 = kyv * kyv + kx2;
                    if (m == 0) k2 = (t == 0 && kc == 0) ? 1.f : k2;                        // k2[0,0] = 1 (fourier.py:130)
                    const cf ab = scale(axpy(wy * kyv, cy, cx0), fast_rcp(k2));            // (aH, bH)
                    // i (aH 2Phi0 + bH 2Phi1) with 2Phi1 = -i D, D = st.v:   i aH 2Phi0 + bH D.  The product with aH is
                    // rounded first and bH D joins it in the FMA, lane for lane what mul_pi(lin2(aH, 2Phi0, bH, -i D)) does
                    st.v[m] = axpy(ab.y, st.v[m], scale(mul_pi(st.va[m]), ab.x));
                }
            } else {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) st.v[m] = mk<float>(0.f, 0.f);
            }
        } else if constexpr (PH == 5) {
            FI::stepA(st.v, t, s0);
            if (st.link.has_next) {       // prefetch the next tile (both buffers' readers are past the barrier)
                const int nkc = st.link.next_bx * G + g, nf = st.link.next_by;
                if (nkc <= p.W / 2) {
                    const int nkcm = (p.W - nkc) & (p.W - 1);
                    const cf* a = col_ptr(p, nf, nkc, t);
                    const cf* b = col_ptr(p, nf, nkcm, t);
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) st.va[m] = a[TPF * m * 4];
                    FCD_UNROLL
                    for (int m = 0; m < 16; ++m) async_copy8(&s1[fft_nat<TPF>(t, m)], &b[TPF * m * 4]);
                }
            }
        } else if constexpr (PH == 6) {
            FI::stepB(st.v, t, s0, tw);
        } else if constexpr (PH == 7) {
            FI::stepC(st.v, t, s0);
        } else {
            FI::stepD(st.v, t, s0, tw);
            if (valid) {
                cf* o = p.w4 + ((long long)f * H + t) * p.w4p + kc;     // row t; the owned rows are TPF rows apart
                const long long stride = (long long)TPF * p.w4p;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) { *o = st.v[m]; o += stride; }
            }
        }
    }
};



// =========================================================================================
// K5  row inverse (c2r, two rows per transform):  w4[f][y][0..W/2] -> height[f][y][x]
// =========================================================================================
struct RowInvParams {
    const cf* w4;
    float* height;          // [F][H][W]
    const uint8_t* mask;    // optional: height *= ~mask  (analyze.py:254-255)
    long long mask_stride;
    const cf* tw;
    int H, w4p;
};

template <int L, int G>
struct RowInv : AllPhases {
    using FI = Fft<L, +1, float, RegenRowPlan<L>>;    // same radix order and table as RowPlan<L> (fft_core.cuh, TwRegen)
    using GL = GroupLayout<L, G>;
    using Params = RowInvParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = true;   // the next tile's two spectrum rows arrive by TMA while this tile finishes
    static constexpr int SYNC_THREADS = (L / 16 >= 32 && G > 1 && G <= 15) ? L / 16 : 0;   // per-group named barriers
    static constexpr int MIN_BLOCKS = ((G * L / 16) <= 128 ? 6 : ((G * L / 16) <= 256 ? 3 : 1));
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 6;
    using TW = SmemTwiddles<FI, THREADS>;
    // The exchange buffer doubles as the TMA landing zone: row a at element 0, row b at ROW_ELEMS
    // (W/2 + 2 elements each, a multiple of 16 bytes; w4 rows are W/2 + 4 elements apart).
    static constexpr int ROW_ELEMS = L / 2 + 2;
    static_assert(2 * ROW_ELEMS + 1 <= GL::STRIDE, "two spectrum rows (16-byte aligned) must fit in the exchange buffer");
    // bulk copies need 16-byte aligned shared addresses; a group's buffer may start on an odd element
    FCD_HD static cf* landing(cf* s, int g) { return s + ((g * GL::STRIDE) & 1); }
    static constexpr int BAR_OFF = TW::TW_BYTES + G * GL::STRIDE * (int)sizeof(cf);
    static constexpr int SMEM_BYTES = BAR_OFF + 16 * G;
    FCD_HD static mbar_t* bar_of(unsigned char* smem_all, int g) { return reinterpret_cast<mbar_t*>(smem_all + BAR_OFF + 16 * g); }
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) {
        TW::load(p.tw, tid, smem);
        if (tid < G) mbar_init(bar_of(smem, tid), 1);
    }
    struct State { cf v[16]; TileLink link; unsigned parity; };

    FCD_HD static void stage_rows(const Params& p, int bx, int by, int g, cf* s, mbar_t* bar) {
        const int ya = (bx * G + g) * 2;
        const cf* ra = p.w4 + ((long long)by * p.H + ya) * p.w4p;
        constexpr unsigned BYTES = ROW_ELEMS * (unsigned)sizeof(cf);
        cf* land = landing(s, g);
        mbar_expect_tx(bar, 2 * BYTES);
        bulk_copy_g2s(land, ra, BYTES, bar);
        bulk_copy_g2s(land + ROW_ELEMS, ra + p.w4p, BYTES, bar);
    }

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        unsigned char* smem = smem_all + TW::TW_BYTES;
        const int g = tid / TPF, t = tid % TPF;
        cf* s = reinterpret_cast<cf*>(smem) + g * GL::STRIDE;
        mbar_t* bar = bar_of(smem_all, g);
        const int W = L;
        const int ya = (bx * G + g) * 2;
        if constexpr (PH == 0) {
            if (st.link.first) {
                st.parity = 0;
#if defined(FCD_EMULATE)
                stage_rows(p, bx, by, g, s, bar);      // sequential emulation: every thread copies for itself (idempotent)
#else
                if (t == 0) stage_rows(p, bx, by, g, s, bar);
#endif
            }
            mbar_wait(bar, st.parity);
            st.parity ^= 1u;
            const cf* ra = landing(s, g);
            const cf* rb = ra + ROW_ELEMS;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const int pidx = t + TPF * m;
                const bool lower = pidx <= W / 2;
                const int k = lower ? pidx : W - pidx;
                cf a = ra[k], b = rb[k];
                if (k == 0 || k == W / 2) { a.y = 0.f; b.y = 0.f; }
                if (!lower) { a = conj(a); b = conj(b); }
                st.v[m] = a + mul_pi(b);
            }
        } else if constexpr (PH == 1) {
            FI::stepA(st.v, t, s);
        } else if constexpr (PH == 2) {
            FI::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 3) {
            FI::stepC(st.v, t, s);
        } else if constexpr (PH == 4) {
            FI::template gather<FI::R3, FI::R1 * FI::R2>(st.v, t, s, tw);
        } else {
            // every thread of the group has taken its last values out of the exchange buffer: hand it to
            // the copy engine for the next tile, then finish this one out of registers
            if (st.link.has_next && t == 0) {
                fence_proxy_async();
                stage_rows(p, st.link.next_bx, st.link.next_by, g, s, bar);
            }
            float* oa = p.height + ((long long)by * p.H + ya) * W;
            float* ob = oa + W;
            if (p.mask) {
                // the mask bytes are requested before the last butterflies so that their latency hides behind the
                // arithmetic; they are folded into two bit sets as they arrive
                const uint8_t* ma = p.mask + (long long)by * p.mask_stride + (long long)ya * W;
                const uint8_t* mb = ma + W;
                unsigned keep_a = 0, keep_b = 0;
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    keep_a |= (ma[t + TPF * m] ? 0u : 1u) << m;
                    keep_b |= (mb[t + TPF * m] ? 0u : 1u) << m;
                }
                FI::template butterflies<FI::R3>(st.v);
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    const int x = t + TPF * m;
                    oa[x] = ((keep_a >> m) & 1u) ? st.v[m].x : 0.f;
                    ob[x] = ((keep_b >> m) & 1u) ? st.v[m].y : 0.f;
                }
                return;
            }
            FI::template butterflies<FI::R3>(st.v);
            {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) {
                    oa[t + TPF * m] = st.v[m].x;
                    ob[t + TPF * m] = st.v[m].y;
                }
            }
        }
    }
};

}  // namespace fcd
