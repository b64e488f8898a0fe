// fft_core.cuh -- register/shared-memory Stockham FFT building blocks for sm_100a.
//
// Every transform in the FCD path is a complex FFT of one image row or column
// (length L = 64..4096, power of two).  L/16 threads cooperate on one transform; each
// thread owns the 16 elements  x[t + (L/16)*m], m = 0..15  both BEFORE and AFTER the
// transform ("natural strided ownership").  Consecutive threads therefore touch
// consecutive addresses, so the first pass can be fed straight from coalesced global
// loads and the last pass can feed coalesced global stores; only the two inter-pass
// exchanges go through shared memory (one padded buffer of L + L/16 elements).
//
// The transform is a decimation-in-time Stockham autosort in up to three passes of radix
// R1*R2*R3 = L (each radix in {2,4,8,16}; R2 = 1 means two passes).  A pass of radix R < 16
// runs 16/R butterflies per thread.  Pass structure (P = product of the previous radices):
//     butterfly i in [0, L/R):  k = i mod P
//        u[a] = x[i + a*L/R] * W_{P*R}^{k*a}          a = 0..R-1      (read, conflict free)
//        u    = DFT_R(u)
//        y[(i-k)*R + k + a*P] = u[a]                                   (write, padded)
// The code is __host__ __device__ so that tests/emul can execute exactly the same
// arithmetic and index maps on the CPU (this container has no GPU).
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>
#include <vector>

#if defined(__CUDACC__)
#define FCD_HD __host__ __device__ __forceinline__
#define FCD_UNROLL _Pragma("unroll")
#else
#define FCD_HD inline
#define FCD_UNROLL
#endif

namespace fcd {

template <class T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};
using cf = cx<float>;
using cd = cx<double>;

template <class T> FCD_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <class T> FCD_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> FCD_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> FCD_HD cx<T> operator*(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class T> FCD_HD cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
template <class T> FCD_HD cx<T> scale(cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// multiply by +i / -i
template <class T> FCD_HD cx<T> mul_pi(cx<T> a) { return mk<T>(-a.y, a.x); }
template <class T> FCD_HD cx<T> mul_mi(cx<T> a) { return mk<T>(a.y, -a.x); }
// multiply by exp(DIR * i * angle) given (c, s) = (cos, sin) of the positive angle
template <int DIR, class T> FCD_HD cx<T> rot(cx<T> a, T c, T s) {
    return DIR < 0 ? mk<T>(a.x * c + a.y * s, a.y * c - a.x * s) : mk<T>(a.x * c - a.y * s, a.y * c + a.x * s);
}
template <int DIR, class T> FCD_HD cx<T> mul_dir_i(cx<T> a) { return DIR < 0 ? mul_mi(a) : mul_pi(a); }
// a * p + b * q with real a, b;   s * p + q with real s
template <class T> FCD_HD cx<T> lin2(T a, cx<T> p, T b, cx<T> q) { return mk<T>(a * p.x + b * q.x, a * p.y + b * q.y); }
template <class T> FCD_HD cx<T> axpy(T s, cx<T> p, cx<T> q) { return mk<T>(s * p.x + q.x, s * p.y + q.y); }

// ---------------------------------------------------------------------------------------
// sm_100a packed float32 arithmetic.  A complex64 value is an aligned register pair, and
// FADD2 / FMUL2 / FFMA2 work on such pairs with free operand modifiers (swap halves, negate
// one half, broadcast a scalar register / immediate), so a complex add is ONE instruction, a
// complex multiply TWO (FMUL2 + FFMA2, the same roundings as the scalar FMUL + FFMA pair the
// compiler contracts to) and a multiplication by +-i folds into the consuming add.  These
// non-template overloads win over the generic templates for cx<float> in device code; the
// CPU emulation (tests/emul) keeps the scalar templates.  -DFCD_NO_PACKED restores scalar
// code on the device for A/B measurements.
// ---------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(FCD_NO_PACKED)
#define FCD_PACKED_F32 1
__device__ __forceinline__ float2 f2_of(cf a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cf cf_of(float2 a) { cf r; r.x = a.x; r.y = a.y; return r; }
__device__ __forceinline__ cf operator+(cf a, cf b) { return cf_of(__fadd2_rn(f2_of(a), f2_of(b))); }
__device__ __forceinline__ cf operator-(cf a, cf b) { return cf_of(__fadd2_rn(f2_of(a), make_float2(-b.x, -b.y))); }
__device__ __forceinline__ cf operator*(cf a, cf b) {
    // (a.x b.x - a.y b.y, a.y b.x + a.x b.y) = a * (b.x, b.x) + (a.y, a.x) * (-b.y, b.y)
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y));
    return cf_of(__ffma2_rn(f2_of(a), make_float2(b.x, b.x), t));
}
__device__ __forceinline__ cf scale(cf a, float s) { return cf_of(__fmul2_rn(f2_of(a), make_float2(s, s))); }
__device__ __forceinline__ cf lin2(float a, cf p, float b, cf q) {
    return cf_of(__ffma2_rn(f2_of(q), make_float2(b, b), __fmul2_rn(f2_of(p), make_float2(a, a))));
}
__device__ __forceinline__ cf axpy(float s, cf p, cf q) { return cf_of(__ffma2_rn(f2_of(p), make_float2(s, s), f2_of(q))); }
template <int DIR> __device__ __forceinline__ cf rot(cf a, float c, float s) {
    // a * (c + DIR * i * s)
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), DIR < 0 ? make_float2(s, -s) : make_float2(-s, s));
    return cf_of(__ffma2_rn(f2_of(a), make_float2(c, c), t));
}
#endif

// ---------------------------------------------------------------------------------------
// in-register DFTs on strided slots v[0], v[S], ..., natural order in and out.
// DIR = -1: forward (exp(-2 pi i nk/R)); DIR = +1: inverse (unnormalised).
// ---------------------------------------------------------------------------------------
template <int DIR, class T> FCD_HD void dft2(cx<T>& a, cx<T>& b) {
    cx<T> t = a - b;
    a = a + b;
    b = t;
}

template <int DIR, class T> FCD_HD void dft4(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    cx<T> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_dir_i<DIR>(a1 - a3);
    a0 = t0 + t2;
    a2 = t0 - t2;
    a1 = t1 + t3;
    a3 = t1 - t3;
}

template <int R, int S, int DIR, class T> struct Dft;

template <int S, int DIR, class T> struct Dft<2, S, DIR, T> {
    FCD_HD static void run(cx<T>* v) { dft2<DIR>(v[0], v[S]); }
};
template <int S, int DIR, class T> struct Dft<4, S, DIR, T> {
    FCD_HD static void run(cx<T>* v) { dft4<DIR>(v[0], v[S], v[2 * S], v[3 * S]); }
};
template <int S, int DIR, class T> struct Dft<8, S, DIR, T> {
    // n = 4*n1 + n2, k = k1 + 2*k2
    FCD_HD static void run(cx<T>* v) {
        const T h = T(0.70710678118654752440);
        cx<T> y0[4], y1[4];
        FCD_UNROLL
        for (int n2 = 0; n2 < 4; ++n2) {
            y0[n2] = v[n2 * S] + v[(n2 + 4) * S];
            y1[n2] = v[n2 * S] - v[(n2 + 4) * S];
        }
        y1[1] = rot<DIR>(y1[1], h, h);
        y1[2] = mul_dir_i<DIR>(y1[2]);
        y1[3] = rot<DIR>(y1[3], -h, h);
        dft4<DIR>(y0[0], y0[1], y0[2], y0[3]);
        dft4<DIR>(y1[0], y1[1], y1[2], y1[3]);
        FCD_UNROLL
        for (int k2 = 0; k2 < 4; ++k2) {
            v[(2 * k2) * S] = y0[k2];
            v[(2 * k2 + 1) * S] = y1[k2];
        }
    }
};
template <int S, int DIR, class T> struct Dft<16, S, DIR, T> {
    // n = 4*n1 + n2, k = k1 + 4*k2
    FCD_HD static void run(cx<T>* v) {
        const T c1 = T(0.92387953251128675613), s1 = T(0.38268343236508977173);
        const T h = T(0.70710678118654752440);
        cx<T> y[4][4];  // y[k1][n2]
        FCD_UNROLL
        for (int n2 = 0; n2 < 4; ++n2) {
            cx<T> a0 = v[n2 * S], a1 = v[(4 + n2) * S], a2 = v[(8 + n2) * S], a3 = v[(12 + n2) * S];
            dft4<DIR>(a0, a1, a2, a3);
            y[0][n2] = a0; y[1][n2] = a1; y[2][n2] = a2; y[3][n2] = a3;
        }
        // twiddles W16^(n2*k1)
        y[1][1] = rot<DIR>(y[1][1], c1, s1);
        y[1][2] = rot<DIR>(y[1][2], h, h);
        y[1][3] = rot<DIR>(y[1][3], s1, c1);
        y[2][1] = rot<DIR>(y[2][1], h, h);
        y[2][2] = mul_dir_i<DIR>(y[2][2]);
        y[2][3] = rot<DIR>(y[2][3], -h, h);
        y[3][1] = rot<DIR>(y[3][1], s1, c1);
        y[3][2] = rot<DIR>(y[3][2], -h, h);
        y[3][3] = rot<DIR>(y[3][3], -c1, -s1);
        FCD_UNROLL
        for (int k1 = 0; k1 < 4; ++k1) {
            dft4<DIR>(y[k1][0], y[k1][1], y[k1][2], y[k1][3]);
            FCD_UNROLL
            for (int k2 = 0; k2 < 4; ++k2) v[(k1 + 4 * k2) * S] = y[k1][k2];
        }
    }
};

// ---------------------------------------------------------------------------------------
// radix plans
// ---------------------------------------------------------------------------------------
template <int L> struct Plan;
template <> struct Plan<64>   { static constexpr int R1 = 4,  R2 = 1,  R3 = 16; };
template <> struct Plan<128>  { static constexpr int R1 = 8,  R2 = 1,  R3 = 16; };
template <> struct Plan<256>  { static constexpr int R1 = 16, R2 = 1,  R3 = 16; };
template <> struct Plan<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8; };
template <> struct Plan<1024> { static constexpr int R1 = 8,  R2 = 8,  R3 = 16; };
template <> struct Plan<2048> { static constexpr int R1 = 8,  R2 = 16, R3 = 16; };
template <> struct Plan<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16; };

// Radix order of the COLUMN kernels (K2, K4), whose warps hold a few consecutive rows of SEVERAL transforms (lanes =
// 32/G rows x G columns, so that a warp request covers whole sectors of the column-blocked planes).  With that lane
// mapping the scatter of a radix-8 first pass hits every bank twice (16 of K4's 80 shared-memory instructions per
// transform, 18 % of its wavefronts in ncu); a radix-16 first pass writes rows 17 slots apart and is conflict free, like
// every other access pattern of both orders (simulated for the half-warp x bank matrix; ncu confirms).
template <int L> struct ColPlan : Plan<L> {};
template <> struct ColPlan<2048> { static constexpr int R1 = 16, R2 = 8, R3 = 16; };
template <> struct ColPlan<1024> { static constexpr int R1 = 16, R2 = 8, R3 = 8; };
template <> struct ColPlan<512>  { static constexpr int R1 = 16, R2 = 4, R3 = 8; };

// Radix order of the ROW kernels (K1, K3, K5, RowPhaseFwd; lanes = consecutive elements of one transform).  The first
// pass stays radix 8 (K3's pruned pass needs it); for 1024 and 512 the order of the other two matters: a radix-8 middle
// pass scatters with stride 8 behind a radix-8 first pass and conflicts two ways, radix 16 in the middle does not.
template <int L> struct RowPlan : Plan<L> {};
template <> struct RowPlan<1024> { static constexpr int R1 = 8, R2 = 16, R3 = 8; };
template <> struct RowPlan<512>  { static constexpr int R1 = 8, R2 = 16, R3 = 4; };

// Twiddle regeneration (per plan type): a radix-16 butterfly loads the powers 1, 2, 4, 8 of its twiddle and multiplies
// the other eleven together -- two packed instructions each instead of an 8-byte shared-memory read.  Measured on a
// B200 at 2048^2 (profiles/README.md, "twiddle regeneration"): K1 RowFwd 6.54 -> 6.23 us and K5 RowInv 5.53 -> 5.38 us
// per frame, K3 and K4 unchanged (they are not bound by shared-memory bandwidth), K2 4.79 -> 4.93 us.  So only the
// plans of K1 and K5 ask for it; -DFCD_TW_REGEN=1|2 forces it everywhere (2: radix-8 butterflies too) for A/B runs.
// Products of two rounded table entries are good to about 1.5 ulp instead of 0.5.
template <class P, class = void> struct TwRegen { static constexpr int value = 0; };
template <class P> struct TwRegen<P, std::void_t<decltype(P::TW_REGEN)>> { static constexpr int value = P::TW_REGEN; };
template <int L> struct RegenRowPlan : RowPlan<L> { static constexpr int TW_REGEN = 1; };

// smem slot of logical element p: one pad element per 16 (keeps radix-strided writes of
// the first pass and 16-aligned runs of the later passes bank-conflict free)
FCD_HD int fft_pos(int p) { return p + (p >> 4); }
// padded offset of a multiple of 16 (exact: fft_pos(p + c) == fft_pos(p) + fft_padc(c) when c % 16 == 0)
FCD_HD constexpr int fft_padc(int c) { return c + (c >> 4); }
// slot of the m-th owned element t + TPF*m (natural strided ownership)
template <int TPF> FCD_HD int fft_nat(int t, int m) {
    return (TPF % 16 == 0) ? fft_pos(t) + fft_padc(TPF * m) : fft_pos(t + TPF * m);
}

template <int L, int DIR, class T, class P = Plan<L>>
struct Fft {
    static constexpr int R1 = P::R1, R2 = P::R2, R3 = P::R3;
    static_assert(R1 * R2 * R3 == L, "radix plan must multiply to L");
    static constexpr int TPF = L / 16;         // threads per transform
    static constexpr int SMEM = L + L / 16;    // elements of the padded exchange buffer
    static constexpr bool THREE = (R2 != 1);

    // Twiddle table, laid out per pass as [a][k] so that consecutive threads (consecutive k)
    // read consecutive elements:  pass with radix R after prior product PP needs
    // W_{PP*R}^{k*a}, k in [0,PP), a in [1,R)  ->  table[off + a*PP + k].
    // Block of the middle pass (three-pass plans) first, then the block of the last pass.
    static constexpr int TW_MID = THREE ? R1 * R2 : 0;
    static constexpr int TW_W8 = TW_MID + L;          // W_8^m, m = 0..7 (pruned first pass)
    static constexpr int TW_ELEMS = TW_W8 + 8;

    // host: build the table (forward sign; DIR=+1 conjugates on load)
    static std::vector<cx<T>> make_table() {
        std::vector<cx<T>> t((size_t)TW_ELEMS);
        auto fill = [&](int off, int R, int PP) {
            for (int a = 0; a < R; ++a)
                for (int k = 0; k < PP; ++k) {
                    const long double ang = -2.0L * 3.14159265358979323846264338327950288L * ((long long)k * a) / (PP * R);
                    t[(size_t)off + (size_t)a * PP + k] = mk<T>((T)cosl(ang), (T)sinl(ang));
                }
        };
        if (THREE) fill(0, R2, R1);
        fill(TW_MID, R3, R1 * R2);
        for (int m = 0; m < 8; ++m) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * m / 8;
            t[(size_t)TW_W8 + m] = mk<T>((T)cosl(ang), (T)sinl(ang));
        }
        // exact zeros / ones where the long-double evaluation leaves 1e-20 residue
        for (int m = 0; m < 8; m += 2) {
            const int q = m / 2;    // exp(-i pi q / 2)
            t[(size_t)TW_W8 + m] = mk<T>(T(q == 0 ? 1 : (q == 2 ? -1 : 0)), T(q == 1 ? -1 : (q == 3 ? 1 : 0)));
        }
        return t;
    }

    // Twiddles W^(k*a), a = 1..R-1, of one butterfly: R - 1 loads from the [a][k] table, or (TwRegen above) the
    // power-of-two powers loaded and the rest multiplied together.
#if defined(FCD_TW_REGEN)
    static constexpr int REGEN = FCD_TW_REGEN;
#else
    static constexpr int REGEN = TwRegen<P>::value;
#endif
    template <int R, int PP>
    FCD_HD static void twiddles(cx<T>* w, const cx<T>* __restrict__ tw) {
        if constexpr ((R == 16 && REGEN >= 1) || (R == 8 && REGEN >= 2)) {
            FCD_UNROLL
            for (int a = 1; a < R; a <<= 1) {
                const cx<T> x = tw[a * PP];
                w[a] = DIR < 0 ? x : conj(x);
            }
            w[3] = w[1] * w[2];
            w[5] = w[1] * w[4];
            w[6] = w[2] * w[4];
            w[7] = w[3] * w[4];
            if constexpr (R == 16) {
                FCD_UNROLL
                for (int a = 1; a < 8; ++a) w[8 + a] = w[8] * w[a];
            }
        } else {
            FCD_UNROLL
            for (int a = 1; a < R; ++a) {
                const cx<T> x = tw[a * PP];
                w[a] = DIR < 0 ? x : conj(x);
            }
        }
    }

    // gather butterfly inputs of a pass with radix R and prior product PP; apply twiddles
    template <int R, int PP>
    FCD_HD static void gather(cx<T>* v, int t, const cx<T>* s, const cx<T>* __restrict__ table) {
        constexpr int NB = 16 / R;
        FCD_UNROLL
        for (int ii = 0; ii < NB; ++ii) {
            const int i = t + TPF * ii;
            const int k = i & (PP - 1);
            constexpr int OFF = (THREE && PP == R1) ? 0 : TW_MID;
            constexpr bool FAST = ((L / R) % 16 == 0);
            const cx<T>* sb = s + fft_pos(i);
            cx<T> w[R];
            if constexpr (PP > 1) twiddles<R, PP>(w, table + OFF + k);
            FCD_UNROLL
            for (int a = 0; a < R; ++a) {
                cx<T> val = FAST ? sb[fft_padc(a * (L / R))] : s[fft_pos(i + a * (L / R))];
                if (PP > 1 && a > 0) val = val * w[a];
                v[ii + NB * a] = val;
            }
        }
    }
    // two independent transforms (same length, same direction) sharing each twiddle load
    template <int R, int PP>
    FCD_HD static void gather2(cx<T>* v0, cx<T>* v1, int t, const cx<T>* s0, const cx<T>* s1,
                               const cx<T>* __restrict__ table) {
        constexpr int NB = 16 / R;
        FCD_UNROLL
        for (int ii = 0; ii < NB; ++ii) {
            const int i = t + TPF * ii;
            const int k = i & (PP - 1);
            constexpr int OFF = (THREE && PP == R1) ? 0 : TW_MID;
            constexpr bool FAST = ((L / R) % 16 == 0);
            const int pb = fft_pos(i);
            cx<T> w[R];
            if constexpr (PP > 1) twiddles<R, PP>(w, table + OFF + k);
            FCD_UNROLL
            for (int a = 0; a < R; ++a) {
                const int pos = FAST ? pb + fft_padc(a * (L / R)) : fft_pos(i + a * (L / R));
                cx<T> x0 = s0[pos], x1 = s1[pos];
                if (PP > 1 && a > 0) {
                    x0 = x0 * w[a];
                    x1 = x1 * w[a];
                }
                v0[ii + NB * a] = x0;
                v1[ii + NB * a] = x1;
            }
        }
    }

    template <int R>
    FCD_HD static void butterflies(cx<T>* v) {
        constexpr int NB = 16 / R;
        FCD_UNROLL
        for (int ii = 0; ii < NB; ++ii) Dft<R, NB, DIR, T>::run(v + ii);
    }
    template <int R, int PP>
    FCD_HD static void scatter(const cx<T>* v, int t, cx<T>* s) {
        constexpr int NB = 16 / R;
        FCD_UNROLL
        for (int ii = 0; ii < NB; ++ii) {
            const int i = t + TPF * ii;
            const int k = i & (PP - 1);
            const int j = (i - k) * R + k;
            // j + a*PP never carries out of the low nibble in a way the constants below miss:
            //   PP == 1 : j = i*R (R | 16), a < R            -> fft_pos(j) + a
            //   PP == 8 : low nibble of j is k < 8            -> fft_pos(j) + 8a + (a >> 1)
            //   PP % 16 == 0                                  -> fft_pos(j) + fft_padc(a*PP)
            constexpr bool FAST = (PP == 1 && 16 % R == 0) || PP == 8 || (PP % 16 == 0);
            cx<T>* sb = s + fft_pos(j);
            FCD_UNROLL
            for (int a = 0; a < R; ++a) {
                if (FAST) sb[PP == 1 ? a : (PP == 8 ? 8 * a + (a >> 1) : fft_padc(a * PP))] = v[ii + NB * a];
                else s[fft_pos(j + a * PP)] = v[ii + NB * a];
            }
        }
    }

    // ---- pruned first pass (radix 8): the butterfly `i` (= t + TPF*ii) has a single non-zero
    // input x, sitting at input index r of the butterfly.  Its outputs are x * W_8^(DIR*r*k),
    // k = 0..7: one inexact multiply (by W_8^r), the rest are multiplications by exact 0 / +-1
    // factors.  Writes the same slots as stepA would.
    FCD_HD static void stepA_single(cx<T> x, int r, int ii, int t, cx<T>* s, const cx<T>* __restrict__ table) {
        static_assert(R1 == 8, "pruned first pass is written for radix 8");
        const cx<T>* w8 = table + TW_W8;            // forward sign; DIR = +1 conjugates
        const int rr = r & 7;
        cx<T> u = w8[rr], v = w8[(2 * rr) & 7];      // W_8^r and W_8^(2r) (the latter exact: 0 / +-1 entries)
        if (DIR > 0) { u = conj(u); v = conj(v); }
        const T sg = w8[(4 * rr) & 7].x;             // W_8^(4r) = (-1)^r
        const cx<T> x0 = x;
        const cx<T> x1 = x * u;
        const cx<T> x2 = x0 * v;
        const cx<T> x3 = x1 * v;
        const int i = t + TPF * ii;
        cx<T>* sb = s + fft_pos(i * R1);
        sb[0] = x0; sb[1] = x1; sb[2] = x2; sb[3] = x3;
        sb[4] = scale(x0, sg); sb[5] = scale(x1, sg); sb[6] = scale(x2, sg); sb[7] = scale(x3, sg);
    }

    // ---- the four steps; a block-wide barrier is required between consecutive steps ----
    // A: v (natural ownership)  -> pass-1 butterflies -> smem
    FCD_HD static void stepA(cx<T>* v, int t, cx<T>* s) {
        butterflies<R1>(v);
        scatter<R1, 1>(v, t, s);
    }
    // B: smem -> v (twiddled)          [three-pass plans only]
    FCD_HD static void stepB(cx<T>* v, int t, const cx<T>* s, const cx<T>* __restrict__ table) {
        if constexpr (THREE) gather<R2, R1>(v, t, s, table);
    }
    // C: pass-2 butterflies -> smem    [three-pass plans only]
    FCD_HD static void stepC(cx<T>* v, int t, cx<T>* s) {
        if constexpr (THREE) {
            butterflies<R2>(v);
            scatter<R2, R1>(v, t, s);
        }
    }
    FCD_HD static void stepB2(cx<T>* v0, cx<T>* v1, int t, const cx<T>* s0, const cx<T>* s1,
                              const cx<T>* __restrict__ table) {
        if constexpr (THREE) gather2<R2, R1>(v0, v1, t, s0, s1, table);
    }
    FCD_HD static void stepD2(cx<T>* v0, cx<T>* v1, int t, const cx<T>* s0, const cx<T>* s1,
                              const cx<T>* __restrict__ table) {
        gather2<R3, R1 * R2>(v0, v1, t, s0, s1, table);
        butterflies<R3>(v0);
        butterflies<R3>(v1);
    }
    // D: smem -> last-pass butterflies -> v (natural ownership)
    FCD_HD static void stepD(cx<T>* v, int t, const cx<T>* s, const cx<T>* __restrict__ table) {
        gather<R3, R1 * R2>(v, t, s, table);
        butterflies<R3>(v);
    }
};

}  // namespace fcd
