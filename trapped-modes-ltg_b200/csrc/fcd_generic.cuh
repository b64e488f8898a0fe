// fcd_generic.cuh -- once-per-reference kernels (float64) and stage-level building blocks:
// generic batched row / column complex FFTs, the shifted high-passed magnitude spectrum used
// for carrier detection (pyfcd/fourier.py:18-35), candidate extraction (fourier.py:153-158),
// disk masking (pyfcd/carriers.py:17-20) and the ccsgn store (carriers.py:22-24).
// Same phase-structured form as fcd_kernels.cuh; see there for the conventions.
#pragma once
#include "fcd_kernels.cuh"

namespace fcd {

// ---- atomics that also compile for the CPU emulation ------------------------------------
FCD_HD void atomic_max_u64(unsigned long long* addr, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMax(addr, v);
#else
    if (v > *addr) *addr = v;
#endif
}
FCD_HD void atomic_add_f64(double* addr, double v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(addr, v);
#else
    *addr += v;
#endif
}
FCD_HD int atomic_inc_i32(int* addr) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(addr, 1);
#else
    return (*addr)++;
#endif
}
FCD_HD unsigned long long f64_bits(double v) {
    union { double d; unsigned long long u; } c;
    c.d = v;
    return c.u;
}

// ---- generic row transform: [rows][L] -----------------------------------------------------
template <class T>
struct GenRowsParams {
    const void* in;      // cx<T>[rows][L]  (kind 0) | float[rows][L] (1) | double[rows][L] (2)
    int in_kind;
    T sub;               // subtracted from real input (image - mean)
    cx<T>* out;          // [rows][L]
    const cx<T>* tw;
    int rows;
    T scale;
};

template <int L, int G, int DIR, class T>
struct GenRows : NoPrologue {
    using F = Fft<L, DIR, T>;
    using Params = GenRowsParams<T>;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = L / 16, THREADS = G * TPF, PHASES = 4;
    static constexpr int STRIDE = L + L / 16;
    static constexpr int SMEM_BYTES = G * STRIDE * (int)sizeof(cx<T>);
    struct State { cx<T> v[16]; };

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State& st) {
        const int g = tid / TPF, t = tid % TPF;
        cx<T>* s = reinterpret_cast<cx<T>*>(smem) + g * STRIDE;
        const int row = bx * G + g;
        const bool valid = row < p.rows;
        if constexpr (PH == 0) {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) {
                const long long idx = (long long)row * L + t + TPF * m;
                cx<T> val = mk<T>(T(0), T(0));
                if (valid) {
                    if (p.in_kind == 0) val = reinterpret_cast<const cx<T>*>(p.in)[idx];
                    else if (p.in_kind == 1) val = mk<T>((T)reinterpret_cast<const float*>(p.in)[idx] - p.sub, T(0));
                    else val = mk<T>((T)reinterpret_cast<const double*>(p.in)[idx] - p.sub, T(0));
                }
                st.v[m] = val;
            }
            F::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            F::stepB(st.v, t, s, p.tw);
        } else if constexpr (PH == 2) {
            F::stepC(st.v, t, s);
        } else {
            F::stepD(st.v, t, s, p.tw);
            if (valid) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) p.out[(long long)row * L + t + TPF * m] = scale(st.v[m], p.scale);
            }
        }
    }
};

// ---- generic column transform: [L][cols] -----------------------------------------------------
template <class T>
struct GenColsParams {
    const cx<T>* in;
    cx<T>* out;
    const cx<T>* tw;
    int cols;
    T scale;
};

template <int L, int G, int DIR, class T>
struct GenCols : NoPrologue {
    using F = Fft<L, DIR, T>;
    using Params = GenColsParams<T>;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = L / 16, THREADS = G * TPF, PHASES = 4;
    static constexpr int STRIDE = L + L / 16;
    static constexpr int SMEM_BYTES = G * STRIDE * (int)sizeof(cx<T>);
    struct State { cx<T> v[16]; };

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State& st) {
        const int g = tid / TPF, t = tid % TPF;
        cx<T>* s = reinterpret_cast<cx<T>*>(smem) + g * STRIDE;
        const int c = bx * G + g;
        const bool valid = c < p.cols;
        if constexpr (PH == 0) {
            FCD_UNROLL
            for (int m = 0; m < 16; ++m)
                st.v[m] = valid ? p.in[(long long)(t + TPF * m) * p.cols + c] : mk<T>(T(0), T(0));
            F::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            F::stepB(st.v, t, s, p.tw);
        } else if constexpr (PH == 2) {
            F::stepC(st.v, t, s);
        } else {
            F::stepD(st.v, t, s, p.tw);
            if (valid) {
                FCD_UNROLL
                for (int m = 0; m < 16; ++m) p.out[(long long)(t + TPF * m) * p.cols + c] = scale(st.v[m], p.scale);
            }
        }
    }
};

// ---- elementwise kernels: one phase, 256 threads, grid-stride over items ---------------------
struct SumParams {
    const void* in;
    int is_f64;
    long long n;
    double* out;   // single accumulator, zero-initialised
    int nblocks;
};
struct SumKernel : NoPrologue {
    using Params = SumParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 2, SMEM_BYTES = THREADS * (int)sizeof(double);
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State&) {
        double* part = reinterpret_cast<double*>(smem);
        if constexpr (PH == 0) {
            double acc = 0.0;
            for (long long i = (long long)bx * THREADS + tid; i < p.n; i += (long long)p.nblocks * THREADS)
                acc += p.is_f64 ? reinterpret_cast<const double*>(p.in)[i] : (double)reinterpret_cast<const float*>(p.in)[i];
            part[tid] = acc;
        } else {
            if (tid == 0) {
                double acc = 0.0;
                for (int q = 0; q < THREADS; ++q) acc += part[q];
                atomic_add_f64(p.out, acc);
            }
        }
    }
};

// ---- arbitrary sizes: Bluestein's chirp convolution around the power-of-two transforms ---------------------------
// X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n],  c[n] = exp(-i pi n^2 / N)  -- a length-N DFT as a circular
// convolution of length L >= 2N - 1 (a power of two).  In two dimensions the chirps are outer products, so one
// padded (L0 x L1) forward transform, a multiplication by the precomputed spectrum of conj(c0) (x) conj(c1) and one
// inverse transform give fft2 of any rows x cols image (scipy.fft.fft2 / ifft2 of the reference take any shape).
struct BluParams {
    const void* in;        // pad: [H][W] cx<double> (kind 0) | float (1) | double (2);  crop / mul: cd [L0][L1]
    int in_kind;
    double sub;            // subtracted from real input
    cd* out;               // pad / mul: [L0][L1];  crop: [H][W]
    const cd* c0;          // [H] chirp along rows
    const cd* c1;          // [W] chirp along columns
    const cd* fb;          // mul: spectrum of the chirp kernel [L0][L1]
    int H, W, L0, L1;
    int conj_io;           // inverse transform: conjugate the input (pad) / the output (crop)
    double scale;
    int nblocks;
};
struct BluPad : NoPrologue {
    using Params = BluParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.L0 * p.L1;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS) {
            const int r = (int)(i / p.L1), c = (int)(i % p.L1);
            cd v = mk<double>(0.0, 0.0);
            if (r < p.H && c < p.W) {
                const long long j = (long long)r * p.W + c;
                if (p.in_kind == 0) v = reinterpret_cast<const cd*>(p.in)[j];
                else if (p.in_kind == 1) v = mk<double>((double)reinterpret_cast<const float*>(p.in)[j] - p.sub, 0.0);
                else v = mk<double>(reinterpret_cast<const double*>(p.in)[j] - p.sub, 0.0);
                if (p.conj_io) v = conj(v);
                v = v * (p.c0[r] * p.c1[c]);
            }
            p.out[i] = v;
        }
    }
};
struct BluMul : BluPad {
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.L0 * p.L1;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS)
            p.out[i] = reinterpret_cast<const cd*>(p.in)[i] * p.fb[i];
    }
};
// out = (outer product of two vectors), used once per plan for the chirp kernel conj(c0) (x) conj(c1)
struct BluOuter : BluPad {
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.L0 * p.L1;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS)
            p.out[i] = p.c0[i / p.L1] * p.c1[i % p.L1];
    }
};
struct BluCrop : BluPad {
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.H * p.W;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS) {
            const int r = (int)(i / p.W), c = (int)(i % p.W);
            cd v = reinterpret_cast<const cd*>(p.in)[(long long)r * p.L1 + c] * (p.c0[r] * p.c1[c]);
            if (p.conj_io) v = conj(v);
            p.out[i] = scale(v, p.scale);
        }
    }
};

// float64 image -> float32 copy (the reference image as the masked workflow substitutes it, analyze.py:231)
struct NarrowParams {
    const double* in;
    float* out;
    long long n;
    int nblocks;
};
struct NarrowF64 : NoPrologue {
    using Params = NarrowParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        for (long long i = (long long)bx * THREADS + tid; i < p.n; i += (long long)p.nblocks * THREADS) p.out[i] = (float)p.in[i];
    }
};

// shifted |F| with high-pass zeroing and global max (fourier.py:18-23,34-35).  Magnitudes of
// columns beyond W/2 are mirrored from the computed half so that conjugate bins are exactly
// equal, as scipy's Hermitian fill makes them.
struct SpecMagParams {
    const cd* spec;          // [H][W] unshifted
    double* mag;             // [H][W] shifted
    const double* kr_sq;     // [H] shifted wavenumber^2 (calibration 1)
    const double* kc_sq;     // [W]
    double kmin_sq;
    unsigned long long* maxbits;
    int H, W, nblocks;
};
struct SpecMag : NoPrologue {
    using Params = SpecMagParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.H * p.W;
        double best = 0.0;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS) {
            const int r = (int)(i / p.W), c = (int)(i % p.W);
            int kr = (r + p.H - p.H / 2) % p.H, kc = (c + p.W - p.W / 2) % p.W;   // unshifted index of shifted (r,c), odd sizes too
            if (kc > p.W / 2) { kr = (p.H - kr) % p.H; kc = p.W - kc; }
            const cd z = p.spec[(long long)kr * p.W + kc];
            double m = hypot(z.x, z.y);
            if (!((p.kr_sq[r] + p.kc_sq[c]) > p.kmin_sq)) m = 0.0;
            p.mag[i] = m;
            if (m > best) best = m;
        }
        atomic_max_u64(p.maxbits, f64_bits(best));
    }
};

struct Candidate { int r, c; double v; };
struct CandidatesParams {
    const double* img;   // [H][W]
    double threshold;
    Candidate* list;
    int* count;
    int capacity;
    int H, W, nblocks;
};
struct Candidates : NoPrologue {
    using Params = CandidatesParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.H * p.W;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS) {
            const int r = (int)(i / p.W), c = (int)(i % p.W);
            if (r == 0 || c == 0 || r == p.H - 1 || c == p.W - 1) continue;  // fourier.py:155-158
            const double v = p.img[i];
            if (v > p.threshold) {
                const int slot = atomic_inc_i32(p.count);
                if (slot < p.capacity) { Candidate cand; cand.r = r; cand.c = c; cand.v = v; p.list[slot] = cand; }
            }
        }
    }
};

// out = spec * disk mask of carrier i (mask given by its chord table in shifted coordinates)
struct MaskMulParams {
    const cd* spec;
    cd* out;
    const int* chord_lo;   // [ncp] for this carrier
    const int* chord_hi;
    int c_lo, nc;
    int H, W, nblocks;
    uint8_t* mask_out;     // optional [H][W] unshifted boolean mask (Carrier.mask)
};
struct MaskMul : NoPrologue {
    using Params = MaskMulParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long n = (long long)p.H * p.W;
        for (long long i = (long long)bx * THREADS + tid; i < n; i += (long long)p.nblocks * THREADS) {
            const int kr = (int)(i / p.W), kc = (int)(i % p.W);
            const int r = (kr + p.H / 2) % p.H, c = (kc + p.W / 2) % p.W;
            const int cc = c - p.c_lo;
            bool keep = false;
            if (cc >= 0 && cc < p.nc) keep = (r >= p.chord_lo[cc]) && (r <= p.chord_hi[cc]);
            if (p.out) p.out[i] = keep ? p.spec[i] : mk<double>(0.0, 0.0);
            if (p.mask_out) p.mask_out[i] = keep ? 1 : 0;
        }
    }
};

// ccsgn = conj(g) (g already carries the 1/(H*W) of ifft2)
struct CcsgnStoreParams {
    const cd* g;
    cf* out_f;    // [H][W] complex64 (pipeline copy)
    cd* out_d;    // optional complex128 copy
    float* theta; // optional angle(ccsgn) as float32 (what the demodulation kernel reads): element i at theta[2 * i],
                  // the two carriers interleaved so that K3 fetches both angles of a pixel in one 8-byte load
    long long n;
    int nblocks;
};
struct CcsgnStore : NoPrologue {
    using Params = CcsgnStoreParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;   // 0: block-wide barrier
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int THREADS = 256, PHASES = 1, SMEM_BYTES = 16;
    struct State { int dummy; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        for (long long i = (long long)bx * THREADS + tid; i < p.n; i += (long long)p.nblocks * THREADS) {
            const cd z = conj(p.g[i]);
            if (p.out_f) p.out_f[i] = mk<float>((float)z.x, (float)z.y);
            if (p.out_d) p.out_d[i] = z;
            if (p.theta) p.theta[2 * i] = (float)atan2(z.y, z.x);
        }
    }
};

}  // namespace fcd
