// fcd_plan.inl -- plan object and C ABI (include/fcd_b200.h).  Included by fcd_b200.cu
// (CUDA, the product) and by tests/emul/fcd_emul.cpp (-DFCD_EMULATE, CPU emulation used
// only by the CPU test-suite).  Host logic here mirrors, with citations, the scalar parts of
// the reference: peak selection (pyfcd/fourier.py:25-41), disk chords (pyfcd/carriers.py:17-20),
// carrier wavevectors (carriers.py:12), wavenumber vectors (fourier.py:44-57,76-92).
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <unordered_map>

#include "fcd_b200.h"
#include "fcd_generic.cuh"
#include "fcd_mask.cuh"
#include "fcd_unwrap.cuh"
#include "fcd_temporal.cuh"
#include "fcd_launch.cuh"

namespace fcd {

// groups per block for each transform length (threads = G * L/16)
template <int L> struct Tune;
template <> struct Tune<64>   { static constexpr int GROW = 16, GCOL = 16, GDEM = 16, GGEN = 16; };
template <> struct Tune<128>  { static constexpr int GROW = 8,  GCOL = 8,  GDEM = 8,  GGEN = 8; };
template <> struct Tune<256>  { static constexpr int GROW = 8,  GCOL = 8,  GDEM = 8,  GGEN = 8; };
template <> struct Tune<512>  { static constexpr int GROW = 4,  GCOL = 4,  GDEM = 4,  GGEN = 4; };
#ifndef FCD_T1024_GDEM
#define FCD_T1024_GDEM 8
#endif
template <> struct Tune<1024> { static constexpr int GROW = 4,  GCOL = 4,  GDEM = FCD_T1024_GDEM,  GGEN = 4; };
#ifndef FCD_T2048_GROW
#define FCD_T2048_GROW 2
#endif
#ifndef FCD_T2048_GCOL
#define FCD_T2048_GCOL 4
#endif
#ifndef FCD_T2048_GDEM
#define FCD_T2048_GDEM 4
#endif
template <> struct Tune<2048> { static constexpr int GROW = FCD_T2048_GROW, GCOL = FCD_T2048_GCOL, GDEM = FCD_T2048_GDEM, GGEN = 2; };
template <> struct Tune<4096> { static constexpr int GROW = 2,  GCOL = 2,  GDEM = 2,  GGEN = 2; };

#define FCD_CASE_L(N, ...) case N: { constexpr int L = N; __VA_ARGS__; } break;
#define FCD_DISPATCH_L(value, ...)                                                      \
    switch (value) {                                                                    \
        FCD_CASE_L(64, __VA_ARGS__) FCD_CASE_L(128, __VA_ARGS__) FCD_CASE_L(256, __VA_ARGS__) \
        FCD_CASE_L(512, __VA_ARGS__) FCD_CASE_L(1024, __VA_ARGS__) FCD_CASE_L(2048, __VA_ARGS__) \
        FCD_CASE_L(4096, __VA_ARGS__)                                                   \
        default: rt::fail("unsupported transform length");                              \
    }

static inline bool supported_dim(int n) { return n >= 64 && n <= 4096 && (n & (n - 1)) == 0; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// fftfreq(n, d)[j] as scipy computes it: integer * (1.0 / (n*d))          (fourier.py:56)
static inline double fftfreq_at(int n, double d, int j) {
    const double val = 1.0 / (n * d);
    const int m = j < (n + 1) / 2 ? j : j - n;
    return m * val;
}
// fftshift(fftfreq(n, d))[j]
static inline double fftfreq_shifted_at(int n, double d, int j) { return fftfreq_at(n, d, (j + n - n / 2) % n); }

struct PlanImpl {
    int H = 0, W = 0, chunk = 1;
    long long launches = 0;

    rt::DevBuf<cf> tw_w_f, tw_h_f;
    rt::DevBuf<cd> tw_w_d, tw_h_d;

    // per-reference float64 scratch
    rt::DevBuf<cd> spec, tmp;
    rt::DevBuf<double> mag, kr_sq, kc_sq, sum;
    rt::DevBuf<unsigned long long> maxbits;
    rt::DevBuf<Candidate> cand;
    rt::DevBuf<int> cand_count;
    static constexpr int kCandCap = 1 << 18;

    // bound reference
    bool bound = false;
    bool spec_valid = false;   // spec holds fft2(reference)
    int peaks[4] = {0, 0, 0, 0};
    double radius = 0, cal = 1, height = 1;
    int c_lo[2] = {0, 0}, nc[2] = {0, 0}, ncp = 4;
    std::vector<int> h_lo, h_hi;           // [2][ncp]
    rt::DevBuf<int> chord_lo, chord_hi;
    rt::DevBuf<cf> ccsgn;                  // [2][H][W] complex64 copy (Carrier.ccsgn)
    rt::DevBuf<float> theta;               // [H][W][2] angle(ccsgn) of the two carriers, read by the demodulation kernel
    rt::DevBuf<float> kx, kxq;
    float dky = 0.f;
    double f[2][2] = {{0, 0}, {0, 0}};     // carrier wavevectors [k_row, k_col]
    double det = 1;

    // per-launch workspaces (chunk frames)
    rt::DevBuf<cf> w1, w2, w3, w4;
    rt::DevBuf<float> colphase, rowoff;
    int w4p = 0;

    template <class K>
    void launch(int gx, int gy, rt::stream_t s, const typename K::Params& p) {
        rt::launch<K>(gx, gy, s, p);
        ++launches;
    }

    // optional per-stage device timing (CUDA events on the launch stream), for bench.py
    static constexpr int kStages = 7;   // K1 K2 K3 K3b PhaseFix K4 K5
    bool profiling = false;
    rt::StageTimer timer;

    // ------------------------------------------------------------------ construction ----
    // Two kinds of plan.  FUSED: rows and cols powers of two in [64, 4096] -- the float32 pipeline K1..K5 and everything
    // around it.  GENERIC: any other shape with 2 <= rows, cols <= 2048 (the reference takes any shape, pyfcd/fcd.py:14):
    // the float64 per-reference path (carrier search, masks, ccsgn), fcd_fft2_c128, the unwrap and the residue count
    // work through Bluestein's chirp convolution on padded power-of-two transforms; the host layer builds the
    // per-frame path from those stage-level pieces (fcd_b200/generic.py).  fcd_execute and the structure-mask entry
    // points need a fused plan.
    bool generic = false;
    int L0 = 0, L1 = 0;                       // generic: padded transform sizes (powers of two >= 2N - 1)
    rt::DevBuf<cd> blu_c0, blu_c1, blu_fb, blu_a;

    static int blu_length(int n) {
        int l = 64;
        while (l < 2 * n - 1) l *= 2;
        return l;
    }
    static std::vector<cd> chirp(int n, int len, bool kernel) {
        // kernel == false: c[m] = exp(-i pi m^2 / n), m < n.  kernel == true: conj(c)[m] for |m| < n laid out circularly
        // in `len` slots.  m^2 is reduced mod 2n in integers so that the angle keeps full precision.
        std::vector<cd> v((size_t)(kernel ? len : n), mk<double>(0.0, 0.0));
        for (int m = 0; m < n; ++m) {
            const long long q = ((long long)m * m) % (2LL * n);
            const long double ang = 3.14159265358979323846264338327950288L * (long double)q / (long double)n;
            const cd w = mk<double>((double)cosl(ang), (double)(kernel ? sinl(ang) : -sinl(ang)));
            v[(size_t)m] = w;
            if (kernel && m) v[(size_t)(len - m)] = w;
        }
        return v;
    }
    void require_fused(const char* what) const {
        if (generic) rt::fail(std::string(what) + " needs a plan whose rows and cols are powers of two in [64, 4096]");
    }

    void init(int rows, int cols, int frames_per_launch) {
        if (frames_per_launch < 1 || frames_per_launch > 16384) rt::fail("frames_per_launch out of range");
        generic = !(supported_dim(rows) && supported_dim(cols));
        if (generic && (rows < 2 || cols < 2 || rows > 2048 || cols > 2048))
            rt::fail("unsupported shape: rows and cols must be powers of two in [64, 4096] (fused float32 pipeline) "
                     "or anything in [2, 2048] (float64 stage-level path)");
        H = rows; W = cols; chunk = frames_per_launch;
        sms = rt::sm_count();
        L0 = generic ? blu_length(H) : H;
        L1 = generic ? blu_length(W) : W;
        FCD_DISPATCH_L(L1, {
            if (!generic) tw_w_f.upload(Fft<L, -1, float, RowPlan<L>>::make_table(), nullptr);   // K1, K3, K5
            tw_w_d.upload(Fft<L, -1, double>::make_table(), nullptr);
        })
        FCD_DISPATCH_L(L0, {
            if (!generic) tw_h_f.upload(Fft<L, -1, float, ColPlan<L>>::make_table(), nullptr);   // K2, K4
            tw_h_d.upload(Fft<L, -1, double>::make_table(), nullptr);
        })
        w4p = W / 2 + 4;
        if (generic) {
            blu_c0.upload(chirp(H, L0, false), nullptr);
            blu_c1.upload(chirp(W, L1, false), nullptr);
            rt::DevBuf<cd> k0, k1;
            k0.upload(chirp(H, L0, true), nullptr);
            k1.upload(chirp(W, L1, true), nullptr);
            const size_t n = (size_t)L0 * L1;
            blu_fb.alloc(n);
            blu_a.alloc(n);
            const int nb = elem_blocks((long long)n);
            BluParams bp{nullptr, 0, 0.0, blu_fb.ptr, k0.ptr, k1.ptr, nullptr, H, W, L0, L1, 0, 1.0, nb};
            launch<BluOuter>(nb, 1, nullptr, bp);
            fft2_pow2(blu_fb.ptr, 0, 0.0, blu_fb.ptr, -1, L0, L1, nullptr);
            rt::sync(nullptr);
        }
    }

    void ensure_reference_scratch() {
        const size_t n = (size_t)H * W;
        spec.alloc(n);
        tmp.alloc(n);
    }

    // ------------------------------------------------------------------ float64 fft2 ----
    // power-of-two transform of a rows x cols array with the plan's tables (rows == L0, cols == L1)
    void fft2_pow2(const void* in, int kind, double sub, cd* out, int dir, int rows, int cols, rt::stream_t s) {
        FCD_DISPATCH_L(cols, {
            constexpr int G = Tune<L>::GGEN;
            GenRowsParams<double> p{in, kind, sub, out, tw_w_d.ptr, rows, 1.0};
            if (dir < 0) launch<GenRows<L, G, -1, double>>(ceil_div(rows, G), 1, s, p);
            else launch<GenRows<L, G, +1, double>>(ceil_div(rows, G), 1, s, p);
        })
        FCD_DISPATCH_L(rows, {
            constexpr int G = Tune<L>::GGEN;
            GenColsParams<double> p{out, out, tw_h_d.ptr, cols, dir < 0 ? 1.0 : 1.0 / ((double)rows * cols)};
            if (dir < 0) launch<GenCols<L, G, -1, double>>(ceil_div(cols, G), 1, s, p);
            else launch<GenCols<L, G, +1, double>>(ceil_div(cols, G), 1, s, p);
        })
    }
    // out = fft2(in) (dir=-1) or ifft2(in) (dir=+1, scaled 1/(HW)); in may alias out for kind 0
    void fft2_d(const void* in, int kind, double sub, cd* out, int dir, rt::stream_t s) {
        if (!generic) {
            fft2_pow2(in, kind, sub, out, dir, H, W, s);
            return;
        }
        // Bluestein in two dimensions; ifft2(x) = conj(fft2(conj(x))) / (H W)
        const int nb = elem_blocks((long long)L0 * L1);
        BluParams bp{in, kind, sub, blu_a.ptr, blu_c0.ptr, blu_c1.ptr, blu_fb.ptr, H, W, L0, L1, dir > 0 ? 1 : 0, 1.0, nb};
        launch<BluPad>(nb, 1, s, bp);
        fft2_pow2(blu_a.ptr, 0, 0.0, blu_a.ptr, -1, L0, L1, s);
        bp.in = blu_a.ptr;
        launch<BluMul>(nb, 1, s, bp);
        fft2_pow2(blu_a.ptr, 0, 0.0, blu_a.ptr, +1, L0, L1, s);
        bp.out = out;
        bp.scale = dir > 0 ? 1.0 / ((double)H * W) : 1.0;
        bp.nblocks = elem_blocks((long long)H * W);
        launch<BluCrop>(bp.nblocks, 1, s, bp);
    }

    int sms = 148;   // multiprocessors of the plan's device (init)
    int elem_blocks(long long n) const { return (int)std::min<long long>((n + 255) / 256, (long long)sms * 8); }

    // ------------------------------------------------------------------ carrier search ----
    double highpass_spectrum(const void* image, int is_f64, double* spectrum_out, rt::stream_t s) {
        ensure_reference_scratch();
        const long long n = (long long)H * W;
        mag.alloc(n);
        sum.alloc(1);
        maxbits.alloc(1);
        rt::dmemset(sum.ptr, 0, sizeof(double), s);
        rt::dmemset(maxbits.ptr, 0, sizeof(unsigned long long), s);
        const int nb = elem_blocks(n);
        launch<SumKernel>(nb, 1, s, SumParams{image, is_f64, n, sum.ptr, nb});
        double total = 0;
        rt::d2h(&total, sum.ptr, sizeof(double), s);
        const double mean = total / (double)n;                      // image - np.mean(image), fourier.py:18
        fft2_d(image, is_f64 ? 2 : 1, mean, spec.ptr, -1, s);
        spec_valid = false;
        if (kr_sq.count != (size_t)H || kc_sq.count != (size_t)W) {
            const double d = 1.0 / (2.0 * M_PI);                    // wavenumber(size, 1, shifted=True)
            std::vector<double> a(H), b(W);
            for (int j = 0; j < H; ++j) { const double k = fftfreq_shifted_at(H, d, j); a[j] = k * k; }
            for (int j = 0; j < W; ++j) { const double k = fftfreq_shifted_at(W, d, j); b[j] = k * k; }
            kr_sq.upload(a, s);
            kc_sq.upload(b, s);
        }
        const double kmin = 4.0 * M_PI / std::min(H, W);            // fourier.py:22
        launch<SpecMag>(nb, 1, s, SpecMagParams{spec.ptr, spectrum_out ? spectrum_out : mag.ptr, kr_sq.ptr,
                                                  kc_sq.ptr, kmin * kmin, maxbits.ptr, H, W, nb});
        if (spectrum_out) rt::d2d(mag.ptr, spectrum_out, n * sizeof(double), s);
        unsigned long long bits = 0;
        rt::d2h(&bits, maxbits.ptr, sizeof(bits), s);
        double mx;
        std::memcpy(&mx, &bits, sizeof(mx));
        return mx;
    }

    std::vector<std::array<int, 2>> peak_locations(const double* image, double threshold, int max_peaks,
                                                   rt::stream_t s) {
        const long long n = (long long)H * W;
        cand.alloc(kCandCap);
        cand_count.alloc(1);
        rt::dmemset(cand_count.ptr, 0, sizeof(int), s);
        const int nb = elem_blocks(n);
        launch<Candidates>(nb, 1, s, CandidatesParams{image, threshold, cand.ptr, cand_count.ptr, kCandCap, H, W, nb});
        int count = 0;
        rt::d2h(&count, cand_count.ptr, sizeof(int), s);
        if (count > kCandCap) rt::fail("too many pixels above the peak threshold (flat spectrum?)");
        std::vector<Candidate> c((size_t)count);
        if (count) rt::d2h(c.data(), cand.ptr, sizeof(Candidate) * (size_t)count, s);
        // raster order, then 8-connected labelling numbered by first pixel (skimage.measure.label)
        std::sort(c.begin(), c.end(), [](const Candidate& a, const Candidate& b) {
            return a.r != b.r ? a.r < b.r : a.c < b.c;
        });
        std::unordered_map<long long, int> at;
        at.reserve(c.size() * 2 + 1);
        for (int i = 0; i < count; ++i) at[(long long)c[i].r * W + c[i].c] = i;
        std::vector<int> lab((size_t)count, -1);
        int nlab = 0;
        std::vector<int> stack;
        for (int i = 0; i < count; ++i) {
            if (lab[i] >= 0) continue;
            lab[i] = nlab;
            stack.push_back(i);
            while (!stack.empty()) {
                const int j = stack.back();
                stack.pop_back();
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        auto it = at.find((long long)(c[j].r + dr) * W + (c[j].c + dc));
                        if (it != at.end() && c[j].c + dc >= 0 && c[j].c + dc < W && lab[it->second] < 0) {
                            lab[it->second] = nlab;
                            stack.push_back(it->second);
                        }
                    }
            }
            ++nlab;
        }
        struct Blob { double v; int r, c; };
        std::vector<Blob> blobs((size_t)nlab, Blob{-1.0, 0, 0});
        for (int i = 0; i < count; ++i) {      // raster order within a blob: first maximum wins
            Blob& b = blobs[lab[i]];
            if (c[i].v > b.v) b = Blob{c[i].v, c[i].r, c[i].c};
        }
        std::stable_sort(blobs.begin(), blobs.end(), [](const Blob& a, const Blob& b) { return a.v < b.v; });
        std::vector<std::array<int, 2>> out;
        for (int i = 0; i < nlab && i < max_peaks; ++i) out.push_back({blobs[i].r, blobs[i].c});
        return out;
    }

    // fourier.find_peaks: fourier.py:25-41
    void find_peaks(const void* image, int is_f64, int out[4], rt::stream_t s) {
        const double mx = highpass_spectrum(image, is_f64, nullptr, s);
        auto locs = peak_locations(mag.ptr, 0.5 * mx, 4, s);
        if (locs.empty()) throw std::domain_error("no carrier peak above threshold");
        const double d = 1.0 / (2.0 * M_PI);
        auto kof = [&](const std::array<int, 2>& p, double k[2]) {
            k[0] = fftfreq_shifted_at(H, d, p[0]);
            k[1] = fftfreq_shifted_at(W, d, p[1]);
        };
        int best = 0;
        double bestv = 0;
        for (size_t i = 0; i < locs.size(); ++i) {
            double k[2];
            kof(locs[i], k);
            const double v = std::fabs(std::atan2(k[0], k[1]));
            if (i == 0 || v < bestv) { best = (int)i; bestv = v; }
        }
        double k1[2];
        kof(locs[best], k1);
        int perp = 0;
        double perpv = 0;
        for (size_t i = 0; i < locs.size(); ++i) {
            double k[2];
            kof(locs[i], k);
            const double v = std::fabs(k1[0] * k[0] + k1[1] * k[1]);
            if (i == 0 || v < perpv) { perp = (int)i; perpv = v; }
        }
        out[0] = locs[best][0]; out[1] = locs[best][1];
        out[2] = locs[perp][0]; out[3] = locs[perp][1];
    }

    // ------------------------------------------------------------------ bind ----
    void compute_chords() {
        // disk ((r-r0)/R)^2 + ((c-c0)/R)^2 < 1 on the shifted grid, clipped to the array
        std::vector<int> lo[2], hi[2];
        for (int i = 0; i < 2; ++i) {
            const double r0 = peaks[2 * i], c0 = peaks[2 * i + 1];
            int first = -1, last = -1;
            std::vector<int> clo(W, 1), chi(W, 0);
            for (int c = 0; c < W; ++c) {
                const double dc = (c - c0) / radius;
                if (!(dc * dc < 1.0)) continue;
                int rlo = -1, rhi = -1;
                const int ra = std::max(0, (int)std::floor(r0 - radius) - 1);
                const int rb = std::min(H - 1, (int)std::ceil(r0 + radius) + 1);
                for (int r = ra; r <= rb; ++r) {
                    const double dr = (r - r0) / radius;
                    if (dr * dr + dc * dc < 1.0) { if (rlo < 0) rlo = r; rhi = r; }
                }
                if (rlo >= 0) {
                    clo[c] = rlo; chi[c] = rhi;
                    if (first < 0) first = c;
                    last = c;
                }
            }
            if (first < 0) { first = 0; last = -1; }
            c_lo[i] = first;
            nc[i] = last - first + 1;
            lo[i].assign(clo.begin() + first, clo.begin() + first + nc[i]);
            hi[i].assign(chi.begin() + first, chi.begin() + first + nc[i]);
        }
        ncp = std::max(4, (std::max(nc[0], nc[1]) + 3) / 4 * 4);
        h_lo.assign((size_t)2 * ncp, 1);
        h_hi.assign((size_t)2 * ncp, 0);
        for (int i = 0; i < 2; ++i)
            for (int c = 0; c < nc[i]; ++c) { h_lo[(size_t)i * ncp + c] = lo[i][c]; h_hi[(size_t)i * ncp + c] = hi[i][c]; }
    }

    void masked_inverse(int i, rt::stream_t s, uint8_t* mask_out) {
        // tmp = ifft2(spec * mask_i)
        const long long n = (long long)H * W;
        const int nb = elem_blocks(n);
        launch<MaskMul>(nb, 1, s, MaskMulParams{spec.ptr, mask_out ? nullptr : tmp.ptr, chord_lo.ptr + (size_t)i * ncp,
                                                  chord_hi.ptr + (size_t)i * ncp, c_lo[i], nc[i], H, W, nb, mask_out});
        if (!mask_out) fft2_d(tmp.ptr, 0, 0.0, tmp.ptr, +1, s);
    }

    void set_height(double h) {
        if (!(h != 0.0)) rt::fail("height must be non-zero");
        height = h;
    }

    void bind(const void* ref, int is_f64, const int pk[4], double rad, double calib, double h, rt::stream_t s) {
        if (!(rad > 0.0)) rt::fail("carrier radius must be positive");
        if (!(calib > 0.0)) rt::fail("calibration factor must be positive");
        for (int i = 0; i < 4; ++i) {
            const int lim = (i % 2 == 0) ? H : W;
            if (pk[i] < 0 || pk[i] >= lim) rt::fail("carrier peak outside the spectrum");
            peaks[i] = pk[i];
        }
        radius = rad; cal = calib;
        set_height(h);
        bound = false;
        ensure_reference_scratch();
        compute_chords();
        if (nc[0] <= 0 || nc[1] <= 0) rt::fail("empty carrier disk");
        chord_lo.upload(h_lo, s);
        chord_hi.upload(h_hi, s);

        // carrier wavevectors: pixel_to_wavenumber(shape, peak, cal)   (carriers.py:12)
        const double d = cal / (2.0 * M_PI);
        for (int i = 0; i < 2; ++i) {
            f[i][0] = fftfreq_shifted_at(H, d, peaks[2 * i]);
            f[i][1] = fftfreq_shifted_at(W, d, peaks[2 * i + 1]);
        }
        det = f[0][1] * f[1][0] - f[0][0] * f[1][1];                 // fcd.py:134-135

        // wavenumber vectors of integrate_in_fourier with the N//2+1 quirk (fourier.py:128-132); fused plans only
        if (!generic) {
            std::vector<float> vkx(W), vkxq(W);
            for (int j = 0; j < W; ++j) { vkx[j] = (float)fftfreq_at(W, d, j); vkxq[j] = vkx[j]; }
            vkxq[W / 2 + 1] = 0.f;
            kx.upload(vkx, s); kxq.upload(vkxq, s);
            dky = (float)fftfreq_at(H, d, 1);
        }

        // ccsgn_i = conj(ifft2(fft2(ref) * mask_i))   (carriers.py:22-24), float64 then stored c64
        fft2_d(ref, is_f64 ? 2 : 1, 0.0, spec.ptr, -1, s);
        spec_valid = true;
        const long long n = (long long)H * W;
        ccsgn.alloc((size_t)2 * n);
        theta.alloc((size_t)2 * n);
        for (int i = 0; i < 2; ++i) {
            masked_inverse(i, s, nullptr);
            const int nb = elem_blocks(n);
            launch<CcsgnStore>(nb, 1, s, CcsgnStoreParams{tmp.ptr, ccsgn.ptr + (size_t)i * n, nullptr,
                                                         theta.ptr + i, n, nb});
        }

        // workspaces of the fused pipeline
        if (!generic) {
            w1.alloc((size_t)chunk * 2 * ncp * H);
            w2.alloc((size_t)chunk * 2 * ncp * H);
            w3.alloc((size_t)chunk * w3_blocks(W) * H * 4);
            w4.alloc((size_t)chunk * H * w4p);
            colphase.alloc((size_t)chunk * 2 * H);
            rowoff.alloc((size_t)chunk * 2 * H);
        }
        rt::sync(s);
        bound = true;
    }

    // ------------------------------------------------------------------ execute ----
    // One wave of frames (at most `chunk`), as pointers to its first frame.  The workspaces w1..w4 hold the
    // wave; `wf` offsets into them let the back half of the pipeline run on a sub-range of a wave.
    struct Wave {
        const void* fr; int kind; int nf;
        float* ho; const uint8_t* mk; long long mask_stride;
    };
    float scale_demod() const { return (float)(1.0 / (2.0 * (double)H * (double)W)); }
    float scale_int() const { return (float)(1.0 / (2.0 * height * det * (double)H * (double)W)); }

    // K1 row forward, K2 column band-pass, K3 row demodulation (+ row unwrap when scan) -> w3, colphase, flags
    void stage_front(const Wave& a, float* po, int scan, int* flags, rt::stream_t s) {
        FCD_DISPATCH_L(W, {
            constexpr int G = Tune<L>::GROW;
            RowFwdParams p{a.fr, a.kind, nullptr, a.mk, a.mask_stride, w1.ptr, tw_w_f.ptr, H, ncp,
                           {nc[0], nc[1]}, {c_lo[0] - W / 2, c_lo[1] - W / 2}};
            p.reference = reference_f32();
            launch<RowFwd<L, G>>(H / (2 * G), a.nf, s, p);
        })
        if (profiling) timer.mark(s, 0);
        FCD_DISPATCH_L(H, {
            constexpr int G = Tune<L>::GCOL;
            ColBandParams p{w1.ptr, w2.ptr, tw_h_f.ptr, chord_lo.ptr, chord_hi.ptr, ncp, {nc[0], nc[1]}, scale_demod()};
            launch<ColBand<L, G>>(ceil_div(ncp, G), a.nf * 2, s, p);
        })
        if (profiling) timer.mark(s, 1);
        stage_demod(0, a.nf, po, scan, flags, s);
    }
    // K3 alone on the band-passed spectra w2 of frames [wf, wf + nf) of the current wave; writes w3 / colphase from
    // frame 0 on (the second look of unwrap "auto" demodulates a run of flagged frames again without redoing K1, K2)
    void stage_demod(int wf, int nf, float* po, int scan, int* flags, rt::stream_t s) {
        FCD_DISPATCH_L(W, {
            constexpr int G = Tune<L>::GDEM;
            RowDemodParams p{w2.ptr + (size_t)wf * 2 * H * ncp, theta.ptr, w3.ptr, colphase.ptr, po, tw_w_f.ptr, H, ncp,
                             {nc[0], nc[1]}, {c_lo[0] - W / 2, c_lo[1] - W / 2}, W / 2, scan, flags};
            bool pruned = false;
            if constexpr (RowPlan<L>::R1 == 8) {
                if (nc[0] <= L / 8 && nc[1] <= L / 8) {
                    pruned = true;
                    launch<RowDemod<L, G, true>>(nf, H / G, s, p);
                }
            }
            if (!pruned) launch<RowDemod<L, G, false>>(nf, H / G, s, p);
        })
        if (profiling) timer.mark(s, 2);
    }
    // K3b: rows linked along the anchor column; materialised phases get their row offsets
    void stage_link(int nf, float* po, int scan, rt::stream_t s) {
        launch<RowLink>(2, nf, s, RowLinkParams{colphase.ptr, rowoff.ptr, H, H / 2, scan});
        if (profiling) timer.mark(s, 3);
        if (po && scan) {
            launch<PhaseFix>(H, nf * 2, s, PhaseFixParams{po, rowoff.ptr, H, W});
            if (profiling) timer.mark(s, 4);
        }
    }
    // reliability-guided unwrap of the wrapped phases of frames [wf, wf + nf) of a wave, in place, then their
    // forward row transform into w3 (what K3 does in the scan modes)
    void stage_guided(float* po_wave, int wf, int nf, rt::stream_t s) {
        const long long n = (long long)H * W;
        float* po = po_wave + (long long)wf * 2 * n;
        unwrap_maps(po, nf * 2, po, s);
        if (profiling) timer.mark(s, 3);
        FCD_DISPATCH_L(W, {
            constexpr int G = Tune<L>::GROW;
            launch<RowPhaseFwd<L, G>>(H / G, nf, s, RowPhaseFwdParams{po, w3.ptr + (size_t)wf * w3_blocks(W) * H * 4, tw_w_f.ptr, H});
        })
        if (profiling) timer.mark(s, 4);
    }
    // K4 column integration, K5 row inverse for frames [wf, wf + nf) of the wave `a`
    void stage_back(const Wave& a, int wf, int nf, int scan, rt::stream_t s) {
        const long long n = (long long)H * W;
        FCD_DISPATCH_L(H, {
            constexpr int G = Tune<L>::GCOL;
            ColIntegrateParams p{w3.ptr + (size_t)wf * w3_blocks(W) * H * 4, rowoff.ptr + (size_t)wf * 2 * H, w4.ptr, tw_h_f.ptr,
                                 kx.ptr, kxq.ptr, dky, W, w4p,
                                 (float)f[0][0], (float)f[0][1], (float)f[1][0], (float)f[1][1], scale_int(), scan};
            launch<ColIntegrate<L, G>>(ceil_div(W / 2 + 1, G), nf, s, p);
        })
        if (profiling) timer.mark(s, 5);
        FCD_DISPATCH_L(W, {
            constexpr int G = Tune<L>::GROW;
            RowInvParams p{w4.ptr, a.ho + (long long)wf * n, a.mk ? a.mk + (long long)wf * a.mask_stride : nullptr,
                           a.mask_stride, tw_w_f.ptr, H, w4p};
            launch<RowInv<L, G>>(H / (2 * G), nf, s, p);
        })
        if (profiling) timer.mark(s, 6);
    }
    static Wave sub_wave(const Wave& a, int wf, int nf, long long n) {
        const size_t px = a.kind == 0 ? 4 : (a.kind == 1 ? 1 : 2);
        return Wave{static_cast<const unsigned char*>(a.fr) + (size_t)wf * n * px, a.kind, nf, a.ho + (long long)wf * n,
                    a.mk ? a.mk + (long long)wf * a.mask_stride : nullptr, a.mask_stride};
    }

    // unwrap = 3 ("auto", the drop-in's unwrap=True): statistics of the last fcd_execute call
    rt::DevBuf<int> frameflag, res_counts;
    long long auto_flagged = 0;               // frames with |phi| > pi/2 somewhere (probed for residues)
    std::vector<int> auto_guided;             // frames whose phases held residues (redone reliability-guided)
    static constexpr int kProbe = 16;         // frames per residue probe (bounds the phase scratch)

    // Second look of unwrap "auto", once per call: `idx` are the frames K3 flagged (ascending).  They are taken in
    // batches of kProbe; a batch whose frames are consecutive is used where it lies, any other batch is first copied
    // together (one device-to-device copy per frame: flagged frames that are few and far between still fill whole
    // launches) and its results copied back.  Per batch: K1-K3 without unwrap -> wrapped phases in scratch, residues
    // counted, and only the frames that hold residues are unwrapped along the reliability-guided tree and
    // re-integrated -- bit for bit what unwrap = 2 does.  A frame without residues keeps its scan result: there
    // every unwrapper gives the same integers.
    rt::DevBuf<unsigned char> g_frames, g_mask;
    rt::DevBuf<float> g_out;
    void auto_second_look(const std::vector<int>& idx, const void* frames, int kind, float* height_out, float* phases,
                          const uint8_t* mask, long long mask_stride, rt::stream_t s) {
        const long long n = (long long)H * W;
        const size_t px = kind == 0 ? 4 : (kind == 1 ? 1 : 2);
        const int batch = std::min(kProbe, chunk);
        ph_ws.grow((size_t)kProbe * 2 * n);
        res_counts.grow((size_t)2 * kProbe);
        for (size_t b0 = 0; b0 < idx.size(); b0 += (size_t)batch) {
            const int m = (int)std::min<size_t>((size_t)batch, idx.size() - b0);
            const int* id = idx.data() + b0;
            auto_flagged += m;
            const bool inplace = id[m - 1] - id[0] == m - 1;
            Wave sub;
            if (inplace) {
                sub = Wave{static_cast<const unsigned char*>(frames) + (size_t)id[0] * n * px, kind, m,
                           height_out + (long long)id[0] * n, mask ? mask + (long long)id[0] * mask_stride : nullptr, mask_stride};
            } else {
                g_frames.grow((size_t)kProbe * n * px);
                g_out.grow((size_t)kProbe * n);
                if (mask && mask_stride) g_mask.grow((size_t)kProbe * n);
                for (int j = 0; j < m; ++j) {
                    rt::d2d(g_frames.ptr + (size_t)j * n * px, static_cast<const unsigned char*>(frames) + (size_t)id[j] * n * px, n * px, s);
                    if (mask && mask_stride) rt::d2d(g_mask.ptr + (size_t)j * n, mask + (long long)id[j] * mask_stride, (size_t)n, s);
                }
                sub = Wave{g_frames.ptr, kind, m, g_out.ptr, mask ? (mask_stride ? g_mask.ptr : mask) : nullptr, mask_stride};
            }
            if (profiling) timer.begin_chunk(s, m);
            stage_front(sub, ph_ws.ptr, 0, nullptr, s);           // wrapped phases of the batch -> scratch
            rt::dmemset(res_counts.ptr, 0, sizeof(int) * 2 * kProbe, s);
            launch<ResidueCount>(H - 1, 2 * m, s, ResidueParams{ph_ws.ptr, res_counts.ptr, H, W});
            int counts[2 * kProbe];
            rt::d2h(counts, res_counts.ptr, sizeof(int) * (size_t)(2 * m), s);
            int q = 0;
            while (q < m) {
                if (!(counts[2 * q] || counts[2 * q + 1])) { ++q; continue; }
                int r = q;
                while (r < m && (counts[2 * r] || counts[2 * r + 1])) ++r;
                stage_guided(ph_ws.ptr, q, r - q, s);              // frames [q, r) of the batch
                stage_back(sub, q, r - q, 0, s);
                for (int k = q; k < r; ++k) {
                    if (!inplace)
                        rt::d2d(height_out + (long long)id[k] * n, g_out.ptr + (long long)k * n, sizeof(float) * (size_t)n, s);
                    if (phases)
                        rt::d2d(phases + (long long)id[k] * 2 * n, ph_ws.ptr + (long long)k * 2 * n, sizeof(float) * (size_t)(2 * n), s);
                    auto_guided.push_back(id[k]);
                }
                q = r;
            }
        }
    }

    void execute(const void* frames, int frame_kind, int n_frames, float* height_out, float* phases,
                 const uint8_t* mask, long long mask_stride, int unwrap, rt::stream_t s) {
        require_fused("fcd_execute (the fused float32 pipeline)");
        if (frame_kind < 0 || frame_kind > 2) rt::fail("frame dtype must be 0 (float32), 1 (uint8) or 2 (uint16)");
        if (unwrap < 0 || unwrap > 3) rt::fail("unwrap must be 0 (off), 1 (scan), 2 (reliability-guided) or 3 (auto)");
        const size_t px = frame_kind == 0 ? 4 : (frame_kind == 1 ? 1 : 2);
        if (!bound) throw std::logic_error("fcd_execute called before fcd_bind_reference");
        if (n_frames < 0) rt::fail("negative frame count");
        if (det == 0.0) rt::fail("carriers are collinear (singular 2x2 system)");
        const long long n = (long long)H * W;
        auto_flagged = 0;
        auto_guided.clear();
        if (unwrap == 3 && n_frames > 0) {
            frameflag.grow((size_t)n_frames);
            rt::dmemset(frameflag.ptr, 0, sizeof(int) * (size_t)n_frames, s);
        }
        auto wave_of = [&](int f0) {
            const int nf = std::min(chunk, n_frames - f0);
            return Wave{static_cast<const unsigned char*>(frames) + (size_t)f0 * n * px, frame_kind, nf,
                        height_out + (long long)f0 * n, mask ? mask + (long long)f0 * mask_stride : nullptr, mask_stride};
        };
        for (int f0 = 0; f0 < n_frames; f0 += chunk) {
            const Wave a = wave_of(f0);
            const int nf = a.nf;
            float* po = phases ? phases + (long long)f0 * 2 * n : nullptr;
            if (profiling) timer.begin_chunk(s, nf);
            // unwrap: 0 none, 1 row/column scan (exact where the wrapped phases have no residues),
            // 2 reliability-guided (Herraez et al., what skimage.restoration.unwrap_phase implements),
            // 3 scan, then 2 for the frames that hold residues
            if (unwrap == 2) {
                if (!po) {
                    ph_ws.grow((size_t)chunk * 2 * n);
                    po = ph_ws.ptr;
                }
                stage_front(a, po, 0, nullptr, s);
                stage_guided(po, 0, nf, s);
                stage_back(a, 0, nf, 0, s);
                continue;
            }
            const int scan = unwrap ? 1 : 0;
            stage_front(a, po, scan, unwrap == 3 ? frameflag.ptr + f0 : nullptr, s);
            stage_link(nf, po, scan, s);
            stage_back(a, 0, nf, scan, s);
        }
        if (unwrap == 3 && n_frames > 0) {
            // ONE read-back per call (4 bytes per frame), after every wave has been queued: on frames that cannot wrap
            // "auto" costs nothing else.
            std::vector<int> flags((size_t)n_frames), idx;
            rt::d2h(flags.data(), frameflag.ptr, sizeof(int) * (size_t)n_frames, s);
            for (int k = 0; k < n_frames; ++k)
                if (flags[(size_t)k]) idx.push_back(k);
            if (!idx.empty()) auto_second_look(idx, frames, frame_kind, height_out, phases, mask, mask_stride, s);
        }
    }

    // ------------------------------------------------------------------ reliability-guided unwrap ----
    // skimage.restoration.unwrap_phase (pyfcd/fcd.py:119): see fcd_unwrap.cuh.  Boruvka rounds until no edge
    // crosses two components; the maps of a wave share every launch.  One host read per round (the length of the
    // next edge list = termination test and grid size); the kernels take the exact lengths from device counters.
    int unwrap_wave_maps() const { return (int)std::max<long long>(1, std::min<long long>(16, (1LL << 26) / ((long long)H * W))); }
    rt::DevBuf<float> ph_ws;
    rt::DevBuf<double> u_rel, u_border;
    rt::DevBuf<po_t> u_po;
    rt::DevBuf<unsigned long long> u_bw, u_key[2];
    rt::DevBuf<unsigned> u_be, u_counters, u_id[2], u_ru[2], u_rv[2];
    rt::DevBuf<unsigned char> u_isroot;
    long long unwrap_rounds = 0;

    void unwrap_maps(const float* wrapped, int n_maps, float* out, rt::stream_t s) {
        const long long n = (long long)H * W;
        if (u_border.count != (size_t)(2 * W + 2 * H)) {
            // the same deterministic filler as oracle/unwrap_herraez.c (scikit-image draws these at random)
            std::vector<double> b((size_t)(2 * W + 2 * H));
            unsigned long long lcg = 0x9E3779B97F4A7C15ull;
            for (long long i = 0; i < n; ++i) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                const int r = (int)(i / W), c = (int)(i % W);
                if (r == 0 || c == 0 || r == H - 1 || c == W - 1) {
                    const int bi = r == 0 ? c : (r == H - 1 ? W + c : (c == 0 ? 2 * W + r : 2 * W + H + r));
                    b[(size_t)bi] = 9999999.0 + (double)(lcg >> 11) / 9007199254740992.0;
                }
            }
            u_border.upload(b, s);
        }
        const int wave = unwrap_wave_maps();
        const size_t cap = (size_t)wave * n;
        u_rel.grow(cap); u_po.grow(cap); u_bw.grow(cap); u_be.grow(cap); u_isroot.grow(cap); u_counters.grow(4);
        for (int k = 0; k < 2; ++k) {                         // a map has fewer than 2n edges
            u_key[k].grow(2 * cap); u_id[k].grow(2 * cap); u_ru[k].grow(2 * cap); u_rv[k].grow(2 * cap);
        }
        for (int m0 = 0; m0 < n_maps; m0 += wave) {
            const int nm = std::min(wave, n_maps - m0);
            const long long total = nm * n;
            const float* w = wrapped + m0 * n;
            launch<MstReliability>(blocks_for(total), 1, s, MstRelParams{w, u_border.ptr, u_rel.ptr, u_po.ptr, total, H, W});
            MstRoundParams base{w, u_rel.ptr, u_po.ptr, u_bw.ptr, u_be.ptr, EdgeList{}, EdgeList{},
                                u_counters.ptr, u_isroot.ptr, total, H, W, nullptr};
            auto list_of = [&](int k) { return EdgeList{u_key[k].ptr, u_id[k].ptr, u_ru[k].ptr, u_rv[k].ptr}; };
            // round 0: every pixel is a component (local minima, no unions), then the first list of cross edges
            rt::dmemset(u_counters.ptr, 0, 4 * sizeof(unsigned), s);
            launch<MstRound0>(blocks_for(total), 1, s, base);
            launch<MstFlatten>(blocks_for(total), 1, s, base);
            MstRoundParams bp = base;
            bp.out = list_of(0);
            launch<MstBuild>(blocks_for(total), 1, s, bp);
            unsigned c[4];
            rt::d2h(c, u_counters.ptr, sizeof(c), s);
            ++unwrap_rounds;
            long long count = c[1];
            int cur = 0;
            for (int round = 1; round < 64 && count > 0; ++round) {
                // counters: [3] = length of this round's list, the others start from zero
                const unsigned init[4] = {0u, 0u, 0u, (unsigned)count};
                rt::h2d(u_counters.ptr, init, sizeof(init), s);
                MstRoundParams rp = base;
                rp.in = list_of(cur); rp.out = list_of(cur ^ 1); rp.count = count;
                launch<MstSelect<0>>(blocks_for(count), 1, s, rp);
                launch<MstSelect<1>>(blocks_for(count), 1, s, rp);
                launch<MstHook>(blocks_for(count), 1, s, rp);
                launch<MstCompact>(blocks_for(count), 1, s, rp);
                rt::d2h(c, u_counters.ptr, sizeof(c), s);
                ++unwrap_rounds;
                if ((long long)c[1] >= count) rt::fail("unwrap: a round merged nothing");
                count = c[1];
                cur ^= 1;
            }
            launch<MstFlattenRoots>(blocks_for(total), 1, s, base);
            MstRoundParams fp = base;
            fp.outp = out + m0 * n;
            launch<MstApply>(blocks_for(total), 1, s, fp);
        }
    }

    // ------------------------------------------------------------------ temporal analysis ----
    // analyze.block_amplitude (pydata/analyze.py:542-641): see fcd_temporal.cuh
    template <int L> struct TuneT { static constexpr int G = L >= 4096 ? 2 : (L >= 2048 ? 4 : 8); };
    int t_len = 0, t_n = 0;                 // cached tables: transform length, number of frames
    rt::DevBuf<cf> t_tw, t_twn, t_ws;
    const cf* t_chirp_ptr = nullptr;
    const cf* t_bspec_ptr = nullptr;
    int t_twn_n = 0;
    rt::DevBuf<double> t_mean, t_twd;
    rt::DevBuf<int> t_valid, t_bins;

    static int temporal_length(int n_frames) {        // single-level transform length; 0: needs the two-level path
        if (n_frames >= 64 && n_frames <= 4096 && (n_frames & (n_frames - 1)) == 0) return n_frames;
        int l = 64;
        while (l < 2 * n_frames - 1) l *= 2;
        return l <= 4096 ? l : 0;
    }
    static void host_fft(std::vector<std::complex<double>>& a) {   // in place, forward, power of two
        const size_t n = a.size();
        for (size_t i = 1, j = 0; i < n; ++i) {
            size_t bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) std::swap(a[i], a[j]);
        }
        for (size_t len = 2; len <= n; len <<= 1) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * M_PI * (double)k / (double)len;
                const std::complex<double> w(std::cos(ang), std::sin(ang));
                for (size_t i = k; i < n; i += len) {
                    const std::complex<double> u = a[i], v = a[i + len / 2] * w;
                    a[i] = u + v;
                    a[i + len / 2] = u - v;
                }
            }
        }
    }
    // engine twiddles + Bluestein tables of one transform length
    struct BlueTables {
        int n = 0, len = 0;
        rt::DevBuf<cf> tw, chirp, bspec;
    };
    static int bluestein_length(int n) {                 // power of two >= 2n - 1 (>= 64); 0: longer than the engine
        int l = 64;
        while (l < 2 * n - 1) l *= 2;
        return l <= 4096 ? l : 0;
    }
    BlueTables t_single, t_lvl[2];
    void build_blue(BlueTables& t, int n, int len, rt::stream_t s) {
        if (t.n == n && t.len == len) return;
        FCD_DISPATCH_L(len, { t.tw.upload(Fft<L, -1, float>::make_table(), s); })
        // Bluestein: X[k] = conj(c[k]) * sum_t (x[t] conj(c[t])) c[k - t],  c[m] = exp(i pi m^2 / N)
        std::vector<std::complex<double>> c((size_t)n), b((size_t)len, 0.0);
        for (int m = 0; m < n; ++m) {
            const long long q = ((long long)m * m) % (2LL * n);
            const double ang = M_PI * (double)q / (double)n;
            c[m] = {std::cos(ang), std::sin(ang)};
            b[m] = c[m];
            if (m) b[len - m] = c[m];
        }
        host_fft(b);
        std::vector<cf> hc((size_t)n), hb((size_t)len);
        for (int m = 0; m < n; ++m) hc[m] = mk<float>((float)c[m].real(), (float)c[m].imag());
        for (int m = 0; m < len; ++m) hb[m] = mk<float>((float)(b[m].real() / len), (float)(b[m].imag() / len));
        t.chirp.upload(hc, s);
        t.bspec.upload(hb, s);
        t.n = n; t.len = len;
    }
    void temporal_tables(int n_frames, rt::stream_t s) {
        const int len = temporal_length(n_frames);
        if (!len) rt::fail("temporal spectrum: unsupported single-level length");
        if (len == t_len && n_frames == t_n) return;
        FCD_DISPATCH_L(len, { t_tw.upload(Fft<L, -1, float>::make_table(), s); })
        if (len != n_frames) {
            build_blue(t_single, n_frames, len, s);
            t_chirp_ptr = t_single.chirp.ptr; t_bspec_ptr = t_single.bspec.ptr;
        }
        t_len = len; t_n = n_frames;
    }
    // N = N1 * N2 with both factors within reach of one chirp convolution (N1 >= N2, as balanced as possible)
    static bool temporal_split(int n_frames, int& n1, int& n2) {
        for (int d = (int)std::sqrt((double)n_frames); d >= 2; --d)
            if (n_frames % d == 0 && n_frames / d <= 2048) { n1 = n_frames / d; n2 = d; return true; }
        return false;
    }
    static bool temporal_supported(int n_frames) {
        int a, b;
        return n_frames >= 2 && (temporal_length(n_frames) != 0 || temporal_split(n_frames, a, b));
    }
    static int temporal_npos(int n_frames) { return n_frames % 2 == 0 ? n_frames / 2 : (n_frames + 1) / 2; }   // fftfreq >= 0

    static void temporal_geometry(int rows, int cols, int bs, int brows, int bcols) {
        if (rows < 1 || cols < 1 || bs < 1 || brows < 1 || bcols < 1) rt::fail("temporal analysis: bad geometry");
        if ((long long)bs * brows > rows || (long long)bs * bcols > cols) rt::fail("temporal analysis: blocks do not fit the maps");
    }

    void temporal_mean_spectrum(const float* maps, int n_frames, int rows, int cols, const float* first, float zero,
                                int bs, int brows, int bcols, int force_n1, double* mean_host, int* valid_host,
                                rt::stream_t s) {
        temporal_geometry(rows, cols, bs, brows, bcols);
        if (n_frames < 2) rt::fail("temporal spectrum needs at least two frames");
        const int nblk = brows * bcols, npos = temporal_npos(n_frames);
        t_mean.alloc((size_t)nblk * npos);
        t_valid.alloc((size_t)nblk);
        rt::dmemset(t_mean.ptr, 0, sizeof(double) * (size_t)nblk * npos, s);
        rt::dmemset(t_valid.ptr, 0, sizeof(int) * (size_t)nblk, s);
        const long long segs = (long long)bs * brows * bcols;
        launch<BlockValidCount>(blocks_for(segs), 1, s, BlockValidParams{first, t_valid.ptr, rows, cols, bs, brows, bcols, segs});
        int n1 = 0, n2 = 0;
        if (force_n1 > 0) {                                    // tests: exercise the two-level path on short series
            if (n_frames % force_n1 != 0) rt::fail("temporal spectrum: n1 does not divide the number of frames");
            n1 = force_n1; n2 = n_frames / force_n1;
        }
        if (force_n1 <= 0 && temporal_length(n_frames) != 0) {
            temporal_tables(n_frames, s);
            TemporalSpecParams p{maps, first, t_tw.ptr, t_chirp_ptr, t_bspec_ptr, t_mean.ptr, zero, n_frames, npos, rows, cols, bs, brows, bcols};
            const bool blue = t_len != n_frames;
            FCD_DISPATCH_L(t_len, {
                constexpr int G = TuneT<L>::G;
                if (bs % G != 0) rt::fail("temporal spectrum: the block size must be a multiple of 8");
                if (blue) launch<TemporalSpectrum<L, G, true>>(bs * (bs / G), nblk, s, p);
                else launch<TemporalSpectrum<L, G, false>>(bs * (bs / G), nblk, s, p);
            })
        } else {
            if (n1 == 0 && !temporal_split(n_frames, n1, n2))
                rt::fail("temporal spectrum: the number of frames has no factorisation into two lengths <= 2048 "
                         "(drop a frame or pass f0)");
            const int l1 = bluestein_length(n1), l2 = bluestein_length(n2);
            if (!l1 || !l2) rt::fail("temporal spectrum: factor too long");
            build_blue(t_lvl[0], n1, l1, s);
            build_blue(t_lvl[1], n2, l2, s);
            if (t_twn_n != n_frames) {
                std::vector<cf> h((size_t)n_frames);
                for (int m = 0; m < n_frames; ++m) {
                    const double ang = -2.0 * M_PI * (double)m / (double)n_frames;
                    h[m] = mk<float>((float)std::cos(ang), (float)std::sin(ang));
                }
                t_twn.upload(h, s);
                t_twn_n = n_frames;
            }
            // chunks of spatial blocks whose [N][pixels] complex workspace stays under 8 GB
            const long long per_block = (long long)n_frames * bs * bs * (long long)sizeof(cf);
            const int chunk_blocks = (int)std::max<long long>(1, std::min<long long>(nblk, (8LL << 30) / per_block));
            t_ws.alloc((size_t)n_frames * chunk_blocks * bs * bs);
            for (int b0 = 0; b0 < nblk; b0 += chunk_blocks) {
                const int nb = std::min(chunk_blocks, nblk - b0);
                TemporalTwoLevelParams p{maps, first, t_ws.ptr, nullptr, nullptr, nullptr, t_twn.ptr, t_mean.ptr, zero,
                                         n_frames, n1, n2, npos, rows, cols, bs, brows, bcols, b0, nb * bs * bs};
                p.tw = t_lvl[0].tw.ptr; p.chirp = t_lvl[0].chirp.ptr; p.bspec = t_lvl[0].bspec.ptr;
                FCD_DISPATCH_L(l1, {
                    constexpr int G = TuneT<L>::G;
                    if (bs % G != 0) rt::fail("temporal spectrum: the block size must be a multiple of 8");
                    launch<TemporalTwoLevel<L, G, 0>>(bs * (bs / G) * n2, nb, s, p);
                })
                p.tw = t_lvl[1].tw.ptr; p.chirp = t_lvl[1].chirp.ptr; p.bspec = t_lvl[1].bspec.ptr;
                FCD_DISPATCH_L(l2, {
                    constexpr int G = TuneT<L>::G;
                    if (bs % G != 0) rt::fail("temporal spectrum: the block size must be a multiple of 8");
                    launch<TemporalTwoLevel<L, G, 1>>(bs * (bs / G), nb * n1, s, p);
                })
            }
        }
        rt::d2h(mean_host, t_mean.ptr, sizeof(double) * (size_t)nblk * npos, s);
        rt::d2h(valid_host, t_valid.ptr, sizeof(int) * (size_t)nblk, s);
        for (int b = 0; b < nblk; ++b)                                   // np.nanmean over the block's valid pixels
            for (int k = 0; k < npos; ++k)
                mean_host[(size_t)b * npos + k] = valid_host[b] ? mean_host[(size_t)b * npos + k] / valid_host[b] : nan_f64();
    }

    void temporal_accumulate(const float* maps, int n_chunk, int t0, int n_total, int rows, int cols, float zero,
                             int bs, int brows, int bcols, const int* bins, int n_bins, double* acc, int init,
                             rt::stream_t s) {
        temporal_geometry(rows, cols, bs, brows, bcols);
        if (n_bins < 1 || n_bins > kMaxHarmonicBins) rt::fail("temporal harmonics: 1..8 bins per block");
        if (n_chunk < 0 || t0 < 0 || n_total < 1 || t0 + n_chunk > n_total) rt::fail("temporal harmonics: bad frame range");
        const int nblk = brows * bcols;
        std::vector<int> hb(bins, bins + (size_t)nblk * n_bins);
        t_bins.upload(hb, s);
        const long long ntw = (long long)std::max(n_chunk, 1) * nblk * n_bins;
        t_twd.alloc((size_t)ntw * 2);
        launch<HarmonicTwiddles>(blocks_for(ntw), 1, s, HarmonicTwParams{t_bins.ptr, t_twd.ptr, n_chunk, nblk, n_bins, t0, n_total, (long long)n_chunk * nblk * n_bins});
        const long long total = (long long)rows * cols;
        const HarmonicAccParams hp{maps, t_twd.ptr, acc, zero, n_chunk, n_bins, rows, cols, bs, brows, bcols, init, total};
        switch (n_bins) {
            case 1: launch<HarmonicAccumulate<1>>(blocks_for(total), 1, s, hp); break;
            case 2: launch<HarmonicAccumulate<2>>(blocks_for(total), 1, s, hp); break;
            case 3: launch<HarmonicAccumulate<3>>(blocks_for(total), 1, s, hp); break;
            case 4: launch<HarmonicAccumulate<4>>(blocks_for(total), 1, s, hp); break;
            case 5: launch<HarmonicAccumulate<5>>(blocks_for(total), 1, s, hp); break;
            case 6: launch<HarmonicAccumulate<6>>(blocks_for(total), 1, s, hp); break;
            case 7: launch<HarmonicAccumulate<7>>(blocks_for(total), 1, s, hp); break;
            default: launch<HarmonicAccumulate<8>>(blocks_for(total), 1, s, hp); break;
        }
    }

    void temporal_finalize(const double* acc, int n_bins, int n_total, int rows, int cols, const float* first,
                           double* amps, double* phases, rt::stream_t s) {
        if (n_bins < 1 || n_bins > kMaxHarmonicBins || n_total < 1 || rows < 1 || cols < 1)
            rt::fail("temporal harmonics: bad arguments (1..8 bins, positive frame count and map size)");
        const long long total = (long long)rows * cols;
        launch<HarmonicFinalize>(blocks_for(total), 1, s, HarmonicFinParams{acc, first, amps, phases, n_bins, n_total, total});
    }

    // ------------------------------------------------------------------ structure mask / centre ----
    // analyze.mask (pydata/analyze.py:43-100) and analyze.center (pydata/analyze.py:104-140)
    // frames per wave: several of these kernels run one thread per image line (the reference's running sums
    // and raster labelling are sequential along a line), so the wave has to be wide to fill the device;
    // capped at 2^28 pixels of workspace rows (~50 bytes per pixel)
    int mask_chunk() const { return (int)std::max<long long>(1, std::min<long long>(64, (1LL << 28) / ((long long)H * W))); }
    rt::DevBuf<float> m_t0, m_smooth, m_ps0, m_ps1;
    rt::DevBuf<int> m_L, m_area, m_bbox, m_centers;
    rt::DevBuf<unsigned> m_bits;
    rt::DevBuf<unsigned long long> m_sums, m_best;

    static int blocks_for(long long total) { return (int)((total + 255) / 256); }

    void mask_workspace(bool with_stats) {
        const size_t n = (size_t)H * W, c = (size_t)mask_chunk();
        m_L.alloc(c * n);
        m_bits.alloc(c * n / 32);
        m_area.alloc(c * n);
        m_best.alloc(c);
        if (with_stats) {
            m_bbox.alloc(4 * c * n);
            m_sums.alloc(2 * c * n);
            m_centers.alloc(2 * c);
        } else {
            m_t0.alloc(c * n);
            m_smooth.alloc(c * n);
            m_ps0.alloc(c * n / 128);
            m_ps1.alloc(c * n / 256 + 1);
        }
    }

    void structure_mask(const float* frames, int n_frames, int smoothed, uint8_t* mask_out, rt::stream_t s) {
        require_fused("fcd_structure_mask");
        if (smoothed < 1 || smoothed > std::min(H, W)) rt::fail("smoothed must be in [1, min(rows, cols)]");
        mask_workspace(false);
        const long long n = (long long)H * W;
        for (int f0 = 0; f0 < n_frames; f0 += mask_chunk()) {
            const int nf = std::min(mask_chunk(), n_frames - f0);
            const long long total = nf * n;
            const float* in = frames + f0 * n;
            launch<BoxLines>(blocks_for((long long)nf * W), 1, s, BoxLinesParams{in, m_t0.ptr, H, W, smoothed, 0, (long long)nf * W});
            if (smoothed <= 32)     // warp-tiled rows: coalesced, same arithmetic
                launch<BoxRowsWarp>((int)(((long long)nf * H / 32 + 3) / 4), 1, s, BoxLinesParams{m_t0.ptr, m_smooth.ptr, H, W, smoothed, 1, (long long)nf * H});
            else
                launch<BoxLines>(blocks_for((long long)nf * H), 1, s, BoxLinesParams{m_t0.ptr, m_smooth.ptr, H, W, smoothed, 1, (long long)nf * H});
            // np.mean(smooth): float32 pairwise sum
            long long m = n / 128;
            launch<PairBlockSum>(blocks_for(8 * nf * m), 1, s, PairBlockParams{m_smooth.ptr, m_ps0.ptr, nf * m});
            float* a = m_ps0.ptr;
            float* b = m_ps1.ptr;
            while (m > 1) {
                m /= 2;
                launch<PairTree>(blocks_for(nf * m), 1, s, PairTreeParams{a, b, nf * m});
                std::swap(a, b);
            }
            launch<LabelInit>(blocks_for(32LL * nf * H), 1, s, LabelInitParams{m_smooth.ptr, a, nullptr, m_L.ptr, m_bits.ptr, (long long)nf * H, H, W, 0});
            launch<LabelMerge>(blocks_for(total / 32), 1, s, LabelMergeParams{m_L.ptr, m_bits.ptr, total / 32, H, W});
            rt::dmemset(m_best.ptr, 0, sizeof(unsigned long long) * (size_t)nf, s);
            RegionStats st{m_area.ptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            launch<RootStatsInit>(blocks_for(total / 32), 1, s, RootStatsInitParams{m_L.ptr, m_bits.ptr, st, total / 32, H, W, 0});
            launch<LabelFlatten>(blocks_for(32LL * nf * H), 1, s, LabelFlattenParams{m_L.ptr, m_bits.ptr, st, (long long)nf * H, H, W, 0});
            launch<LargestRegion>(blocks_for(total / 32), 1, s, LargestParams{m_L.ptr, m_bits.ptr, st, m_best.ptr, total / 32, H, W, 0});
            launch<MaskOut>(blocks_for(total / 32), 1, s, MaskOutParams{m_L.ptr, m_bits.ptr, m_best.ptr, mask_out + f0 * n, total / 32, H, W});
        }
    }

    void mask_center(const uint8_t* mask, int n_frames, int* centers_host, rt::stream_t s) {
        require_fused("fcd_mask_center");
        mask_workspace(true);
        const long long n = (long long)H * W;
        for (int f0 = 0; f0 < n_frames; f0 += mask_chunk()) {
            const int nf = std::min(mask_chunk(), n_frames - f0);
            const long long total = nf * n;
            launch<LabelInit>(blocks_for(32LL * nf * H), 1, s, LabelInitParams{nullptr, nullptr, mask + f0 * n, m_L.ptr, m_bits.ptr, (long long)nf * H, H, W, 1});
            launch<LabelMerge>(blocks_for(total / 32), 1, s, LabelMergeParams{m_L.ptr, m_bits.ptr, total / 32, H, W});
            rt::dmemset(m_best.ptr, 0, sizeof(unsigned long long) * (size_t)nf, s);
            int* minr = m_bbox.ptr; int* maxr = minr + total; int* minc = maxr + total; int* maxc = minc + total;
            RegionStats st{m_area.ptr, minr, maxr, minc, maxc, m_sums.ptr, m_sums.ptr + total};
            launch<RootStatsInit>(blocks_for(total / 32), 1, s, RootStatsInitParams{m_L.ptr, m_bits.ptr, st, total / 32, H, W, 1});
            launch<LabelFlatten>(blocks_for(32LL * nf * H), 1, s, LabelFlattenParams{m_L.ptr, m_bits.ptr, st, (long long)nf * H, H, W, 1});
            launch<LargestRegion>(blocks_for(total / 32), 1, s, LargestParams{m_L.ptr, m_bits.ptr, st, m_best.ptr, total / 32, H, W, 1});
            launch<CenterOut>(blocks_for(nf), 1, s, CenterOutParams{m_best.ptr, st, m_centers.ptr, nf, (int)n});
            rt::d2h(centers_host + 2 * f0, m_centers.ptr, sizeof(int) * 2 * (size_t)nf, s);
        }
    }

    // float32 copy of the reference for mask substitution (analyze.py:231), made on demand
    rt::DevBuf<float> ref_f32;
    bool ref_f32_valid = false;
    const float* reference_f32() const { return ref_f32_valid ? ref_f32.ptr : nullptr; }
};

}  // namespace fcd

// =========================================================================================
// C ABI
// =========================================================================================
struct fcd_plan {
    fcd::PlanImpl impl;
};

static thread_local std::string g_fcd_error;

template <class Fn>
static int fcd_guard(Fn&& fn) {
    try {
        fn();
        return FCD_OK;
    } catch (const std::domain_error& e) {
        g_fcd_error = e.what();
        return FCD_ERR_NOPEAKS;
    } catch (const std::logic_error& e) {
        g_fcd_error = e.what();
        return FCD_ERR_STATE;
    } catch (const fcd::rt::cuda_error& e) {
        g_fcd_error = e.what();
        return FCD_ERR_RUNTIME;
    } catch (const std::runtime_error& e) {      // rt::fail: bad argument / unsupported request
        g_fcd_error = e.what();
        return FCD_ERR_INVALID;
    } catch (const std::exception& e) {
        g_fcd_error = e.what();
        return FCD_ERR_RUNTIME;
    }
}

extern "C" {

const char* fcd_last_error(void) { return g_fcd_error.c_str(); }

int fcd_plan_create(int rows, int cols, int frames_per_launch, fcd_plan** out) {
    if (!out) { g_fcd_error = "null output pointer"; return FCD_ERR_INVALID; }
    *out = nullptr;
    return fcd_guard([&] {
        fcd_plan* p = new fcd_plan();
        try {
            p->impl.init(rows, cols, frames_per_launch);
        } catch (...) {
            delete p;
            throw;
        }
        *out = p;
    });
}

int fcd_plan_destroy(fcd_plan* plan) {
    delete plan;
    return FCD_OK;
}

int fcd_highpass_spectrum(fcd_plan* plan, const void* image_dev, int image_is_f64, double* spectrum_dev,
                          double* max_out, void* stream) {
    if (!plan || !image_dev) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        const double mx = plan->impl.highpass_spectrum(image_dev, image_is_f64, spectrum_dev, stream);
        if (max_out) *max_out = mx;
    });
}

int fcd_peak_locations(fcd_plan* plan, const double* image_dev, double threshold, int max_peaks, int* rc_out,
                       int* count_out, void* stream) {
    if (!plan || !image_dev || !rc_out || !count_out || max_peaks < 0) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        auto v = plan->impl.peak_locations(image_dev, threshold, max_peaks, stream);
        *count_out = (int)v.size();
        for (size_t i = 0; i < v.size(); ++i) { rc_out[2 * i] = v[i][0]; rc_out[2 * i + 1] = v[i][1]; }
    });
}

int fcd_find_peaks(fcd_plan* plan, const void* image_dev, int image_is_f64, int peaks_out[4], void* stream) {
    if (!plan || !image_dev || !peaks_out) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.find_peaks(image_dev, image_is_f64, peaks_out, stream); });
}

int fcd_bind_reference(fcd_plan* plan, const void* reference_dev, int reference_is_f64, const int peaks[4],
                       double radius, double calibration_factor, double height, void* stream) {
    if (!plan || !reference_dev || !peaks) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        auto& im = plan->impl;
        im.bind(reference_dev, reference_is_f64, peaks, radius, calibration_factor, height, stream);
        // float32 copy of the reference for the masked workflow
        const long long n = (long long)im.H * im.W;
        im.ref_f32.alloc((size_t)n);
        if (!reference_is_f64) {
            fcd::rt::d2d(im.ref_f32.ptr, reference_dev, n * sizeof(float), stream);
            fcd::rt::sync(stream);
            im.ref_f32_valid = true;
        } else {
            const int nb = im.elem_blocks(n);
            im.launch<fcd::NarrowF64>(nb, 1, stream, fcd::NarrowParams{static_cast<const double*>(reference_dev), im.ref_f32.ptr, n, nb});
            fcd::rt::sync(stream);
            im.ref_f32_valid = true;
        }
    });
}

int fcd_execute(fcd_plan* plan, const float* frames_dev, int n_frames, float* height_dev, float* phases_dev,
                const uint8_t* mask_dev, long long mask_stride, int unwrap, void* stream) {
    if (!plan || (n_frames > 0 && (!frames_dev || !height_dev))) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.execute(frames_dev, 0, n_frames, height_dev, phases_dev, mask_dev, mask_stride, unwrap, stream);
    });
}

int fcd_execute_typed(fcd_plan* plan, const void* frames_dev, int frame_dtype, int n_frames, float* height_dev,
                      float* phases_dev, const uint8_t* mask_dev, long long mask_stride, int unwrap, void* stream) {
    if (!plan || (n_frames > 0 && (!frames_dev || !height_dev))) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.execute(frames_dev, frame_dtype, n_frames, height_dev, phases_dev, mask_dev, mask_stride, unwrap, stream);
    });
}

int fcd_set_height(fcd_plan* plan, double height) {
    if (!plan) { g_fcd_error = "null plan"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.set_height(height); });
}

int fcd_count_residues(fcd_plan* plan, const float* phases_dev, int n_maps, int* counts_out, void* stream) {
    if (!plan || !phases_dev || !counts_out || n_maps < 0) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        auto& im = plan->impl;
        fcd::rt::DevBuf<int> counts;
        counts.alloc((size_t)std::max(n_maps, 1));
        fcd::rt::dmemset(counts.ptr, 0, sizeof(int) * (size_t)std::max(n_maps, 1), stream);
        if (n_maps > 0) {
            im.launch<fcd::ResidueCount>(im.H - 1, n_maps, stream, fcd::ResidueParams{phases_dev, counts.ptr, im.H, im.W});
            fcd::rt::d2h(counts_out, counts.ptr, sizeof(int) * (size_t)n_maps, stream);
        }
    });
}

int fcd_get_carrier_mask(fcd_plan* plan, int carrier, uint8_t* mask_dev, void* stream) {
    if (!plan || !mask_dev || carrier < 0 || carrier > 1) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        if (!plan->impl.bound) throw std::logic_error("no reference bound");
        plan->impl.masked_inverse(carrier, stream, mask_dev);
    });
}

int fcd_get_carrier_ccsgn(fcd_plan* plan, int carrier, void* ccsgn_dev, int as_c128, void* stream) {
    if (!plan || !ccsgn_dev || carrier < 0 || carrier > 1) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        auto& im = plan->impl;
        if (!im.bound || !im.spec_valid) throw std::logic_error("no reference bound");
        const long long n = (long long)im.H * im.W;
        if (!as_c128) {
            fcd::rt::d2d(ccsgn_dev, im.ccsgn.ptr + (size_t)carrier * n, n * sizeof(fcd::cf), stream);
            return;
        }
        im.masked_inverse(carrier, stream, nullptr);
        const int nb = im.elem_blocks(n);
        im.launch<fcd::CcsgnStore>(nb, 1, stream,
                                   fcd::CcsgnStoreParams{im.tmp.ptr, nullptr, static_cast<fcd::cd*>(ccsgn_dev), nullptr, n, nb});
    });
}

int fcd_fft2_c128(fcd_plan* plan, const void* in_dev, void* out_dev, int direction, void* stream) {
    if (!plan || !in_dev || !out_dev || (direction != 1 && direction != -1)) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.fft2_d(in_dev, 0, 0.0, static_cast<fcd::cd*>(out_dev), direction, stream); });
}

int fcd_set_profiling(fcd_plan* plan, int enable) {
    if (!plan) { g_fcd_error = "null plan"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.profiling = enable != 0;
        plan->impl.timer.reset();
    });
}

int fcd_stage_times(fcd_plan* plan, double ms_out[7], long long launches_out[7], long long frames_out[7]) {
    if (!plan || !ms_out || !launches_out || !frames_out) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.timer.collect(ms_out, launches_out, frames_out); });
}

int fcd_unwrap_phase(fcd_plan* plan, const float* wrapped_dev, int n_maps, float* unwrapped_dev, void* stream) {
    if (!plan || n_maps < 0 || (n_maps > 0 && (!wrapped_dev || !unwrapped_dev))) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.unwrap_maps(wrapped_dev, n_maps, unwrapped_dev, stream);
        fcd::rt::sync(stream);
    });
}

int fcd_temporal_mean_spectrum(fcd_plan* plan, const float* maps_dev, int n_frames, int rows, int cols,
                               const float* first_map_dev, float zero, int block_size, int block_rows, int block_cols,
                               double* mean_out, int* valid_out, void* stream) {
    if (!plan || !maps_dev || !mean_out || !valid_out) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.temporal_mean_spectrum(maps_dev, n_frames, rows, cols, first_map_dev, zero, block_size, block_rows,
                                          block_cols, 0, mean_out, valid_out, stream);
    });
}

int fcd_temporal_mean_spectrum_split(fcd_plan* plan, const float* maps_dev, int n_frames, int rows, int cols,
                                     const float* first_map_dev, float zero, int block_size, int block_rows, int block_cols,
                                     int n1, double* mean_out, int* valid_out, void* stream) {
    if (!plan || !maps_dev || !mean_out || !valid_out || n1 < 1) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.temporal_mean_spectrum(maps_dev, n_frames, rows, cols, first_map_dev, zero, block_size, block_rows,
                                          block_cols, n1, mean_out, valid_out, stream);
    });
}

int fcd_temporal_frames_supported(int n_frames) { return fcd::PlanImpl::temporal_supported(n_frames) ? 1 : 0; }

int fcd_temporal_accumulate(fcd_plan* plan, const float* maps_dev, int n_chunk, int t0, int n_total, int rows, int cols,
                            float zero, int block_size, int block_rows, int block_cols, const int* bins, int n_bins,
                            double* acc_dev, int init, void* stream) {
    if (!plan || (n_chunk > 0 && !maps_dev) || !bins || !acc_dev) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.temporal_accumulate(maps_dev, n_chunk, t0, n_total, rows, cols, zero, block_size, block_rows, block_cols,
                                       bins, n_bins, acc_dev, init, stream);
    });
}

int fcd_temporal_finalize(fcd_plan* plan, const double* acc_dev, int n_bins, int n_total, int rows, int cols,
                          const float* first_map_dev, double* amps_dev, double* phases_dev, void* stream) {
    if (!plan || !acc_dev || !amps_dev || !phases_dev) { g_fcd_error = "null argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] {
        plan->impl.temporal_finalize(acc_dev, n_bins, n_total, rows, cols, first_map_dev, amps_dev, phases_dev, stream);
    });
}

int fcd_structure_mask(fcd_plan* plan, const float* frames_dev, int n_frames, int smoothed, uint8_t* mask_dev,
                       void* stream) {
    if (!plan || n_frames < 0 || (n_frames > 0 && (!frames_dev || !mask_dev))) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.structure_mask(frames_dev, n_frames, smoothed, mask_dev, stream); });
}

int fcd_mask_center(fcd_plan* plan, const uint8_t* mask_dev, int n_frames, int* centers_out, void* stream) {
    if (!plan || n_frames < 0 || (n_frames > 0 && (!mask_dev || !centers_out))) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    return fcd_guard([&] { plan->impl.mask_center(mask_dev, n_frames, centers_out, stream); });
}

int fcd_last_auto(const fcd_plan* plan, long long* flagged_out, int* guided_count_out, int* guided_frames_out, int capacity) {
    if (!plan || capacity < 0) { g_fcd_error = "bad argument"; return FCD_ERR_INVALID; }
    const auto& im = plan->impl;
    if (flagged_out) *flagged_out = im.auto_flagged;
    if (guided_count_out) *guided_count_out = (int)im.auto_guided.size();
    if (guided_frames_out)
        for (int i = 0; i < capacity && i < (int)im.auto_guided.size(); ++i) guided_frames_out[i] = im.auto_guided[(size_t)i];
    return FCD_OK;
}

long long fcd_launch_count(const fcd_plan* plan) { return plan ? plan->impl.launches : 0; }
int fcd_band_columns(const fcd_plan* plan) { return plan ? plan->impl.ncp : 0; }
int fcd_plan_is_fused(const fcd_plan* plan) { return plan && !plan->impl.generic ? 1 : 0; }

}  // extern "C"
