// cufft_compare.cu -- BENCHMARK UTILITY ONLY (libfcd_cufft_compare.so): the transforms of every stage of the FCD
// pipeline on cuFFT, timed with CUDA events, so that each hand-written kernel K1..K5 can be put next to the cuFFT
// calls a library user would make for the same stage (`north_star`: "each is also compared with a cuFFT-based
// version").  Only the FFTs are timed here -- none of the work the hand-written kernels fuse around them (band
// extraction and transposition, disk mask, ccsgn product, atan2, unwrap, 2x2 solve, k-space coefficients) -- so the
// sum over the stages is a FLOOR for any pipeline built on these cuFFT calls.  The product (libfcd_b200.so) does
// not link cuFFT and never loads this library.
//
// Stage shapes (frames F, image H x W, ncp band columns per carrier), matching csrc/fcd_kernels.cuh:
//   K1 row forward       R2C, F*H transforms of length W
//   K2 column band-pass  C2C forward + C2C inverse, F*2*ncp contiguous transforms of length H
//   K3 row demodulation  C2C inverse, F*2*H transforms of length W (both carriers) + C2C forward, F*H of length W
//   K4 column integrate  C2C forward over the W columns + C2C inverse over W/2+1 columns, length H, stride W
//   K5 row inverse       C2R, F*H transforms of length W
#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdio>
#include <string>

namespace {

thread_local std::string g_err;

struct Fail {
    std::string what;
};
void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Fail{std::string(what) + ": " + cudaGetErrorString(e)};
}
void ckf(cufftResult r, const char* what) {
    if (r != CUFFT_SUCCESS) throw Fail{std::string(what) + ": cufft error " + std::to_string((int)r)};
}

struct Plans {
    cufftHandle h[8];
    int n = 0;
    cufftHandle make_many(int len, int istride, int idist, int ostride, int odist, cufftType type, int batch, cudaStream_t s) {
        cufftHandle p;
        int dims[1] = {len};
        int inembed[1] = {len}, onembed[1] = {len};
        ckf(cufftPlanMany(&p, 1, dims, inembed, istride, idist, onembed, ostride, odist, type, batch), "cufftPlanMany");
        ckf(cufftSetStream(p, s), "cufftSetStream");
        h[n++] = p;
        return p;
    }
    ~Plans() {
        for (int i = 0; i < n; ++i) cufftDestroy(h[i]);
    }
};

struct Buf {
    void* p = nullptr;
    explicit Buf(size_t bytes) { ck(cudaMalloc(&p, bytes), "cudaMalloc"); ck(cudaMemset(p, 0, bytes), "cudaMemset"); }
    ~Buf() { cudaFree(p); }
};

template <class Fn>
double time_us_per_frame(Fn&& fn, int frames, int reps, cudaStream_t s) {
    cudaEvent_t a, b;
    ck(cudaEventCreate(&a), "event");
    ck(cudaEventCreate(&b), "event");
    fn();                                       // warm-up (plan workspaces, clocks)
    ck(cudaStreamSynchronize(s), "sync");
    ck(cudaEventRecord(a, s), "record");
    for (int r = 0; r < reps; ++r) fn();
    ck(cudaEventRecord(b, s), "record");
    ck(cudaEventSynchronize(b), "sync");
    float ms = 0.f;
    ck(cudaEventElapsedTime(&ms, a, b), "elapsed");
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return (double)ms * 1e3 / ((double)reps * frames);
}

}  // namespace

extern "C" {

const char* fcdcmp_last_error(void) { return g_err.c_str(); }

// us_out[6]: microseconds per frame of the cuFFT calls of stages K1..K5 ([3]: K4 on the row-major plane, stride W)
// and [5]: K4 on contiguous columns.  Returns 0 on success.
int fcdcmp_time_stages(int H, int W, int ncp, int frames, int reps, double us_out[6], void* stream) {
    try {
        if (H < 2 || W < 2 || ncp < 1 || frames < 1 || reps < 1 || !us_out) throw Fail{"bad argument"};
        cudaStream_t s = (cudaStream_t)stream;
        const size_t P = (size_t)H * W, F = (size_t)frames;
        const int Wh = W / 2 + 1;
        Plans pl;
        // work buffers sized for the largest stage: F*2*P complex (K3's two carriers)
        Buf a(F * 2 * P * sizeof(cufftComplex)), b(F * 2 * P * sizeof(cufftComplex));
        cufftComplex* ca = static_cast<cufftComplex*>(a.p);
        cufftComplex* cb = static_cast<cufftComplex*>(b.p);
        float* ra = static_cast<float*>(a.p);
        float* rb = static_cast<float*>(b.p);

        cufftHandle k1 = pl.make_many(W, 1, W, 1, Wh, CUFFT_R2C, (int)(F * H), s);
        us_out[0] = time_us_per_frame([&] { ckf(cufftExecR2C(k1, ra, cb), "K1 R2C"); }, frames, reps, s);

        cufftHandle k2 = pl.make_many(H, 1, H, 1, H, CUFFT_C2C, (int)(F * 2 * ncp), s);
        us_out[1] = time_us_per_frame([&] {
            ckf(cufftExecC2C(k2, ca, cb, CUFFT_FORWARD), "K2 fwd");
            ckf(cufftExecC2C(k2, cb, ca, CUFFT_INVERSE), "K2 inv");
        }, frames, reps, s);

        cufftHandle k3a = pl.make_many(W, 1, W, 1, W, CUFFT_C2C, (int)(F * 2 * H), s);
        cufftHandle k3b = pl.make_many(W, 1, W, 1, W, CUFFT_C2C, (int)(F * H), s);
        us_out[2] = time_us_per_frame([&] {
            ckf(cufftExecC2C(k3a, ca, cb, CUFFT_INVERSE), "K3 inv");
            ckf(cufftExecC2C(k3b, cb, ca, CUFFT_FORWARD), "K3 fwd");
        }, frames, reps, s);

        // columns of a row-major [H][W] plane: stride W, consecutive transforms 1 apart; one call per frame
        cufftHandle k4a = pl.make_many(H, W, 1, W, 1, CUFFT_C2C, W, s);
        cufftHandle k4b = pl.make_many(H, W, 1, W, 1, CUFFT_C2C, Wh, s);
        us_out[3] = time_us_per_frame([&] {
            for (size_t f = 0; f < F; ++f) {
                ckf(cufftExecC2C(k4a, ca + f * P, cb + f * P, CUFFT_FORWARD), "K4 fwd");
                ckf(cufftExecC2C(k4b, cb + f * P, ca + f * P, CUFFT_INVERSE), "K4 inv");
            }
        }, frames, reps, s);

        // the same transforms on contiguous columns (as if the neighbouring stages wrote / read a transposed plane for
        // free, which is what the column-blocked w3 layout of the hand-written path amounts to)
        cufftHandle k4c = pl.make_many(H, 1, H, 1, H, CUFFT_C2C, (int)(F * W), s);
        cufftHandle k4d = pl.make_many(H, 1, H, 1, H, CUFFT_C2C, (int)(F * Wh), s);
        us_out[5] = time_us_per_frame([&] {
            ckf(cufftExecC2C(k4c, ca, cb, CUFFT_FORWARD), "K4 fwd contiguous");
            ckf(cufftExecC2C(k4d, cb, ca, CUFFT_INVERSE), "K4 inv contiguous");
        }, frames, reps, s);

        cufftHandle k5 = pl.make_many(W, 1, Wh, 1, W, CUFFT_C2R, (int)(F * H), s);
        us_out[4] = time_us_per_frame([&] { ckf(cufftExecC2R(k5, ca, rb), "K5 C2R"); }, frames, reps, s);
        return 0;
    } catch (const Fail& f) {
        g_err = f.what;
        return -1;
    }
}

}  // extern "C"
