// TEST INFRASTRUCTURE ONLY: CPU emulation build of the product sources (see fcd_launch.cuh).
//   g++ -O2 -std=c++17 -DFCD_EMULATE -shared -fPIC -I include -I trapped-modes-ltg_b200/csrc \
//       -o tests/emul/libfcd_emul.so tests/emul/fcd_emul.cpp
#ifndef FCD_EMULATE
#error "build with -DFCD_EMULATE"
#endif
#include "fcd_plan.inl"

// emulation-only knob (not part of include/fcd_b200.h): thread order inside a phase
extern "C" void fcd_emul_set_thread_order(int order) { fcd::rt::emu_thread_order() = order; }

// emulation-only probes of the mask arithmetic (bit-exactness of the box filter in front of the threshold)
extern "C" double fcd_emul_div_by_const(double a, double d) { return fcd::div_by_const(a, d, 1.0 / d); }
// counts the elements of a[0..n) whose quotient by d differs (as bits) from the division
extern "C" long long fcd_emul_div_mismatches(const double* a, long long n, double d) {
    const double r = 1.0 / d;
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        volatile double want = a[i] / d;
        const double got = fcd::div_by_const(a[i], d, r);
        const double w = want;
        if (std::memcmp(&w, &got, sizeof(double)) != 0 && !(w == 0.0 && got == 0.0)) ++bad;
    }
    return bad;
}
// uniform_filter of [frames][H][W] float32 images as the mask path runs it: axis 0 (BoxLines), then axis 1
// (BoxRowsWarp for windows up to 32, BoxLines above)
extern "C" void fcd_emul_box_filter(const float* in, float* tmp, float* out, int frames, int H, int W, int size) {
    using namespace fcd;
    auto blocks = [](long long threads) { return (int)((threads + 255) / 256); };
    rt::launch<BoxLines>(blocks((long long)frames * W), 1, nullptr, BoxLinesParams{in, tmp, H, W, size, 0, (long long)frames * W});
    if (size <= 32)
        rt::launch<BoxRowsWarp>((int)(((long long)frames * H / 32 + 3) / 4), 1, nullptr,
                                BoxLinesParams{tmp, out, H, W, size, 1, (long long)frames * H});
    else
        rt::launch<BoxLines>(blocks((long long)frames * H), 1, nullptr, BoxLinesParams{tmp, out, H, W, size, 1, (long long)frames * H});
}
