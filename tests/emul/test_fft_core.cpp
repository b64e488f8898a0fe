// Host check of fft_core.cuh: emulate the L/16 cooperating threads step by step and
// compare with a naive long-double DFT.  Build: g++ -O2 -std=c++17 -I<csrc> ...
#include <cmath>
#include <cstdio>
#include <vector>
#include <complex>
#include "fft_core.cuh"
using namespace fcd;

template <int L, int DIR, class T> double run_one() {
    using F = Fft<L, DIR, T>;
    std::vector<cx<T>> table = F::make_table();
    std::vector<std::complex<long double>> x(L), ref(L);
    unsigned s = 12345u + L;
    for (int n = 0; n < L; ++n) {
        s = s * 1664525u + 1013904223u; double a = (s >> 8) / 16777216.0 - 0.5;
        s = s * 1664525u + 1013904223u; double b = (s >> 8) / 16777216.0 - 0.5;
        x[n] = {a, b};
    }
    for (int k = 0; k < L; ++k) {
        std::complex<long double> acc = 0;
        for (int n = 0; n < L; ++n) {
            long double a = DIR * 2.0L * M_PIl * ((long long)n * k % L) / L;
            acc += x[n] * std::complex<long double>(cosl(a), sinl(a));
        }
        ref[k] = acc;
    }
    const int TPF = F::TPF;
    std::vector<cx<T>> regs(TPF * 16), smem(F::SMEM);
    for (int t = 0; t < TPF; ++t)
        for (int m = 0; m < 16; ++m) regs[t * 16 + m] = mk<T>((T)x[t + TPF * m].real(), (T)x[t + TPF * m].imag());
    for (int t = 0; t < TPF; ++t) F::stepA(&regs[t * 16], t, smem.data());
    for (int t = 0; t < TPF; ++t) F::stepB(&regs[t * 16], t, smem.data(), table.data());
    for (int t = 0; t < TPF; ++t) F::stepC(&regs[t * 16], t, smem.data());
    for (int t = 0; t < TPF; ++t) F::stepD(&regs[t * 16], t, smem.data(), table.data());
    long double num = 0, den = 0;
    for (int t = 0; t < TPF; ++t)
        for (int m = 0; m < 16; ++m) {
            auto r = ref[t + TPF * m];
            long double dx = regs[t * 16 + m].x - r.real(), dy = regs[t * 16 + m].y - r.imag();
            num += dx * dx + dy * dy; den += std::norm(r);
        }
    return (double)sqrtl(num / den);
}

template <int L> int check() {
    double e1 = run_one<L, -1, float>(), e2 = run_one<L, +1, float>();
    double e3 = run_one<L, -1, double>(), e4 = run_one<L, +1, double>();
    printf("L=%4d  f32 fwd %.2e inv %.2e   f64 fwd %.2e inv %.2e\n", L, e1, e2, e3, e4);
    return (e1 < 2e-6 && e2 < 2e-6 && e3 < 1e-14 && e4 < 1e-14) ? 0 : 1;
}

int main() {
    int bad = check<64>() + check<128>() + check<256>() + check<512>() + check<1024>() + check<2048>() + check<4096>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
