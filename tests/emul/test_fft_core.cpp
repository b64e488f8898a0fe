// Host check of fft_core.cuh: emulate the L/16 cooperating threads step by step and
// compare with a naive long-double DFT.  Build: g++ -O2 -std=c++17 -I<csrc> ...
#include <algorithm>
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

// pruned first pass: band of nc <= L/8 consecutive (mod L) non-zero inputs
template <int L, int DIR, class T> double run_pruned(int p0, int nc) {
    using F = Fft<L, DIR, T>;
    std::vector<cx<T>> table = F::make_table();
    std::vector<std::complex<long double>> x(L, 0), ref(L);
    unsigned s = 777u + L + p0;
    for (int c = 0; c < nc; ++c) {
        s = s * 1664525u + 1013904223u; double a = (s >> 8) / 16777216.0 - 0.5;
        s = s * 1664525u + 1013904223u; double b = (s >> 8) / 16777216.0 - 0.5;
        x[(p0 + c) % L] = {a, b};
    }
    for (int k = 0; k < L; ++k) {
        std::complex<long double> acc = 0;
        for (int c = 0; c < nc; ++c) {
            int n = (p0 + c) % L;
            long double a = DIR * 2.0L * M_PIl * ((long long)n * k % L) / L;
            acc += x[n] * std::complex<long double>(cosl(a), sinl(a));
        }
        ref[k] = acc;
    }
    const int TPF = F::TPF, M1 = L / 8;
    std::vector<cx<T>> regs(TPF * 16), smem(F::SMEM);
    for (int t = 0; t < TPF; ++t)
        for (int ii = 0; ii < 2; ++ii) {
            const int i = t + TPF * ii;
            const int c = ((i - p0) % M1 + M1) % M1;
            const int pp = (p0 + c) % L;
            cx<T> v = mk<T>(T(0), T(0));
            if (c < nc) v = mk<T>((T)x[pp].real(), (T)x[pp].imag());
            F::stepA_single(v, pp / M1, ii, t, smem.data());
        }
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

template <int L> int check_pruned() {
    double worst = 0;
    const int starts[] = {0, 5, L / 8 - 3, L / 2 - 7, L - 40, L - L / 8, 3 * L / 8 + 1};
    for (int p0 : starts)
        for (int nc : {1, L / 16 + 1, L / 8 - 15, L / 8}) {
            worst = std::max(worst, run_pruned<L, +1, float>(p0, nc));
            worst = std::max(worst, run_pruned<L, -1, float>(p0, nc));
        }
    printf("L=%4d  pruned first pass worst rel err %.2e\n", L, worst);
    return worst < 2e-6 ? 0 : 1;
}

template <int L> int check() {
    double e1 = run_one<L, -1, float>(), e2 = run_one<L, +1, float>();
    double e3 = run_one<L, -1, double>(), e4 = run_one<L, +1, double>();
    printf("L=%4d  f32 fwd %.2e inv %.2e   f64 fwd %.2e inv %.2e\n", L, e1, e2, e3, e4);
    return (e1 < 2e-6 && e2 < 2e-6 && e3 < 1e-14 && e4 < 1e-14) ? 0 : 1;
}

int main() {
    int bad = check<64>() + check<128>() + check<256>() + check<512>() + check<1024>() + check<2048>() + check<4096>();
    bad += check_pruned<512>() + check_pruned<1024>() + check_pruned<2048>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
