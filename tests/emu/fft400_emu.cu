// Host emulation of the n400 kernel's task functions (spectrograms_b200/csrc/fft400_core.cuh): runs pass1_task /
// pass2_task for every (frame, role) of one tile sequentially on the CPU, with the same shared-memory layouts.
// Test infrastructure: lets the CPU-only container check the index maps, twiddles and butterflies against the oracle.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../spectrograms_b200/csrc/fft400_core.cuh"

using namespace sgx::f400;

extern "C" int emu_fft400_tile(const float *samples, long long n_samples, long long f0, const float *win, float *power_out /*[201][32]*/) {
    static Consts c;
    for (int i = 0; i < kN; ++i) c.win[i] = win[i];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int n2 = 0; n2 < 20; ++n2) {
            const long double a = -2.0L * pi * (long double)((n2 * k1) % 400) / 400.0L;
            const double s = (k1 == 0 || k1 == 10) ? 1.0 : 0.5;
            c.tw2[k1][n2] = make_float2((float)(s * (double)cosl(a)), (float)(s * (double)sinl(a)));
        }
    std::vector<float> sig(kSigWords, 0.f), ybuf(kYWords, 0.f), p(kPWords, 0.f);
    const long long s0 = f0 * kHop - kN / 2;
    for (int u = 0; u < kTileSamples; ++u) {
        const long long s = s0 + u;
        sig[sig_word(u)] = (s >= 0 && s < n_samples) ? samples[s] : 0.f;
    }
    for (int t = 0; t < 10; ++t)
        for (int f = 0; f < kFT; ++f) pass1_task(sig.data(), ybuf.data(), c.win, f, t);
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int f = 0; f < kFT; ++f) pass2_task(ybuf.data(), p.data(), c, f, k1);
    for (int b = 0; b < kBins; ++b)
        for (int f = 0; f < kFT; ++f) power_out[b * kFT + f] = p[b * kFT + frame_col(f)];
    return 0;
}
