// test_cpp_host.cpp -- the reference's behavioural checklist through the C++ host layer
// (include/sgx_b200.hpp). Built and run by tests/test_cpp_host.py with g++:
//   host-only part  : parameter validation and messages, shapes, axes, windows -- runs without a GPU;
//   compute part    : known answers and properties of the reference's tests -- runs when a CUDA device is present,
//                     otherwise it is verified that compute fails loudly (no CPU fallback).
// Prints "CPP_HOST_OK host" or "CPP_HOST_OK host+gpu"; any failed expectation aborts with its line.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sgx_b200.hpp"

#define EXPECT(cond)                                                                  \
    do {                                                                              \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
    } while (0)

template <typename E, typename F> static std::string throws(F &&f) {
    try { f(); } catch (const E &e) { return e.what(); } catch (...) { return "<other exception>"; }
    return "<no exception>";
}
static bool contains(const std::string &s, const char *sub) { return s.find(sub) != std::string::npos; }

static std::vector<double> sine(size_t n, double sr, double f) {
    std::vector<double> x(n);
    for (size_t i = 0; i < n; ++i) x[i] = std::sin(2.0 * M_PI * f * static_cast<double>(i) / sr);   // tests/spectrogram_tests.rs:10-16
    return x;
}

int main() {
    using namespace sgx;
    // ---------------------------------------------------------------- host only
    EXPECT(contains(throws<InvalidInputError>([] { StftParams(512, 600); }), "hop_size must be <= n_fft"));           // tests/params_tests.rs
    EXPECT(contains(throws<InvalidInputError>([] { SpectrogramParams(StftParams(512, 256), 0.0); }), "sample_rate_hz"));
    EXPECT(contains(throws<InvalidInputError>([] { LogParams(NAN); }), "floor_db must be finite"));
    EXPECT(contains(throws<InvalidInputError>([] { ChromaParams(440.0, 0.0, 100.0); }), "f_min must be finite and > 0"));
    SpectrogramParams p(StftParams(512, 256, WindowType::hanning(), true), 16000.0);
    SpectrogramPlanner planner;
    auto lin = planner.linear_plan<double>(p);
    EXPECT(lin.output_shape(16000) == std::make_pair(size_t(257), size_t(63)));                                       // src/spectrogram.rs:505-507
    EXPECT(lin.output_shape(5) == std::make_pair(size_t(257), size_t(1)));                                            // tests/spectrogram_tests.rs:111-121
    auto f = lin.freq_axis();
    EXPECT(f.size() == 257 && f[0] == 0.0 && std::fabs(f[256] - 8000.0) < 1e-9);                                      // tests/spectrogram_tests.rs:183-236
    auto t = lin.times(63);
    EXPECT(t[0] == 0.0 && std::fabs(t[1] - 256.0 / 16000.0) < 1e-12);
    auto w = lin.window();
    EXPECT(w.size() == 512 && w[0] == 0.0 && std::fabs(w[255] - w[256]) < 1e-12);                                     // symmetric Hann (N - 1 denominator)
    EXPECT(contains(throws<InvalidInputError>([&] { planner.mel_plan<float>(p, MelParams(40, 0.0, 9000.0)); }), "Nyquist"));   // tests/spectrogram_tests.rs:147-158
    auto mel = planner.mel_plan<float>(SpectrogramParams(StftParams(400, 160), 16000.0), MelParams(128, 0.0, 8000.0), LogParams(-80.0), Amp::Decibels);
    EXPECT(mel.output_shape(480000) == std::make_pair(size_t(128), size_t(3001)));                                    // BASELINE configs[1]
    EXPECT(mel.kernel_name() == "r2c_fused_n400_tm");
    EXPECT(SpectrogramParams::speech_default(16000.0).stft().hop_size() == 160 && SpectrogramParams::music_default(44100.0).stft().n_fft() == 2048);
    EXPECT(contains(throws<InvalidInputError>([] { WindowType::custom({}); }), "cannot be empty"));
    EXPECT(contains(throws<InvalidInputError>([] { StftParams(8, 4, WindowType::custom({1, 2, 3}), true); }), "must match n_fft"));

    // ---------------------------------------------------------------- compute (needs a CUDA device)
    const auto x = sine(16000, 16000.0, 440.0);
    bool gpu = true;
    Matrix<double> spec;
    try { spec = lin.compute(x); } catch (const FftBackendError &e) {
        gpu = false;
        EXPECT(contains(e.what(), "no CPU fallback") || contains(e.what(), "CUDA"));                                  // fails loudly, never computes on the host
    }
    if (!gpu) { std::puts("CPP_HOST_OK host"); return 0; }

    EXPECT(spec.rows == 257 && spec.cols == 63);
    size_t peak = 0;
    for (size_t k = 1; k < 257; ++k) if (spec(k, 30) > spec(peak, 30)) peak = k;
    EXPECT(peak == 14);                                                                                                // examples/basic_linear.rs:50-62
    // compute_into with a wrong-size buffer: DimensionMismatch, rows first (tests/stft_plan_tests.rs:84-96)
    Matrix<double> bad(100, 63);
    try { lin.compute_into(x, bad); EXPECT(false); } catch (const DimensionMismatchError &e) { EXPECT(e.expected == 257 && e.got == 100); }
    // plan == one-shot (tests/stft_plan_tests.rs:59-82) and compute_frame == column (tests/streaming_tests.rs)
    StftPlan<double> sp(p);
    auto S = sp.compute(x);
    auto S2 = stft<double>(x, 512, 256);
    double d = 0.0;
    for (size_t i = 0; i < S.data.size(); ++i) d = std::fmax(d, std::abs(S.data[i] - S2.data[i]));
    EXPECT(d < 1e-10);
    auto col = sp.compute_frame_simple(x, 7);
    for (size_t k = 0; k < 257; ++k) EXPECT(std::abs(col[k] - S(k, 7)) < 1e-10);
    // DC known answer: fft([1, 1, 1] padded to 8)[0] has norm 3 (tests/fft_padding_tests.rs:149-158); longer input is an error
    auto X = rfft<double>({1.0, 1.0, 1.0}, 8);
    EXPECT(std::fabs(std::abs(X[0]) - 3.0) < 1e-10);
    EXPECT(contains(throws<InvalidInputError>([] { rfft<double>(std::vector<double>(9, 1.0), 8); }), "exceeds"));
    // inverse path: irfft(rfft(x)) == x and istft(stft(x)) == x away from the edges
    std::vector<double> y8 = {0.5, -1.0, 2.0, 0.25, 0.0, 1.5, -0.75, 3.0};
    auto back = irfft<double>(rfft<double>(y8, 8), 8);
    for (size_t i = 0; i < 8; ++i) EXPECT(std::fabs(back[i] - y8[i]) < 1e-12);
    auto rec = sp.istft(S);
    EXPECT(rec.size() == 62 * 256 + 512 - 512);
    for (size_t i = 512; i + 512 < rec.size(); ++i) EXPECT(std::fabs(rec[i] - x[i]) < 1e-10);
    // dB floor (tests/spectrogram_tests.rs:57-60) and f32 mel; MFCC shapes and the silence known answer c0 = -3200
    std::vector<float> xf(x.begin(), x.end());
    auto db = planner.mel_plan<float>(p, MelParams(40, 0.0, 8000.0), LogParams(-80.0), Amp::Decibels).compute(xf);
    EXPECT(db.rows == 40 && db.cols == 63);
    for (float v : db.data) EXPECT(v >= -80.0f - 1e-4f);
    auto c = mfcc<double>(std::vector<double>(8000, 0.0), StftParams(512, 160), 16000.0, 40, MfccParams(13, true, 0));
    EXPECT(c.rows == 13 && std::fabs(c(0, 3) + 3200.0) < 1e-9 && std::fabs(c(1, 3)) < 1e-9);
    EXPECT(mfcc<double>(x, StftParams(512, 160), 16000.0, 40, MfccParams(13, false, 22)).rows == 12);                  // tests/mfcc_tests.rs
    // chroma: a 440 Hz tone lands on pitch class A (index 9), L2-normalised frames have unit norm (src/chroma.rs)
    auto ch = chromagram<double>(x, StftParams(2048, 512), 16000.0, ChromaParams::music_standard());
    EXPECT(ch.rows == 12);
    size_t pc = 0;
    double n2 = 0.0;
    for (size_t r = 0; r < 12; ++r) { if (ch(r, 10) > ch(pc, 10)) pc = r; n2 += ch(r, 10) * ch(r, 10); }
    EXPECT(pc == 9 && std::fabs(n2 - 1.0) < 1e-12);
    // binaural: right = left / 2 -> ILD = 20 log10(2) dB and ILR = 0.5 on every bin with energy; identical channels -> IPD 0
    std::vector<double> xr(x.size());
    for (size_t i = 0; i < x.size(); ++i) xr[i] = 0.5 * x[i];
    BandParams bp(p, 300.0, 600.0);
    EXPECT(contains(throws<InvalidInputError>([&] { BandParams(p, 600.0, 300.0); }), "Start frequency must be less than end frequency."));
    auto ild = compute_ild_spectrogram<double>(x, xr, bp, sp);
    auto ilr = compute_ilr_spectrogram<double>(x, xr, bp, sp);
    auto ipd = compute_ipd_spectrogram<double>(x, x, bp, sp);
    EXPECT(ild.rows == bp.bins().second - bp.bins().first && ild.cols == 63);
    EXPECT(std::fabs(ild(4, 30) - 20.0 * std::log10(2.0)) < 1e-9 && std::fabs(ilr(4, 30) - 0.5) < 1e-12 && ipd(4, 30) == 0.0);
    std::puts("CPP_HOST_OK host+gpu");
    return 0;
}
