// sgx_b200.hpp -- C++17 host-side mirror of the reference's plan API on top of the C ABI (sgx_b200.h).
//
// The reference is a compiled (Rust) crate; no Rust toolchain exists in this repository's build image, so the compiled
// host layer above the C ABI is this header: the same names, argument meaning and error behaviour as the crate
// (bare :N = src/spectrogram.rs:N of the reference checkout), RAII plans instead of `&mut self` borrows, exceptions
// instead of `SpectrogramResult`. Header only; link with -lsgx_b200. Nothing here computes on the CPU: every number
// comes from the CUDA library, and without a device every compute call throws sgx::FftBackendError.
//
//   sgx::SpectrogramParams params(sgx::StftParams(400, 160, sgx::WindowType::hanning(), true), 16000.0);
//   auto plan = sgx::SpectrogramPlanner().mel_plan<float>(params, sgx::MelParams(128, 0.0, 8000.0),
//                                                         sgx::LogParams(-80.0), sgx::Amp::Decibels);
//   sgx::Matrix<float> spec = plan.compute(samples);            // reference-style call, host buffers
//   plan.compute_batch(d_clips, n_clips, n_samples, n_samples, d_out, stream);   // batched, device resident
#pragma once

#include <complex>
#include <cstddef>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sgx_b200.h"

namespace sgx {

// ---- SpectrogramError (src/error.rs:13-28): one exception type per variant
struct SpectrogramError : std::runtime_error {
    sgx_status status;
    explicit SpectrogramError(sgx_status s, const std::string &m) : std::runtime_error(m), status(s) {}
};
struct InvalidInputError : SpectrogramError { using SpectrogramError::SpectrogramError; };
struct DimensionMismatchError : SpectrogramError {
    size_t expected = 0, got = 0;
    DimensionMismatchError(const std::string &m, size_t e, size_t g) : SpectrogramError(SGX_DIMENSION_MISMATCH, m), expected(e), got(g) {}
};
struct FftBackendError : SpectrogramError { using SpectrogramError::SpectrogramError; };
struct InternalError : SpectrogramError { using SpectrogramError::SpectrogramError; };

inline void check(sgx_status st) {
    if (st == SGX_OK) return;
    const std::string msg = sgx_last_error_message();
    switch (st) {
        case SGX_INVALID_INPUT: throw InvalidInputError(st, msg);
        case SGX_DIMENSION_MISMATCH: {
            size_t e = 0, g = 0;
            sgx_last_dimension_mismatch(&e, &g);
            throw DimensionMismatchError(msg, e, g);
        }
        case SGX_BACKEND_ERROR: throw FftBackendError(st, msg);
        default: throw InternalError(st, msg);
    }
}

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr sgx_dtype value = SGX_F32; };
template <> struct dtype_of<double> { static constexpr sgx_dtype value = SGX_F64; };

// ---- WindowType (src/window.rs:19-50)
struct WindowType {
    sgx_window kind = SGX_WIN_HANNING;
    double param = 0.0;
    std::vector<double> coefficients;                // Custom only
    static WindowType rectangular() { return {SGX_WIN_RECTANGULAR, 0.0, {}}; }
    static WindowType hanning() { return {SGX_WIN_HANNING, 0.0, {}}; }
    static WindowType hamming() { return {SGX_WIN_HAMMING, 0.0, {}}; }
    static WindowType blackman() { return {SGX_WIN_BLACKMAN, 0.0, {}}; }
    static WindowType kaiser(double beta) { return {SGX_WIN_KAISER, beta, {}}; }
    static WindowType gaussian(double std_samples) { return {SGX_WIN_GAUSSIAN, std_samples, {}}; }
    static WindowType custom(std::vector<double> c) {
        if (c.empty()) throw InvalidInputError(SGX_INVALID_INPUT, "Custom window coefficients cannot be empty");   // src/window.rs
        return {SGX_WIN_CUSTOM, 0.0, std::move(c)};
    }
};

// ---- StftParams (:3452-3506), SpectrogramParams (:4108-4140)
class StftParams {
public:
    StftParams(size_t n_fft, size_t hop_size, WindowType window = WindowType::hanning(), bool centre = true)
        : n_fft_(n_fft), hop_(hop_size), window_(std::move(window)), centre_(centre) {
        if (n_fft == 0) throw InvalidInputError(SGX_INVALID_INPUT, "n_fft must be set");
        if (hop_size == 0) throw InvalidInputError(SGX_INVALID_INPUT, "hop_size must be set");
        if (hop_size > n_fft) throw InvalidInputError(SGX_INVALID_INPUT, "hop_size must be <= n_fft");            // :3485
        if (window_.kind == SGX_WIN_CUSTOM && window_.coefficients.size() != n_fft)                                 // :3490-3497
            throw InvalidInputError(SGX_INVALID_INPUT, "Custom window size (" + std::to_string(window_.coefficients.size()) +
                                                           ") must match n_fft (" + std::to_string(n_fft) + ")");
    }
    size_t n_fft() const { return n_fft_; }
    size_t hop_size() const { return hop_; }
    const WindowType &window() const { return window_; }
    bool centre() const { return centre_; }

private:
    size_t n_fft_, hop_;
    WindowType window_;
    bool centre_;
};

class SpectrogramParams {
public:
    SpectrogramParams(StftParams stft, double sample_rate_hz) : stft_(std::move(stft)), sr_(sample_rate_hz) {
        if (!(sample_rate_hz > 0.0) || sample_rate_hz != sample_rate_hz || sample_rate_hz > 1.7e308)
            throw InvalidInputError(SGX_INVALID_INPUT, "sample_rate_hz must be finite and > 0");                    // :4130
    }
    static SpectrogramParams speech_default(double sr) { return {StftParams(512, 160), sr}; }                      // :4215-4230
    static SpectrogramParams music_default(double sr) { return {StftParams(2048, 512), sr}; }                       // :4232-4248
    const StftParams &stft() const { return stft_; }
    double sample_rate_hz() const { return sr_; }
    double nyquist_hz() const { return sr_ * 0.5; }
    double frame_period_seconds() const { return static_cast<double>(stft_.hop_size()) / sr_; }                    // :4268-4271

private:
    StftParams stft_;
    double sr_;
};

// ---- scale parameters: MelParams (:3744-3920), ErbParams (src/erb.rs:30-88), LogHzParams (:3955-3975), LogParams (:4052-4100)
enum class MelNorm { None = SGX_MELNORM_NONE, Slaney = SGX_MELNORM_SLANEY, L1 = SGX_MELNORM_L1, L2 = SGX_MELNORM_L2 };
struct MelParams {
    size_t n_mels; double f_min, f_max; MelNorm norm;
    MelParams(size_t n, double lo, double hi, MelNorm nm = MelNorm::None) : n_mels(n), f_min(lo), f_max(hi), norm(nm) {}
};
enum class ErbSpacing { Linear = SGX_ERB_LINEAR, AppleTr35 = SGX_ERB_APPLE_TR35 };
struct ErbParams {
    size_t n_filters; double f_min, f_max; ErbSpacing spacing;
    ErbParams(size_t n, double lo, double hi, ErbSpacing s = ErbSpacing::Linear) : n_filters(n), f_min(lo), f_max(hi), spacing(s) {}
};
struct LogHzParams {
    size_t n_bins; double f_min, f_max;
    LogHzParams(size_t n, double lo, double hi) : n_bins(n), f_min(lo), f_max(hi) {}
};
struct LogParams {
    double floor_db;
    explicit LogParams(double floor) : floor_db(floor) {
        if (floor != floor || floor > 1.7e308 || floor < -1.7e308) throw InvalidInputError(SGX_INVALID_INPUT, "floor_db must be finite");   // :4072
    }
};
struct MfccParams {                                                   // src/mfcc.rs:21-141
    size_t n_mfcc = 13; bool include_c0 = true; size_t lifter = 22;
    MfccParams() = default;
    MfccParams(size_t n, bool c0 = true, size_t lift = 22) : n_mfcc(n), include_c0(c0), lifter(lift) {
        if (n == 0) throw InvalidInputError(SGX_INVALID_INPUT, "n_mfcc must be non-zero");
    }
    static MfccParams speech_standard() { return MfccParams(13); }
};
enum class ChromaNorm { None = SGX_CHROMANORM_NONE, L1 = SGX_CHROMANORM_L1, L2 = SGX_CHROMANORM_L2, Max = SGX_CHROMANORM_MAX };
struct ChromaParams {                                                 // src/chroma.rs:18-182
    double tuning = 440.0, f_min = 32.7, f_max = 4186.0; ChromaNorm norm = ChromaNorm::L2;
    ChromaParams() = default;
    ChromaParams(double t, double lo, double hi, ChromaNorm n = ChromaNorm::L2) : tuning(t), f_min(lo), f_max(hi), norm(n) {
        if (!(t > 0.0) || t > 1.7e308) throw InvalidInputError(SGX_INVALID_INPUT, "tuning must be finite and > 0");   // :82-86
        if (!(lo > 0.0) || lo > 1.7e308) throw InvalidInputError(SGX_INVALID_INPUT, "f_min must be finite and > 0");  // :87-91
        if (hi <= lo) throw InvalidInputError(SGX_INVALID_INPUT, "f_max must be > f_min");                            // :92-94
    }
    static ChromaParams music_standard() { return ChromaParams(); }
};

enum class Amp { Power = SGX_AMP_POWER, Magnitude = SGX_AMP_MAGNITUDE, Decibels = SGX_AMP_DECIBELS };   // AmpScaleSpec implementors (:1986-2037)

// ---- results: Array2<T> / Array2<Complex<T>>, row-major (rows, n_frames)
template <typename V> struct Matrix {
    size_t rows = 0, cols = 0;
    std::vector<V> data;
    Matrix() = default;
    Matrix(size_t r, size_t c) : rows(r), cols(c), data(r * c) {}
    V &operator()(size_t r, size_t c) { return data[r * cols + c]; }
    const V &operator()(size_t r, size_t c) const { return data[r * cols + c]; }
};

// ---- one owned sgx_plan*
class NativePlan {
public:
    NativePlan() = default;
    explicit NativePlan(const sgx_plan_desc &d, std::vector<double> custom = {}) : custom_(std::move(custom)) {
        sgx_plan_desc dd = d;
        if (!custom_.empty()) { dd.custom_window = custom_.data(); dd.custom_window_len = custom_.size(); }
        check(sgx_plan_create(&dd, &h_));
    }
    NativePlan(NativePlan &&o) noexcept : h_(o.h_), custom_(std::move(o.custom_)) { o.h_ = nullptr; }
    NativePlan &operator=(NativePlan &&o) noexcept {
        if (this != &o) { reset(); h_ = o.h_; custom_ = std::move(o.custom_); o.h_ = nullptr; }
        return *this;
    }
    NativePlan(const NativePlan &) = delete;
    NativePlan &operator=(const NativePlan &) = delete;
    ~NativePlan() { reset(); }
    sgx_plan *get() const { return h_; }

private:
    void reset() { if (h_) { sgx_plan_destroy(h_); h_ = nullptr; } }
    sgx_plan *h_ = nullptr;
    std::vector<double> custom_;
};

inline sgx_plan_desc base_desc(const SpectrogramParams &p, sgx_dtype dt, int device) {
    sgx_plan_desc d;
    std::memset(&d, 0, sizeof d);
    d.dtype = dt;
    d.n_fft = p.stft().n_fft();
    d.hop_size = p.stft().hop_size();
    d.centre = p.stft().centre() ? 1 : 0;
    d.window = p.stft().window().kind;
    d.window_param = p.stft().window().param;
    d.sample_rate_hz = p.sample_rate_hz();
    d.device = device;
    return d;
}

// ---- SpectrogramPlan<F, A, T> (:172-519); the frequency scale and amplitude scale are run-time members here
template <typename T> class SpectrogramPlan {
public:
    SpectrogramPlan(const sgx_plan_desc &d, const SpectrogramParams &p) : params_(p), plan_(d, p.stft().window().coefficients), rows_(0) {
        size_t nf = 0;
        check(sgx_plan_output_shape(plan_.get(), 1, &rows_, &nf));
    }
    const SpectrogramParams &params() const { return params_; }
    std::pair<size_t, size_t> output_shape(size_t signal_length) const {                     // :512-519
        size_t r = 0, c = 0;
        check(sgx_plan_output_shape(plan_.get(), signal_length, &r, &c));
        return {r, c};
    }
    std::vector<double> freq_axis() const {                                                   // :202-215
        std::vector<double> f(rows_);
        check(sgx_plan_axes(plan_.get(), 0, f.data(), nullptr));
        return f;
    }
    std::vector<double> times(size_t n_frames) const {                                        // :2128-2139
        std::vector<double> t(n_frames);
        check(sgx_plan_axes(plan_.get(), n_frames, nullptr, t.data()));
        return t;
    }
    Matrix<T> compute(const std::vector<T> &samples) {                                        // :240-294
        if (samples.empty()) throw InvalidInputError(SGX_INVALID_INPUT, "samples must be non-empty");
        const auto [r, c] = output_shape(samples.size());
        Matrix<T> out(r, c);
        compute_into(samples, out);
        return out;
    }
    void compute_into(const std::vector<T> &samples, Matrix<T> &out) {                        // :414-477
        check(sgx_plan_compute_batch(plan_.get(), samples.data(), 1, samples.size(), samples.size(), out.data.data(), out.rows, out.cols, 0, nullptr));
    }
    std::vector<T> compute_frame(const std::vector<T> &samples, size_t frame_idx) {           // :335-372
        std::vector<T> out(rows_);
        check(sgx_plan_compute_frame(plan_.get(), samples.data(), samples.size(), frame_idx, out.data(), nullptr));
        return out;
    }
    // New relative to the reference: `for s in clips { plan.compute_into(s, out[i]) }` as one call. Host or device
    // pointers (both the same kind); device calls are asynchronous on `stream` (a cudaStream_t).
    void compute_batch(const T *clips, size_t n_clips, size_t n_samples, size_t clip_stride, T *out, void *stream = nullptr) {
        const auto [r, c] = output_shape(n_samples);
        check(sgx_plan_compute_batch(plan_.get(), clips, n_clips, n_samples, clip_stride, out, r, c, 0, stream));
    }
    std::vector<T> window() const {
        std::vector<T> w(params_.stft().n_fft());
        check(sgx_plan_window(plan_.get(), w.data()));
        return w;
    }
    std::string kernel_name() const { return sgx_plan_kernel_name(plan_.get()); }
    sgx_plan *native() const { return plan_.get(); }

private:
    SpectrogramParams params_;
    NativePlan plan_;
    size_t rows_;
};

// ---- StftPlan<T> (:1173-1637) and the inverse (:4813-4911)
template <typename T> class StftPlan {
public:
    using C = std::complex<T>;
    explicit StftPlan(const SpectrogramParams &p, int device = -1) : params_(p), plan_(desc(p, device), p.stft().window().coefficients) {}
    std::pair<size_t, size_t> output_shape(size_t signal_length) const {                     // :1596-1602
        size_t r = 0, c = 0;
        check(sgx_plan_output_shape(plan_.get(), signal_length, &r, &c));
        return {r, c};
    }
    Matrix<C> compute(const std::vector<T> &samples) {                                        // :1424-1458
        if (samples.empty()) throw InvalidInputError(SGX_INVALID_INPUT, "samples must be non-empty");
        const auto [r, c] = output_shape(samples.size());
        Matrix<C> out(r, c);
        compute_into(samples, out);
        return out;
    }
    void compute_into(const std::vector<T> &samples, Matrix<C> &out) {                        // :1548-1580
        check(sgx_plan_compute_batch(plan_.get(), samples.data(), 1, samples.size(), samples.size(), out.data.data(), out.rows, out.cols, 0, nullptr));
    }
    std::vector<C> compute_frame_simple(const std::vector<T> &samples, size_t frame_idx) {    // :1500-1507
        std::vector<C> out(params_.stft().n_fft() / 2 + 1);
        check(sgx_plan_compute_frame(plan_.get(), samples.data(), samples.size(), frame_idx, out.data(), nullptr));
        return out;
    }
    std::vector<T> istft(const Matrix<C> &stft_matrix) {                                      // istft() :4813-4911 with this plan's parameters
        if (stft_matrix.rows != params_.stft().n_fft() / 2 + 1)
            throw DimensionMismatchError("Dimension mismatch: expected " + std::to_string(params_.stft().n_fft() / 2 + 1) + ", got " +
                                             std::to_string(stft_matrix.rows), params_.stft().n_fft() / 2 + 1, stft_matrix.rows);
        size_t n = 0;
        check(sgx_plan_istft(plan_.get(), nullptr, 1, stft_matrix.cols, nullptr, &n, nullptr));
        std::vector<T> out(n);
        check(sgx_plan_istft(plan_.get(), stft_matrix.data.data(), 1, stft_matrix.cols, out.data(), &n, nullptr));
        return out;
    }
    std::string kernel_name() const { return sgx_plan_kernel_name(plan_.get()); }
    sgx_plan *native() const { return plan_.get(); }

private:
    static sgx_plan_desc desc(const SpectrogramParams &p, int device) {
        sgx_plan_desc d = base_desc(p, dtype_of<T>::value, device);
        d.mapping = SGX_MAP_LINEAR; d.amp = SGX_AMP_POWER; d.output = SGX_OUT_COMPLEX_STFT;
        return d;
    }
    SpectrogramParams params_;
    NativePlan plan_;
};

// ---- SpectrogramPlanner (:640-1152)
class SpectrogramPlanner {
public:
    explicit SpectrogramPlanner(int device = -1) : device_(device) {}
    template <typename T> SpectrogramPlan<T> linear_plan(const SpectrogramParams &p, std::optional<LogParams> db = std::nullopt, Amp amp = Amp::Power) const {   // :893-917
        sgx_plan_desc d = base_desc(p, dtype_of<T>::value, device_);
        d.mapping = SGX_MAP_LINEAR;
        return finish<T>(d, p, db, amp);
    }
    template <typename T> SpectrogramPlan<T> mel_plan(const SpectrogramParams &p, const MelParams &mel, std::optional<LogParams> db = std::nullopt, Amp amp = Amp::Power) const {   // :944-977
        sgx_plan_desc d = base_desc(p, dtype_of<T>::value, device_);
        d.mapping = SGX_MAP_MEL; d.n_bands = mel.n_mels; d.f_min = mel.f_min; d.f_max = mel.f_max; d.mel_norm = static_cast<sgx_mel_norm>(mel.norm);
        return finish<T>(d, p, db, amp);
    }
    template <typename T> SpectrogramPlan<T> erb_plan(const SpectrogramParams &p, const ErbParams &erb, std::optional<LogParams> db = std::nullopt, Amp amp = Amp::Power) const {   // :1005-1040
        sgx_plan_desc d = base_desc(p, dtype_of<T>::value, device_);
        d.mapping = SGX_MAP_ERB; d.n_bands = erb.n_filters; d.f_min = erb.f_min; d.f_max = erb.f_max; d.erb_spacing = static_cast<sgx_erb_spacing>(erb.spacing);
        return finish<T>(d, p, db, amp);
    }
    template <typename T> SpectrogramPlan<T> log_hz_plan(const SpectrogramParams &p, const LogHzParams &lh, std::optional<LogParams> db = std::nullopt, Amp amp = Amp::Power) const {   // :1067-1102
        sgx_plan_desc d = base_desc(p, dtype_of<T>::value, device_);
        d.mapping = SGX_MAP_LOGHZ; d.n_bands = lh.n_bins; d.f_min = lh.f_min; d.f_max = lh.f_max;
        return finish<T>(d, p, db, amp);
    }
    template <typename T> StftPlan<T> stft_plan(const SpectrogramParams &p) const { return StftPlan<T>(p, device_); }
    template <typename T> Matrix<std::complex<T>> compute_stft(const std::vector<T> &samples, const SpectrogramParams &p) const {   // :722-729
        return StftPlan<T>(p, device_).compute(samples);
    }

private:
    template <typename T> SpectrogramPlan<T> finish(sgx_plan_desc d, const SpectrogramParams &p, const std::optional<LogParams> &db, Amp amp) const {
        d.amp = static_cast<sgx_amp>(amp);
        d.has_floor_db = db.has_value() ? 1 : 0;
        d.floor_db = db ? db->floor_db : 0.0;
        d.output = SGX_OUT_SPECTROGRAM;
        return SpectrogramPlan<T>(d, p);
    }
    int device_;
};

// ---- free functions: stft() (:4733-4747), istft() (:4813-4911), fft()/rfft (:4490-4520), irfft (:4789-4811)
template <typename T> Matrix<std::complex<T>> stft(const std::vector<T> &samples, size_t n_fft, size_t hop_size, WindowType window = WindowType::hanning(), bool centre = true) {
    return StftPlan<T>(SpectrogramParams(StftParams(n_fft, hop_size, std::move(window), centre), 1.0)).compute(samples);
}
template <typename T> std::vector<T> istft(const Matrix<std::complex<T>> &m, size_t n_fft, size_t hop_size, WindowType window = WindowType::hanning(), bool centre = true) {
    return StftPlan<T>(SpectrogramParams(StftParams(n_fft, hop_size, std::move(window), centre), 1.0)).istft(m);
}
template <typename T> std::vector<std::complex<T>> rfft(const std::vector<T> &samples, size_t n_fft) {
    std::vector<std::complex<T>> out(n_fft / 2 + 1);
    check(sgx_rfft(dtype_of<T>::value, samples.data(), samples.size(), n_fft, out.data(), -1, nullptr));
    return out;
}
template <typename T> std::vector<T> irfft(const std::vector<std::complex<T>> &spectrum, size_t n_fft) {
    std::vector<T> out(n_fft);
    check(sgx_irfft(dtype_of<T>::value, spectrum.data(), spectrum.size(), n_fft, out.data(), -1, nullptr));
    return out;
}

// ---- MFCC: mfcc_from_log_mel (src/mfcc.rs:224-273) and the fused mfcc() (:359-379)
template <typename T> Matrix<T> mfcc_from_log_mel(const Matrix<T> &log_mel, const MfccParams &mp) {
    if (mp.n_mfcc > log_mel.rows) throw InvalidInputError(SGX_INVALID_INPUT, "n_mfcc must be <= n_mels");            // :231-233
    Matrix<T> out(mp.n_mfcc - ((mp.include_c0 || mp.n_mfcc == 1) ? 0 : 1), log_mel.cols);
    check(sgx_mfcc_from_log_mel(dtype_of<T>::value, log_mel.data.data(), 1, log_mel.rows, log_mel.cols, mp.n_mfcc, mp.include_c0 ? 1 : 0,
                                mp.lifter, out.data.data(), -1, nullptr));
    return out;
}
template <typename T> class MfccPlan {
public:
    MfccPlan(const StftParams &stft, double sample_rate, size_t n_mels, const MfccParams &mp, int device = -1)
        : plan_(desc(stft, sample_rate, n_mels, mp, device), SpectrogramParams(stft, sample_rate)) {}
    Matrix<T> compute(const std::vector<T> &samples) { return plan_.compute(samples); }
    SpectrogramPlan<T> &plan() { return plan_; }

private:
    static sgx_plan_desc desc(const StftParams &stft, double sr, size_t n_mels, const MfccParams &mp, int device) {
        sgx_plan_desc d = base_desc(SpectrogramParams(stft, sr), dtype_of<T>::value, device);
        d.mapping = SGX_MAP_MEL; d.n_bands = n_mels; d.f_min = 0.0; d.f_max = sr / 2.0; d.mel_norm = SGX_MELNORM_NONE;   // src/mfcc.rs:366-371
        d.amp = SGX_AMP_DECIBELS; d.has_floor_db = 1; d.floor_db = -80.0;
        d.output = SGX_OUT_MFCC; d.n_mfcc = mp.n_mfcc; d.include_c0 = mp.include_c0 ? 1 : 0; d.lifter = mp.lifter;
        return d;
    }
    SpectrogramPlan<T> plan_;
};
template <typename T> Matrix<T> mfcc(const std::vector<T> &samples, const StftParams &stft, double sample_rate, size_t n_mels, const MfccParams &mp) {
    return MfccPlan<T>(stft, sample_rate, n_mels, mp).compute(samples);
}

// ---- chroma: chromagram() (src/chroma.rs:487-503), chromagram_from_spectrogram (:365-404)
template <typename T> Matrix<T> chromagram(const std::vector<T> &samples, const StftParams &stft, double sample_rate, const ChromaParams &cp, int device = -1) {
    sgx_plan_desc d = base_desc(SpectrogramParams(stft, sample_rate), dtype_of<T>::value, device);
    d.mapping = SGX_MAP_CHROMA; d.n_bands = 12; d.f_min = cp.f_min; d.f_max = cp.f_max;
    d.chroma_tuning = cp.tuning; d.chroma_norm = static_cast<sgx_chroma_norm>(cp.norm);
    d.amp = SGX_AMP_MAGNITUDE; d.output = SGX_OUT_SPECTROGRAM;
    return SpectrogramPlan<T>(d, SpectrogramParams(stft, sample_rate)).compute(samples);
}
template <typename T> Matrix<T> chromagram_from_spectrogram(const Matrix<T> &spec, double sample_rate, size_t n_fft, const ChromaParams &cp) {
    Matrix<T> out(12, spec.cols);
    check(sgx_chroma_from_spectrogram(dtype_of<T>::value, spec.data.data(), 1, spec.rows, spec.cols, sample_rate, n_fft, cp.tuning, cp.f_min,
                                      cp.f_max, static_cast<sgx_chroma_norm>(cp.norm), out.data.data(), -1, nullptr));
    return out;
}

// ---- binaural cues (src/binaural.rs): compute_{itd,ipd,ild,ilr}_spectrogram(audio = [left, right], params, plan)
struct BandParams {                                                   // ITD/IPD/ILD/ILRSpectrogramParams::new (:410-444, :775-806, ...)
    SpectrogramParams spectrogram_params;
    double start_freq, end_freq;
    size_t magphase_power = 1;                                        // ITD only
    bool wrapped = true;                                              // IPD only
    BandParams(SpectrogramParams sp, double start, double stop, size_t power = 1, bool wrap = true)
        : spectrogram_params(std::move(sp)), start_freq(start), end_freq(stop), magphase_power(power), wrapped(wrap) {
        if (start <= 0.0 || stop <= 0.0) throw InvalidInputError(SGX_INVALID_INPUT, "Start and end frequencies must be positive.");
        if (start >= stop) throw InvalidInputError(SGX_INVALID_INPUT, "Start frequency must be less than end frequency.");
        if (stop > spectrogram_params.sample_rate_hz() / 2.0) throw InvalidInputError(SGX_INVALID_INPUT, "End frequency must be less than Nyquist frequency.");
        if (power == 0) throw InvalidInputError(SGX_INVALID_INPUT, "magphase_power must be non-zero");
    }
    std::pair<size_t, size_t> bins() const {                          // (f / bin_width).round() as usize (:476-481)
        const double bw = spectrogram_params.sample_rate_hz() / static_cast<double>(spectrogram_params.stft().n_fft());
        return {static_cast<size_t>(start_freq / bw + 0.5), static_cast<size_t>(end_freq / bw + 0.5)};
    }
};
template <typename T> Matrix<T> compute_binaural(sgx_binaural_cue cue, const std::vector<T> &left, const std::vector<T> &right, const BandParams &bp, StftPlan<T> &plan) {
    if (left.empty() || left.size() != right.size()) throw InvalidInputError(SGX_INVALID_INPUT, "left and right must be equally long and non-empty");
    const auto [b0, b1] = bp.bins();
    if (b1 <= b0) throw InvalidInputError(SGX_INVALID_INPUT, "Frequency range should have at least one bin");
    const size_t nf = plan.output_shape(left.size()).second;
    Matrix<T> out(b1 - b0, nf);
    check(sgx_plan_compute_binaural(plan.native(), cue, left.data(), right.data(), 1, left.size(), left.size(), bp.start_freq, bp.end_freq,
                                    bp.magphase_power, bp.wrapped ? 1 : 0, out.data.data(), out.rows, out.cols, nullptr));
    return out;
}
template <typename T> Matrix<T> compute_itd_spectrogram(const std::vector<T> &l, const std::vector<T> &r, const BandParams &bp, StftPlan<T> &plan) { return compute_binaural<T>(SGX_CUE_ITD, l, r, bp, plan); }   // :472-580
template <typename T> Matrix<T> compute_ipd_spectrogram(const std::vector<T> &l, const std::vector<T> &r, const BandParams &bp, StftPlan<T> &plan) { return compute_binaural<T>(SGX_CUE_IPD, l, r, bp, plan); }   // :830-917
template <typename T> Matrix<T> compute_ild_spectrogram(const std::vector<T> &l, const std::vector<T> &r, const BandParams &bp, StftPlan<T> &plan) { return compute_binaural<T>(SGX_CUE_ILD, l, r, bp, plan); }   // :1187-1262
template <typename T> Matrix<T> compute_ilr_spectrogram(const std::vector<T> &l, const std::vector<T> &r, const BandParams &bp, StftPlan<T> &plan) { return compute_binaural<T>(SGX_CUE_ILR, l, r, bp, plan); }   // :1530-1620

}  // namespace sgx
