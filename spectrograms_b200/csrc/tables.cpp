// tables.cpp -- per-plan host tables in f64 (see tables.hpp). Bare ":N" citations are src/spectrogram.rs:N of the
// reference. Build with -ffp-contract=off.
#include "tables.hpp"

#include <cmath>
#include <cstdio>
#include <limits>

namespace sgx {

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

[[noreturn]] void invalid(const std::string &m) { throw Error{SGX_INVALID_INPUT, m}; }

// modified_bessel_i0 (:2237-2259)
double bessel_i0(double x) {
    const double ax = std::fabs(x);
    if (ax <= 3.75) {
        const double t = x / 3.75, t2 = t * t;
        return 1.0 + t2 * (3.5156229 + t2 * (3.0899424 + t2 * (1.2067492 + t2 * (0.2659732 + t2 * (0.0360768 + t2 * 0.0045813)))));
    }
    const double t = 3.75 / ax;
    const double poly =
        0.39894228 +
        t * (0.01328592 +
             t * (0.00225319 +
                  t * (-0.00157565 + t * (0.00916281 + t * (-0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377)))))));
    return (std::exp(ax) / (std::sqrt(ax) * std::sqrt(2.0 * kPi))) * poly;
}

// make_window (:2159-2235)
void make_window(const sgx_plan_desc &d, std::vector<double> &w) {
    const size_t n = d.n_fft;
    w.assign(n, 0.0);
    const double n1 = static_cast<double>(n - 1);
    switch (d.window) {
        case SGX_WIN_RECTANGULAR:
            for (auto &v : w) v = 1.0;
            break;
        case SGX_WIN_HANNING:
            for (size_t i = 0; i < n; ++i) w[i] = std::fma(0.5, -std::cos(2.0 * kPi * static_cast<double>(i) / n1), 0.5);
            break;
        case SGX_WIN_HAMMING:
            for (size_t i = 0; i < n; ++i) w[i] = std::fma(0.46, -std::cos(2.0 * kPi * static_cast<double>(i) / n1), 0.54);
            break;
        case SGX_WIN_BLACKMAN:
            for (size_t i = 0; i < n; ++i) {
                const double a = 2.0 * kPi * static_cast<double>(i) / n1;
                w[i] = std::fma(0.08, std::cos(2.0 * a), std::fma(0.5, -std::cos(a), 0.42));
            }
            break;
        case SGX_WIN_KAISER: {
            if (n == 1) { w[0] = 1.0; break; }
            const double beta = d.window_param;
            const double denom = bessel_i0(beta);
            const double n_max = n1 / 2.0;
            for (size_t i = 0; i < n; ++i) {
                double ratio = 0.0;
                if (n_max != 0.0) {
                    const double nrm = (static_cast<double>(i) - n_max) / n_max;
                    ratio = std::fmax(1.0 - nrm * nrm, 0.0);
                }
                w[i] = denom == 0.0 ? 0.0 : bessel_i0(beta * std::sqrt(ratio)) / denom;
            }
            break;
        }
        case SGX_WIN_GAUSSIAN: {
            const double centre = n1 / 2.0;
            for (size_t i = 0; i < n; ++i) {
                const double q = (static_cast<double>(i) - centre) / d.window_param;
                w[i] = std::exp(-0.5 * (q * q));
            }
            break;
        }
        case SGX_WIN_CUSTOM:
            for (size_t i = 0; i < n; ++i) w[i] = d.custom_window[i];
            break;
    }
}

// Slaney mel scale (:2268-2300)
constexpr double kFsp = 200.0 / 3.0;
constexpr double kMinLogHz = 1000.0;
constexpr double kMinLogMel = kMinLogHz / kFsp;
constexpr double kLogStep = 0.06875177742094923;

double hz_to_mel(double hz) { return hz >= kMinLogHz ? kMinLogMel + std::log(hz / kMinLogHz) / kLogStep : hz / kFsp; }
double mel_to_hz(double mel) {
    return mel >= kMinLogMel ? kMinLogHz * std::exp(kLogStep * (mel - kMinLogMel)) : std::fma(kFsp, mel, 0.0);
}

struct Csr {
    HostTables &t;
    void open_row() { t.row_ptr.push_back(static_cast<int>(t.val.size())); }
    // SparseMatrix::set (:69-87): entries with |v| <= 1e-10 are dropped
    void set(size_t col, double v) {
        if (std::fabs(v) > 1e-10) { t.col.push_back(static_cast<int>(col)); t.val.push_back(v); }
    }
    void close() { t.row_ptr.push_back(static_cast<int>(t.val.size())); }
};

// build_mel_filterbank_matrix (:2302-2432) + mel_band_centres_hz (:2510-2530)
void build_mel(const sgx_plan_desc &d, HostTables &t) {
    const double sr = d.sample_rate_hz;
    if (d.f_min < 0.0 || std::isinf(d.f_min)) invalid("f_min must be >= 0");
    if (d.f_max <= d.f_min) invalid("f_max must be > f_min");
    if (d.f_max > sr * 0.5) invalid("f_max must be <= Nyquist");
    const size_t n_mels = d.n_bands;
    const double df = sr / static_cast<double>(d.n_fft);
    const double mel_min = hz_to_mel(d.f_min), mel_max = hz_to_mel(d.f_max);
    const size_t n_points = n_mels + 2;
    const double step = (mel_max - mel_min) / static_cast<double>(n_points - 1);
    std::vector<double> mel_pts(n_points), hz_pts(n_points);
    for (size_t i = 0; i < n_points; ++i) mel_pts[i] = std::fma(static_cast<double>(i), step, mel_min);
    for (size_t i = 0; i < n_points; ++i) hz_pts[i] = mel_to_hz(mel_pts[i]);

    Csr csr{t};
    for (size_t m = 0; m < n_mels; ++m) {
        csr.open_row();
        const double left = hz_pts[m], centre = hz_pts[m + 1], right = hz_pts[m + 2];
        const double dl = centre - left, dr = right - centre;
        if (dl == 0.0 || dr == 0.0) continue;   // degenerate triangle
        for (size_t k = 0; k < t.out_len; ++k) {
            const double f = static_cast<double>(k) * df;
            double wgt = std::fmin((f - left) / dl, (right - f) / dr);
            if (wgt < 0.0) wgt = 0.0;
            if (wgt > 1.0) wgt = 1.0;
            if (wgt > 0.0) csr.set(k, wgt);
        }
    }
    csr.close();
    for (size_t m = 0; m < n_mels; ++m) {
        const int s = t.row_ptr[m], e = t.row_ptr[m + 1];
        double scale = 1.0;
        bool apply = false;
        if (d.mel_norm == SGX_MELNORM_SLANEY) {
            scale = 2.0 / (mel_to_hz(mel_pts[m + 2]) - mel_to_hz(mel_pts[m]));
            apply = true;
        } else if (d.mel_norm == SGX_MELNORM_L1) {
            double sum = 0.0;
            for (int i = s; i < e; ++i) sum += t.val[i];
            if (sum > 0.0) { scale = 1.0 / sum; apply = true; }
        } else if (d.mel_norm == SGX_MELNORM_L2) {
            double sum = 0.0;
            for (int i = s; i < e; ++i) sum += t.val[i] * t.val[i];
            const double nv = std::sqrt(sum);
            if (nv > 0.0) { scale = 1.0 / nv; apply = true; }
        }
        if (apply)
            for (int i = s; i < e; ++i) t.val[i] *= scale;
    }
    // frequency axis ignores MelParams f_min/f_max (:1924-1931)
    t.freq_axis.resize(n_mels);
    const double f_hi = std::fmin(sr * 0.5, sr * 0.5);
    const double a_min = hz_to_mel(0.0), a_max = hz_to_mel(f_hi);
    const double a_step = (a_max - a_min) / static_cast<double>(n_mels + 1);
    for (size_t i = 0; i < n_mels; ++i) t.freq_axis[i] = mel_to_hz(std::fma(static_cast<double>(i) + 1.0, a_step, a_min));
}

size_t saturating_usize(double v) {   // Rust `as usize`
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return std::numeric_limits<size_t>::max();
    return static_cast<size_t>(v);
}

// build_loghz_matrix (:2438-2508)
void build_loghz(const sgx_plan_desc &d, HostTables &t) {
    const double sr = d.sample_rate_hz;
    if (d.f_min <= 0.0 || std::isinf(d.f_min)) invalid("f_min must be finite and > 0");
    if (d.f_max <= d.f_min) invalid("f_max must be > f_min");
    if (d.f_max > sr * 0.5) invalid("f_max must be <= Nyquist");
    const size_t n_bins = d.n_bands;
    const double df = sr / static_cast<double>(d.n_fft);
    const double lo = std::log(d.f_min), hi = std::log(d.f_max);
    const double step = (hi - lo) / static_cast<double>(n_bins - 1);
    t.freq_axis.resize(n_bins);
    for (size_t i = 0; i < n_bins; ++i) t.freq_axis[i] = std::exp(std::fma(static_cast<double>(i), step, lo));
    Csr csr{t};
    for (size_t i = 0; i < n_bins; ++i) {
        csr.open_row();
        const double exact = t.freq_axis[i] / df;
        const size_t lower = saturating_usize(std::floor(exact));
        size_t upper = saturating_usize(std::ceil(exact));
        if (upper > t.out_len - 1) upper = t.out_len - 1;
        if (lower >= t.out_len) continue;
        if (lower == upper) {
            csr.set(lower, 1.0);
        } else {
            const double frac = exact - static_cast<double>(lower);
            csr.set(lower, 1.0 - frac);
            if (upper < t.out_len) csr.set(upper, frac);
        }
    }
    csr.close();
}

// ErbFilterbank::generate (src/erb.rs:266-332)
void build_erb(const sgx_plan_desc &d, HostTables &t) {
    const size_t nf = d.n_bands;
    std::vector<double> cf(nf);
    if (d.erb_spacing == SGX_ERB_LINEAR) {
        const double e_min = 24.7 * (4.37 * d.f_min / 1000.0 + 1.0);
        const double e_max = 24.7 * (4.37 * d.f_max / 1000.0 + 1.0);
        const double e_step = (e_max - e_min) / static_cast<double>(nf - 1);
        for (size_t i = 0; i < nf; ++i) cf[i] = (std::fma(static_cast<double>(i), e_step, e_min) / 24.7 - 1.0) * 1000.0 / 4.37;
    } else {   // apple_tr35_center_freqs (src/erb.rs:221-236), reversed to low -> high
        const double shift = 9.26449 * 24.7;
        const double e = (std::log(d.f_min + shift) - std::log(d.f_max + shift)) / static_cast<double>(nf);
        for (size_t i = 0; i < nf; ++i) cf[nf - 1 - i] = -shift + std::exp((static_cast<double>(i) + 1.0) * e) * (d.f_max + shift);
    }
    const double res = d.sample_rate_hz / static_cast<double>(d.n_fft);
    t.dense.resize(nf * t.out_len);
    for (size_t f = 0; f < nf; ++f) {
        const double bw = 1.019 * (24.7 * (4.37 * cf[f] / 1000.0 + 1.0));
        for (size_t k = 0; k < t.out_len; ++k) {
            const double x = (static_cast<double>(k) * res - cf[f]) / bw;
            // (1 + jx)^2, then squared again, with num_complex's (ac - bd, ad + bc)
            const double r2 = 1.0 * 1.0 - x * x, i2 = 1.0 * x + x * 1.0;
            const double r4 = r2 * r2 - i2 * i2, i4 = r2 * i2 + i2 * r2;
            t.dense[f * t.out_len + k] = 1.0 / (r4 * r4 + i4 * i4);
        }
    }
    t.freq_axis = cf;
}

}  // namespace

// ChromaParams::new (src/chroma.rs:81-112)
void validate_chroma(double sample_rate_hz, double tuning, double f_min, double f_max) {
    if (!(sample_rate_hz > 0.0 && std::isfinite(sample_rate_hz))) invalid("sample_rate must be finite and > 0");   // :286-290
    if (!(tuning > 0.0 && std::isfinite(tuning))) invalid("tuning must be finite and > 0");
    if (!(f_min > 0.0 && std::isfinite(f_min))) invalid("f_min must be finite and > 0");
    if (f_max <= f_min) invalid("f_max must be > f_min");
}

// build_chroma_filterbank (src/chroma.rs:279-346): Gaussian (sigma = 1 semitone) weights on the circular pitch-class
// distance of every FFT bin inside [f_min, f_max], rows normalised to unit sum. dense: [12][n_fft/2 + 1].
void build_chroma_filterbank(double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max,
                             std::vector<double> &dense) {
    const size_t n_bins = n_fft / 2 + 1;
    const double freq_resolution = sample_rate_hz / static_cast<double>(n_fft);
    const double ln2 = 0.693147180559945309417232121458176568;     // std::f64::consts::LN_2
    dense.assign(12 * n_bins, 0.0);
    for (size_t bin = 0; bin < n_bins; ++bin) {
        const double freq = static_cast<double>(bin) * freq_resolution;
        if (freq < f_min || freq > f_max || freq <= 0.0) continue;
        const double midi_note = 69.0 + 12.0 * std::log(freq / tuning) / ln2;
        double pitch_class = std::fmod(midi_note, 12.0);           // f64::rem_euclid
        if (pitch_class < 0.0) pitch_class += 12.0;
        for (size_t c = 0; c < 12; ++c) {
            const double dist = std::fabs(pitch_class - static_cast<double>(c));
            const double circular = std::fmin(dist, 12.0 - dist);
            const double q = circular / 1.0;
            dense[c * n_bins + bin] = std::exp(-0.5 * (q * q));
        }
    }
    for (size_t c = 0; c < 12; ++c) {
        double row_sum = 0.0;
        for (size_t i = 0; i < n_bins; ++i) row_sum += dense[c * n_bins + i];
        if (row_sum > 0.0)
            for (size_t i = 0; i < n_bins; ++i) dense[c * n_bins + i] /= row_sum;
    }
}

size_t frame_count(size_t n_samples, size_t n_fft, size_t hop, bool centre) {
    const size_t pad = centre ? n_fft / 2 : 0;
    const size_t padded = n_samples + 2 * pad;
    if (padded < n_fft) return 1;
    return (padded - n_fft) / hop + 1;
}

void validate_desc(const sgx_plan_desc &d) {
    if (d.dtype != SGX_F32 && d.dtype != SGX_F64) invalid("dtype must be f32 or f64");
    if (d.n_fft == 0) invalid("n_fft must be set");                                   // NonZeroUsize / :3686
    if (d.hop_size == 0) invalid("hop_size must be set");                             // :3689
    if (d.hop_size > d.n_fft) invalid("hop_size must be <= n_fft");                   // :3485
    if (d.window < SGX_WIN_RECTANGULAR || d.window > SGX_WIN_CUSTOM) invalid("unknown window type");
    if (d.window == SGX_WIN_CUSTOM) {
        if (d.custom_window == nullptr || d.custom_window_len == 0) invalid("Custom window coefficients cannot be empty");
        if (d.custom_window_len != d.n_fft) {                                         // :3490-3497
            char buf[160];
            std::snprintf(buf, sizeof buf, "Custom window size (%zu) must match n_fft (%zu)", d.custom_window_len, d.n_fft);
            invalid(buf);
        }
    }
    if (!(d.sample_rate_hz > 0.0 && std::isfinite(d.sample_rate_hz))) invalid("sample_rate_hz must be finite and > 0");   // :4130
    const double nyquist = d.sample_rate_hz * 0.5;
    char buf[200];
    switch (d.mapping) {
        case SGX_MAP_LINEAR:
            break;
        case SGX_MAP_MEL:
            if (d.n_bands == 0) invalid("n_mels must be non-zero");
            if (d.f_min < 0.0) invalid("f_min must be >= 0");                         // MelParams::with_norm :3799
            if (d.f_max <= d.f_min) invalid("f_max must be > f_min");                 // :3803
            if (d.f_max > nyquist) invalid("mel f_max must be <= Nyquist");           // mel_plan :954-959
            if (d.n_bands > 10000) invalid("n_mels is unreasonably large");           // :1696
            break;
        case SGX_MAP_ERB:
            if (d.n_bands < 2) invalid("n_filters must be >= 2 (single filter would cause division by zero)");   // src/erb.rs:67-71
            if (d.f_min < 0.0 || std::isinf(d.f_min)) invalid("f_min must be finite and >= 0");
            if (d.f_max <= d.f_min) invalid("f_max must be > f_min");
            if (d.f_max > nyquist) {                                                  // erb_plan :1016-1022
                std::snprintf(buf, sizeof buf, "f_max=%g exceeds Nyquist=%g", d.f_max, nyquist);
                invalid(buf);
            }
            if (d.n_bands > 10000) invalid("n_filters is unreasonably large");        // :1769
            break;
        case SGX_MAP_LOGHZ:
            if (d.n_bands == 0) invalid("n_bins must be non-zero");
            if (!(d.f_min > 0.0 && std::isfinite(d.f_min))) invalid("f_min must be finite and > 0");   // LogHzParams::new :3961
            if (d.f_max <= d.f_min) invalid("f_max must be > f_min");
            if (d.f_max > nyquist) {                                                  // log_hz_plan :1078-1084
                std::snprintf(buf, sizeof buf, "f_max=%g exceeds Nyquist=%g", d.f_max, nyquist);
                invalid(buf);
            }
            if (d.n_bands > 10000) invalid("n_bins is unreasonably large");           // :1732
            break;
        case SGX_MAP_CHROMA:
            validate_chroma(d.sample_rate_hz, d.chroma_tuning, d.f_min, d.f_max);
            if (d.chroma_norm < SGX_CHROMANORM_NONE || d.chroma_norm > SGX_CHROMANORM_MAX) invalid("unknown chroma normalisation");
            // chromagram() always runs on Spectrogram::<LinearHz, Magnitude, T> with no dB stage (src/chroma.rs:493-499)
            if (d.amp != SGX_AMP_MAGNITUDE || d.has_floor_db) invalid("chroma plans take the magnitude spectrogram (amp = magnitude, no dB floor)");
            if (d.output != SGX_OUT_SPECTROGRAM) invalid("chroma plans produce a (12, n_frames) matrix (output = spectrogram)");
            break;
        default:
            invalid("unknown frequency mapping");
    }
    if (d.amp < SGX_AMP_POWER || d.amp > SGX_AMP_DECIBELS) invalid("unknown amplitude scale");
    if (d.has_floor_db && !std::isfinite(d.floor_db)) invalid("floor_db must be finite");   // LogParams::new :4072
    if (d.output == SGX_OUT_MFCC) {
        if (d.mapping != SGX_MAP_MEL) invalid("MFCC output requires a mel mapping");
        if (d.n_mfcc == 0) invalid("n_mfcc must be non-zero");
        if (d.n_mfcc > d.n_bands) invalid("n_mfcc must be <= n_mels");                // src/mfcc.rs:231-233
    } else if (d.output != SGX_OUT_SPECTROGRAM && d.output != SGX_OUT_COMPLEX_STFT) {
        invalid("unknown output kind");
    }
}

void build_dct(size_t n_mfcc, size_t n_mels, size_t lifter, std::vector<double> &basis, std::vector<double> &lift) {
    basis.resize(n_mfcc * n_mels);
    for (size_t k = 0; k < n_mfcc; ++k)
        for (size_t i = 0; i < n_mels; ++i)   // src/mfcc.rs:285-286
            basis[k * n_mels + i] =
                std::cos(kPi * static_cast<double>(k) * (static_cast<double>(i) + 0.5) / static_cast<double>(n_mels));
    lift.assign(n_mfcc, 1.0);
    if (lifter > 0)
        for (size_t i = 0; i < n_mfcc; ++i)   // src/mfcc.rs:304-307
            lift[i] = std::fma(static_cast<double>(lifter) / 2.0,
                               std::sin(kPi * static_cast<double>(i) / static_cast<double>(lifter)), 1.0);
}

void build_tables(const sgx_plan_desc &d, HostTables &t) {
    t = HostTables{};
    t.out_len = d.n_fft / 2 + 1;   // r2c_output_size, src/fft_backend.rs:16-18
    make_window(d, t.window);
    switch (d.mapping) {
        case SGX_MAP_LINEAR: {
            t.n_bins = t.out_len;
            t.freq_axis.resize(t.out_len);
            const double df = d.sample_rate_hz / static_cast<double>(d.n_fft);   // :1911-1922
            for (size_t k = 0; k < t.out_len; ++k) t.freq_axis[k] = static_cast<double>(k) * df;
            break;
        }
        case SGX_MAP_MEL:
            t.n_bins = d.n_bands;
            build_mel(d, t);
            break;
        case SGX_MAP_LOGHZ:
            t.n_bins = d.n_bands;
            build_loghz(d, t);
            break;
        case SGX_MAP_ERB:
            t.n_bins = d.n_bands;
            build_erb(d, t);
            break;
        case SGX_MAP_CHROMA:
            t.n_bins = 12;
            build_chroma_filterbank(d.sample_rate_hz, d.n_fft, d.chroma_tuning, d.f_min, d.f_max, t.dense);
            t.freq_axis.resize(12);
            for (size_t c = 0; c < 12; ++c) t.freq_axis[c] = static_cast<double>(c);   // pitch classes C .. B (Chromagram::labels, src/chroma.rs:238-242)
            break;
    }
    if (d.output == SGX_OUT_MFCC) build_dct(d.n_mfcc, d.n_bands, d.lifter, t.dct, t.lifter);
}

}  // namespace sgx
