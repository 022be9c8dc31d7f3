/*
 * oracle.c -- CPU restatement of the jmg049/Spectrograms hot path (spectrograms crate v2.1.0).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product (spectrograms_b200/) never links or calls it.
 *
 * Parity pinning status: PINNED for the north_star path -- STFT, window, power / magnitude / dB, the mel, ERB and
 * LogHz filterbanks and their spectrograms, the DCT-II / lifter and MFCC. The reference crate cannot be compiled here
 * (no cargo/rustc; realfft 3.5.0 / rustfft 6.4.1 are un-vendored crates.io dependencies, Cargo.lock:920-922,1002-1004)
 * and ships no golden vectors, so the pins are
 *   (1) every known-answer / shape / property assertion the reference's own tests make for this path
 *       (tests/test_oracle_pinning.py lists them with file:line);
 *   (2) outputs of the reference's own NumPy restatement python/examples/numpy_impls.py, imported unmodified from
 *       /root/reference by tests/golden/make_golden.py -> tests/golden/ref_numpy_impls.npz: stft, hann_window, power,
 *       magnitude, dB, erb_centers + gammatone_response + erb_spectrogram (= src/erb.rs:266-402) and
 *       log_frequency_matrix + logfreq_spectrogram (= src/spectrogram.rs:2438-2508);
 *   (3) third-party implementations of the algorithms the crate names -> tests/golden/third_party_pins.npz:
 *       torchaudio.functional.melscale_fbanks(mel_scale="slaney", norm=None|"slaney") for the librosa-style mel
 *       filterbank (src/spectrogram.rs:2262-2432; exact 394 / 2018 non-zero pattern, weights to torchaudio's own f32
 *       rounding) and scipy.fft.dct(type=2)/2 for dct_ii (src/mfcc.rs:278-292);
 *   tests/test_reference_pins.py checks the oracle, the product's host tables AND (under -m gpu) the CUDA output
 *   directly against (2) and (3).
 * STILL UNPINNED ("parity unpinned"): the chroma filterbank and the binaural cues -- the reference ships no vectors, its
 * NumPy chroma is a different (hard-assignment) algorithm and no third-party implementation is installed; they are
 * restated line by line, cross-checked by oracle/oracle_np.py and by derived known answers. The inverse path is
 * checked against SciPy's pocketfft irfft and by round trips. L1 / L2 mel norms and the Apple-TR35 ERB spacing are
 * covered by the two-restatement cross-check only.
 *
 * All citations are relative to the reference checkout (src/spectrogram.rs unless a file is named).
 * Compile with -ffp-contract=off: the reference (rustc) never contracts a*b+c unless mul_add is written.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* ---------------------------------------------------------------------------------------------- */
/* frame_count  (src/spectrogram.rs:1230-1250)                                                      */
size_t orc_frame_count(size_t n_samples, size_t n_fft, size_t hop, int centre) {
    const size_t pad = centre ? n_fft / 2 : 0;
    const size_t padded = n_samples + 2 * pad;
    if (padded < n_fft) return 1;
    return (padded - n_fft) / hop + 1;
}

/* modified_bessel_i0 (:2237-2259), Abramowitz & Stegun polynomial */
static double bessel_i0(double x) {
    const double ax = fabs(x);
    if (ax <= 3.75) {
        const double t = x / 3.75;
        const double t2 = t * t;
        return 1.0 + t2 * (3.5156229 + t2 * (3.0899424 + t2 * (1.2067492 + t2 * (0.2659732 + t2 * (0.0360768 + t2 * 0.0045813)))));
    }
    const double t = 3.75 / ax;
    const double poly = 0.39894228 + t * (0.01328592 + t * (0.00225319 + t * (-0.00157565 + t * (0.00916281
                       + t * (-0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377)))))));
    return (exp(ax) / (sqrt(ax) * sqrt(2.0 * M_PI))) * poly;
}

/* make_window (:2159-2235), f64 */
static int make_window(const orc_desc *d, double *w, char *err, size_t errlen) {
    const size_t n = d->n_fft;
    switch (d->window_kind) {
    case ORC_WIN_RECT:
        for (size_t i = 0; i < n; ++i) w[i] = 1.0;
        break;
    case ORC_WIN_HANN: {
        const double n1 = (double)(n - 1);
        for (size_t i = 0; i < n; ++i) w[i] = fma(0.5, -cos(2.0 * M_PI * (double)i / n1), 0.5);
        break;
    }
    case ORC_WIN_HAMMING: {
        const double n1 = (double)(n - 1);
        for (size_t i = 0; i < n; ++i) w[i] = fma(0.46, -cos(2.0 * M_PI * (double)i / n1), 0.54);
        break;
    }
    case ORC_WIN_BLACKMAN: {
        const double n1 = (double)(n - 1);
        for (size_t i = 0; i < n; ++i) {
            const double a = 2.0 * M_PI * (double)i / n1;
            w[i] = fma(0.08, cos(2.0 * a), fma(0.5, -cos(a), 0.42));
        }
        break;
    }
    case ORC_WIN_KAISER:
        if (n == 1) { w[0] = 1.0; break; }
        {
            const double beta = d->window_param;
            const double denom = bessel_i0(beta);
            const double n_max = (double)(n - 1) / 2.0;
            for (size_t i = 0; i < n; ++i) {
                const double nn = (double)i - n_max;
                double ratio = 0.0;
                if (n_max != 0.0) {
                    const double nrm = nn / n_max;
                    ratio = fmax(1.0 - nrm * nrm, 0.0);
                }
                const double arg = beta * sqrt(ratio);
                w[i] = (denom == 0.0) ? 0.0 : bessel_i0(arg) / denom;
            }
        }
        break;
    case ORC_WIN_GAUSSIAN: {
        const double centre = (double)(n - 1) / 2.0;
        for (size_t i = 0; i < n; ++i) {
            const double q = ((double)i - centre) / d->window_param;
            w[i] = exp(-0.5 * (q * q));          /* powi(2) == q*q */
        }
        break;
    }
    case ORC_WIN_CUSTOM:
        if (!d->custom_window || d->custom_window_len != n) {
            snprintf(err, errlen, "Custom window size (%zu) must match n_fft (%zu)", d->custom_window_len, n);
            return 1;
        }
        memcpy(w, d->custom_window, sizeof(double) * n);
        break;
    default:
        snprintf(err, errlen, "unknown window kind");
        return 1;
    }
    return 0;
}

/* hz_to_mel / mel_to_hz (:2268-2300), Slaney */
static double hz_to_mel(double hz) {
    const double F_SP = 200.0 / 3.0, MIN_LOG_HZ = 1000.0, MIN_LOG_MEL = (1000.0 - 0.0) / (200.0 / 3.0);
    const double LOGSTEP = 0.06875177742094923;
    if (hz >= MIN_LOG_HZ) return MIN_LOG_MEL + log(hz / MIN_LOG_HZ) / LOGSTEP;
    return (hz - 0.0) / F_SP;
}
static double mel_to_hz(double mel) {
    const double F_SP = 200.0 / 3.0, MIN_LOG_HZ = 1000.0, MIN_LOG_MEL = (1000.0 - 0.0) / (200.0 / 3.0);
    const double LOGSTEP = 0.06875177742094923;
    if (mel >= MIN_LOG_MEL) return MIN_LOG_HZ * exp(LOGSTEP * (mel - MIN_LOG_MEL));
    return fma(F_SP, mel, 0.0);
}

typedef struct { size_t *row_ptr, *col; double *val; size_t cap, nnz; } csr_builder;
static void csr_push(csr_builder *b, size_t col, double v) {
    /* SparseMatrix::set (:69-87): keep only |v| > 1e-10 */
    if (!(fabs(v) > 1e-10)) return;
    if (b->nnz == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 1024;
        b->col = (size_t *)realloc(b->col, sizeof(size_t) * b->cap);
        b->val = (double *)realloc(b->val, sizeof(double) * b->cap);
    }
    b->col[b->nnz] = col; b->val[b->nnz] = v; b->nnz++;
}

/* build_mel_filterbank_matrix (:2302-2432) */
static int build_mel(const orc_desc *d, size_t out_len, orc_tables *t, char *err, size_t errlen) {
    const double sr = d->sample_rate, f_min = d->f_min, f_max = d->f_max;
    if (f_min < 0.0 || isinf(f_min)) { snprintf(err, errlen, "f_min must be >= 0"); return 1; }
    if (f_max <= f_min) { snprintf(err, errlen, "f_max must be > f_min"); return 1; }
    if (f_max > sr * 0.5) { snprintf(err, errlen, "f_max must be <= Nyquist"); return 1; }
    const size_t n_mels = d->n_bands;
    const double df = sr / (double)d->n_fft;
    const double mel_min = hz_to_mel(f_min), mel_max = hz_to_mel(f_max);
    const size_t n_points = n_mels + 2;
    const double step = (mel_max - mel_min) / (double)(n_points - 1);
    double *mel_points = (double *)malloc(sizeof(double) * n_points);
    double *hz_points = (double *)malloc(sizeof(double) * n_points);
    for (size_t i = 0; i < n_points; ++i) mel_points[i] = fma((double)i, step, mel_min);
    for (size_t i = 0; i < n_points; ++i) hz_points[i] = mel_to_hz(mel_points[i]);
    csr_builder b; memset(&b, 0, sizeof(b));
    b.row_ptr = (size_t *)calloc(n_mels + 1, sizeof(size_t));
    for (size_t m = 0; m < n_mels; ++m) {
        b.row_ptr[m] = b.nnz;
        const double fl = hz_points[m], fc = hz_points[m + 1], fr = hz_points[m + 2];
        const double dl = fc - fl, dr = fr - fc;
        if (dl == 0.0 || dr == 0.0) continue;
        for (size_t k = 0; k < out_len; ++k) {
            const double bf = (double)k * df;
            const double lower = (bf - fl) / dl;
            const double upper = (fr - bf) / dr;
            /* f64::min then clamp(0,1) */
            double wgt = fmin(lower, upper);
            if (wgt < 0.0) wgt = 0.0;
            if (wgt > 1.0) wgt = 1.0;
            if (wgt > 0.0) csr_push(&b, k, wgt);
        }
    }
    b.row_ptr[n_mels] = b.nnz;
    /* normalisation (:2385-2429), applied to the stored values */
    for (size_t m = 0; m < n_mels; ++m) {
        const size_t s = b.row_ptr[m], e = b.row_ptr[m + 1];
        if (d->mel_norm == ORC_MELNORM_SLANEY) {
            const double enorm = 2.0 / (mel_to_hz(mel_points[m + 2]) - mel_to_hz(mel_points[m]));
            for (size_t i = s; i < e; ++i) b.val[i] *= enorm;
        } else if (d->mel_norm == ORC_MELNORM_L1) {
            double sum = 0.0;
            for (size_t i = s; i < e; ++i) sum += b.val[i];
            if (sum > 0.0) { const double nz = 1.0 / sum; for (size_t i = s; i < e; ++i) b.val[i] *= nz; }
        } else if (d->mel_norm == ORC_MELNORM_L2) {
            double sum = 0.0;
            for (size_t i = s; i < e; ++i) sum += b.val[i] * b.val[i];
            const double nv = sqrt(sum);
            if (nv > 0.0) { const double nz = 1.0 / nv; for (size_t i = s; i < e; ++i) b.val[i] *= nz; }
        }
    }
    t->row_ptr = b.row_ptr; t->col = b.col; t->val = b.val; t->nnz = b.nnz;
    /* axis: mel_band_centres_hz (:2510-2530) -- always 0..Nyquist (quirk F9, :1924-1931) */
    t->freq_axis = (double *)malloc(sizeof(double) * n_mels);
    {
        const double fmx = fmin(sr * 0.5, sr * 0.5);
        const double mmin = hz_to_mel(0.0), mmax = hz_to_mel(fmx);
        const double st = (mmax - mmin) / (double)(n_mels + 1);
        for (size_t i = 0; i < n_mels; ++i) t->freq_axis[i] = mel_to_hz(fma((double)i + 1.0, st, mmin));
    }
    free(mel_points); free(hz_points);
    return 0;
}

/* build_loghz_matrix (:2438-2508) */
static int build_loghz(const orc_desc *d, size_t out_len, orc_tables *t, char *err, size_t errlen) {
    const double sr = d->sample_rate, f_min = d->f_min, f_max = d->f_max;
    if (f_min <= 0.0 || isinf(f_min)) { snprintf(err, errlen, "f_min must be finite and > 0"); return 1; }
    if (f_max <= f_min) { snprintf(err, errlen, "f_max must be > f_min"); return 1; }
    if (f_max > sr * 0.5) { snprintf(err, errlen, "f_max must be <= Nyquist"); return 1; }
    const size_t n_bins = d->n_bands;
    const double df = sr / (double)d->n_fft;
    const double lmin = log(f_min), lmax = log(f_max);
    const double lstep = (lmax - lmin) / (double)(n_bins - 1);
    t->freq_axis = (double *)malloc(sizeof(double) * n_bins);
    for (size_t i = 0; i < n_bins; ++i) t->freq_axis[i] = exp(fma((double)i, lstep, lmin));
    csr_builder b; memset(&b, 0, sizeof(b));
    b.row_ptr = (size_t *)calloc(n_bins + 1, sizeof(size_t));
    for (size_t i = 0; i < n_bins; ++i) {
        b.row_ptr[i] = b.nnz;
        const double exact = t->freq_axis[i] / df;
        const double fl = floor(exact), ce = ceil(exact);
        /* `as usize` saturates: negative/NaN -> 0, huge -> usize::MAX */
        const size_t lower = (fl >= 1.8446744073709552e19) ? SIZE_MAX : (fl > 0.0 ? (size_t)fl : 0);
        size_t upper = (ce >= 1.8446744073709552e19) ? SIZE_MAX : (ce > 0.0 ? (size_t)ce : 0);
        if (upper > out_len - 1) upper = out_len - 1;
        if (lower >= out_len) continue;
        if (lower == upper) {
            csr_push(&b, lower, 1.0);
        } else {
            const double frac = exact - (double)lower;
            csr_push(&b, lower, 1.0 - frac);
            if (upper < out_len) csr_push(&b, upper, frac);
        }
    }
    b.row_ptr[n_bins] = b.nnz;
    t->row_ptr = b.row_ptr; t->col = b.col; t->val = b.val; t->nnz = b.nnz;
    return 0;
}

/* ErbFilterbank::generate (src/erb.rs:266-332) */
static int build_erb(const orc_desc *d, size_t out_len, orc_tables *t, char *err, size_t errlen) {
    const size_t nf = d->n_bands;
    if (nf < 2) { snprintf(err, errlen, "n_filters must be >= 2 (single filter would cause division by zero)"); return 1; }
    if (d->f_min < 0.0 || isinf(d->f_min)) { snprintf(err, errlen, "f_min must be finite and >= 0"); return 1; }
    if (d->f_max <= d->f_min) { snprintf(err, errlen, "f_max must be > f_min"); return 1; }
    double *cf = (double *)malloc(sizeof(double) * nf);
    if (d->erb_spacing == ORC_ERB_LINEAR) {
        /* hz_to_erb :208-210 / erb_to_hz :249-251 */
        const double emin = 24.7 * (4.37 * d->f_min / 1000.0 + 1.0);
        const double emax = 24.7 * (4.37 * d->f_max / 1000.0 + 1.0);
        const double estep = (emax - emin) / (double)(nf - 1);
        for (size_t i = 0; i < nf; ++i) {
            const double e = fma((double)i, estep, emin);
            cf[i] = (e / 24.7 - 1.0) * 1000.0 / 4.37;
        }
    } else {
        /* apple_tr35_center_freqs :221-236 */
        const double shift = 9.26449 * 24.7;
        const double a = -shift, dd = d->f_max + shift;
        const double e = (log(d->f_min + shift) - log(d->f_max + shift)) / (double)nf;
        for (size_t i = 0; i < nf; ++i) cf[nf - 1 - i] = a + exp(((double)i + 1.0) * e) * dd;
    }
    const double res = d->sample_rate / (double)d->n_fft;
    t->dense = (double *)malloc(sizeof(double) * nf * out_len);
    for (size_t f = 0; f < nf; ++f) {
        const double bw = 1.019 * (24.7 * (4.37 * cf[f] / 1000.0 + 1.0));
        for (size_t k = 0; k < out_len; ++k) {
            const double freq = (double)k * res;
            /* denom = 1 + j x ; denom^2 ; denom^4 ; 1/|denom^4|^2, with num_complex Mul: (ac-bd, ad+bc) */
            const double x = (freq - cf[f]) / bw;
            const double re1 = 1.0, im1 = x;
            const double re2 = re1 * re1 - im1 * im1, im2 = re1 * im1 + im1 * re1;
            const double re4 = re2 * re2 - im2 * im2, im4 = re2 * im2 + im2 * re2;
            t->dense[f * out_len + k] = 1.0 / (re4 * re4 + im4 * im4);
        }
    }
    t->freq_axis = cf;
    return 0;
}

int orc_build_tables(const orc_desc *d, orc_tables *t, char *err, size_t errlen) {
    memset(t, 0, sizeof(*t));
    if (d->n_fft == 0 || d->hop == 0) { snprintf(err, errlen, "n_fft and hop_size must be non-zero"); return 1; }
    if (d->hop > d->n_fft) { snprintf(err, errlen, "hop_size must be <= n_fft"); return 1; }              /* :3485 */
    if (!(d->sample_rate > 0.0 && isfinite(d->sample_rate))) {                                           /* :4130 */
        snprintf(err, errlen, "sample_rate_hz must be finite and > 0"); return 1;
    }
    const size_t out_len = d->n_fft / 2 + 1;                    /* r2c_output_size, src/fft_backend.rs:16-18 */
    t->out_len = out_len;
    t->window = (double *)malloc(sizeof(double) * d->n_fft);
    if (make_window(d, t->window, err, errlen)) { free(t->window); t->window = NULL; return 1; }
    int rc = 0;
    switch (d->mapping) {
    case ORC_MAP_IDENTITY:
        t->n_bins = out_len;
        t->freq_axis = (double *)malloc(sizeof(double) * out_len);
        {   /* :1911-1922 */
            const double df = d->sample_rate / (double)d->n_fft;
            for (size_t k = 0; k < out_len; ++k) t->freq_axis[k] = (double)k * df;
        }
        break;
    case ORC_MAP_MEL:
        if (d->n_bands == 0 || d->n_bands > 10000) { snprintf(err, errlen, "n_mels is unreasonably large"); rc = 1; break; }
        if (d->f_max > d->sample_rate * 0.5) { snprintf(err, errlen, "mel f_max must be <= Nyquist"); rc = 1; break; }   /* :954-959 */
        t->n_bins = d->n_bands;
        rc = build_mel(d, out_len, t, err, errlen);
        break;
    case ORC_MAP_LOGHZ:
        if (d->n_bands == 0 || d->n_bands > 10000) { snprintf(err, errlen, "n_bins is unreasonably large"); rc = 1; break; }
        if (d->f_max > d->sample_rate * 0.5) { snprintf(err, errlen, "f_max=%g exceeds Nyquist=%g", d->f_max, d->sample_rate * 0.5); rc = 1; break; }
        t->n_bins = d->n_bands;
        rc = build_loghz(d, out_len, t, err, errlen);
        break;
    case ORC_MAP_ERB:
        if (d->n_bands == 0 || d->n_bands > 10000) { snprintf(err, errlen, "n_filters is unreasonably large"); rc = 1; break; }
        if (d->f_max > d->sample_rate * 0.5) { snprintf(err, errlen, "f_max=%g exceeds Nyquist=%g", d->f_max, d->sample_rate * 0.5); rc = 1; break; }
        t->n_bins = d->n_bands;
        rc = build_erb(d, out_len, t, err, errlen);
        break;
    default:
        snprintf(err, errlen, "unknown mapping"); rc = 1;
    }
    if (rc) { orc_free_tables(t); return rc; }
    if (d->amp == ORC_AMP_DECIBELS && d->has_floor_db && !isfinite(d->floor_db)) {                        /* :2021, :4072 */
        snprintf(err, errlen, "floor_db must be finite"); orc_free_tables(t); return 1;
    }
    return 0;
}

void orc_free_tables(orc_tables *t) {
    free(t->window); free(t->row_ptr); free(t->col); free(t->val); free(t->dense); free(t->freq_axis);
    memset(t, 0, sizeof(*t));
}

/* ---------------------------------------------------------------------------------------------- */
#define FN(name) name##_f32
#define REAL float
#define R_SQRT sqrtf
#define R_LOG10 log10f
#define R_FMAX fmaxf
#define R_FMA fmaf
#include "oracle_impl.inc"
#undef FN
#undef REAL
#undef R_SQRT
#undef R_LOG10
#undef R_FMAX
#undef R_FMA

#define FN(name) name##_f64
#define REAL double
#define R_SQRT sqrt
#define R_LOG10 log10
#define R_FMAX fmax
#define R_FMA fma
#include "oracle_impl.inc"
#undef FN
#undef REAL
#undef R_SQRT
#undef R_LOG10
#undef R_FMAX
#undef R_FMA

/* ---------------------------------------------------------------------------------------------- */
/* Exported API                                                                                     */
struct orc_plan { int dtype; plan_f32 *p32; plan_f64 *p64; orc_desc d; double *custom; };

static __thread char g_err[512];
const char *orc_last_error(void) { return g_err; }

orc_plan *orc_plan_create(const orc_desc *d) {
    g_err[0] = 0;
    orc_plan *p = (orc_plan *)calloc(1, sizeof(*p));
    p->dtype = d->dtype;
    p->d = *d;
    if (d->custom_window && d->custom_window_len) {
        p->custom = (double *)malloc(sizeof(double) * d->custom_window_len);
        memcpy(p->custom, d->custom_window, sizeof(double) * d->custom_window_len);
        p->d.custom_window = p->custom;
    }
    if (d->dtype == ORC_F32) p->p32 = plan_new_f32(&p->d, g_err, sizeof(g_err));
    else p->p64 = plan_new_f64(&p->d, g_err, sizeof(g_err));
    if (!p->p32 && !p->p64) { free(p->custom); free(p); return NULL; }
    return p;
}

void orc_plan_destroy(orc_plan *p) {
    if (!p) return;
    plan_free_f32(p->p32); plan_free_f64(p->p64); free(p->custom); free(p);
}

size_t orc_plan_n_bins(const orc_plan *p) { return p->p32 ? p->p32->n_bins : p->p64->n_bins; }
size_t orc_plan_out_len(const orc_plan *p) { return p->p32 ? p->p32->out_len : p->p64->out_len; }

void orc_plan_window(const orc_plan *p, void *out) {
    if (p->p32) memcpy(out, p->p32->window, sizeof(float) * p->d.n_fft);
    else memcpy(out, p->p64->window, sizeof(double) * p->d.n_fft);
}

void orc_plan_freq_axis(const orc_plan *p, double *out) {
    const double *src = p->p32 ? p->p32->freq_axis : p->p64->freq_axis;
    memcpy(out, src, sizeof(double) * orc_plan_n_bins(p));
}

size_t orc_plan_filterbank_nnz(const orc_plan *p) {
    if (p->d.mapping != ORC_MAP_MEL && p->d.mapping != ORC_MAP_LOGHZ) return 0;
    const size_t *rp = p->p32 ? p->p32->row_ptr : p->p64->row_ptr;
    return rp[orc_plan_n_bins(p)];
}

/* dense (n_bins, out_len) f64 copy of the mapping matrix (zeros for identity) */
void orc_plan_filterbank_dense(const orc_plan *p, double *out) {
    const size_t nb = orc_plan_n_bins(p), ol = orc_plan_out_len(p);
    memset(out, 0, sizeof(double) * nb * ol);
    if (p->d.mapping == ORC_MAP_MEL || p->d.mapping == ORC_MAP_LOGHZ) {
        const size_t *rp = p->p32 ? p->p32->row_ptr : p->p64->row_ptr;
        const size_t *col = p->p32 ? p->p32->col : p->p64->col;
        const double *val = p->p32 ? p->p32->val : p->p64->val;
        for (size_t r = 0; r < nb; ++r)
            for (size_t e = rp[r]; e < rp[r + 1]; ++e) out[r * ol + col[e]] = val[e];
    } else if (p->d.mapping == ORC_MAP_ERB) {
        memcpy(out, p->p32 ? p->p32->dense : p->p64->dense, sizeof(double) * nb * ol);
    } else {
        for (size_t r = 0; r < nb; ++r) out[r * ol + r] = 1.0;
    }
}

void orc_compute_spectrogram(orc_plan *p, const void *samples, size_t n_samples, void *out) {
    if (p->p32) compute_spectrogram_f32(p->p32, (const float *)samples, n_samples, (float *)out);
    else compute_spectrogram_f64(p->p64, (const double *)samples, n_samples, (double *)out);
}

void orc_compute_stft(orc_plan *p, const void *samples, size_t n_samples, void *out) {
    if (p->p32) compute_stft_f32(p->p32, (const float *)samples, n_samples, (float *)out);
    else compute_stft_f64(p->p64, (const double *)samples, n_samples, (double *)out);
}

/* SpectrogramPlan::compute_frame (:335-372) -- frame_idx is not range-checked by the reference */
void orc_compute_frame(orc_plan *p, const void *samples, size_t n_samples, size_t frame_idx, void *out) {
    if (p->p32) {
        frame_spectrogram_f32(p->p32, (const float *)samples, n_samples, frame_idx);
        memcpy(out, p->p32->mapped, sizeof(float) * p->p32->n_bins);
    } else {
        frame_spectrogram_f64(p->p64, (const double *)samples, n_samples, frame_idx);
        memcpy(out, p->p64->mapped, sizeof(double) * p->p64->n_bins);
    }
}

int orc_mfcc_from_log_mel(int dtype, const void *log_mel, size_t n_mels, size_t n_frames, size_t n_mfcc,
                          int include_c0, size_t lifter, int faithful, void *out) {
    g_err[0] = 0;
    int rc;
    if (dtype == ORC_F32)
        rc = mfcc_from_log_mel_f32((const float *)log_mel, n_mels, n_frames, n_mfcc, include_c0, lifter, faithful, (float *)out);
    else
        rc = mfcc_from_log_mel_f64((const double *)log_mel, n_mels, n_frames, n_mfcc, include_c0, lifter, faithful, (double *)out);
    if (rc) snprintf(g_err, sizeof(g_err), "n_mfcc must be <= n_mels");       /* src/mfcc.rs:231-233 */
    return rc;
}

/* ---------------------------------------------------------------------------------------------- */
/* Chroma (src/chroma.rs). ChromaParams::new checks :81-97; build_chroma_filterbank :279-346.        */
static int chroma_check(double sample_rate, double tuning, double f_min, double f_max) {
    if (!(sample_rate > 0.0 && isfinite(sample_rate))) { snprintf(g_err, sizeof(g_err), "sample_rate must be finite and > 0"); return 1; }
    if (!(tuning > 0.0 && isfinite(tuning))) { snprintf(g_err, sizeof(g_err), "tuning must be finite and > 0"); return 1; }
    if (!(f_min > 0.0 && isfinite(f_min))) { snprintf(g_err, sizeof(g_err), "f_min must be finite and > 0"); return 1; }
    if (f_max <= f_min) { snprintf(g_err, sizeof(g_err), "f_max must be > f_min"); return 1; }
    return 0;
}

int orc_chroma_filterbank(double sample_rate, size_t n_fft, double tuning, double f_min, double f_max, double *out) {
    g_err[0] = 0;
    if (chroma_check(sample_rate, tuning, f_min, f_max)) return 1;
    const size_t n_bins = n_fft / 2 + 1;
    const double freq_resolution = sample_rate / (double)n_fft;                      /* :293 */
    memset(out, 0, sizeof(double) * 12 * n_bins);
    for (size_t bin = 0; bin < n_bins; ++bin) {
        const double freq = (double)bin * freq_resolution;                           /* :296 */
        if (freq < f_min || freq > f_max || freq <= 0.0) continue;                   /* :308-310 */
        const double midi_note = 69.0 + 12.0 * log(freq / tuning) / M_LN2;           /* :313 */
        double pitch_class = fmod(midi_note, 12.0);                                  /* rem_euclid :316 */
        if (pitch_class < 0.0) pitch_class += 12.0;
        for (size_t c = 0; c < 12; ++c) {
            const double dist = fabs(pitch_class - (double)c);                       /* :324 */
            const double circular_dist = fmin(dist, 12.0 - dist);
            const double q = circular_dist / 1.0;                                    /* sigma = 1 semitone :328 */
            out[c * n_bins + bin] = exp(-0.5 * (q * q));                             /* :329 */
        }
    }
    for (size_t c = 0; c < 12; ++c) {                                                /* rows to unit sum :336-343 */
        double row_sum = 0.0;
        for (size_t i = 0; i < n_bins; ++i) row_sum += out[c * n_bins + i];
        if (row_sum > 0.0)
            for (size_t i = 0; i < n_bins; ++i) out[c * n_bins + i] /= row_sum;
    }
    return 0;
}

#define ORC_CHROMA_APPLY(NAME, REAL, SQRT, FMAXF)                                                                     \
    static void NAME(const REAL *spec, size_t n_bins, size_t n_frames, const double *fb, int norm, REAL *out) {       \
        for (size_t f = 0; f < n_frames; ++f) {                                                                       \
            REAL c[12];                                                                                               \
            for (size_t r = 0; r < 12; ++r) {                      /* :384-394: sum += T::from_f64(w) * x, ascending */ \
                REAL sum = (REAL)0;                                                                                   \
                for (size_t k = 0; k < n_bins; ++k) sum += (REAL)fb[r * n_bins + k] * spec[k * n_frames + f];         \
                c[r] = sum;                                                                                           \
            }                                                                                                         \
            REAL d = (REAL)0;                                      /* apply_chroma_normalization :406-453 */          \
            if (norm == ORC_CHROMANORM_L1) { for (int i = 0; i < 12; ++i) d = d + c[i]; }                             \
            else if (norm == ORC_CHROMANORM_L2) { for (int i = 0; i < 12; ++i) d = d + c[i] * c[i]; d = SQRT(d); }    \
            else if (norm == ORC_CHROMANORM_MAX) { for (int i = 0; i < 12; ++i) d = FMAXF(d, c[i]); }                 \
            if (norm != ORC_CHROMANORM_NONE && d > (REAL)0) for (int i = 0; i < 12; ++i) c[i] /= d;                   \
            for (size_t r = 0; r < 12; ++r) out[r * n_frames + f] = c[r];                                             \
        }                                                                                                             \
    }
ORC_CHROMA_APPLY(chroma_apply_f32, float, sqrtf, fmaxf)
ORC_CHROMA_APPLY(chroma_apply_f64, double, sqrt, fmax)

int orc_chroma_from_spectrogram(int dtype, const void *spec, size_t n_bins, size_t n_frames, double sample_rate,
                                size_t n_fft, double tuning, double f_min, double f_max, int norm, void *out) {
    g_err[0] = 0;
    if (n_bins != n_fft / 2 + 1) {                                                   /* :376-379 */
        snprintf(g_err, sizeof(g_err), "Dimension mismatch: expected %zu, got %zu", n_fft / 2 + 1, n_bins);
        return 2;
    }
    double *fb = (double *)malloc(sizeof(double) * 12 * n_bins);
    if (orc_chroma_filterbank(sample_rate, n_fft, tuning, f_min, f_max, fb)) { free(fb); return 1; }
    if (dtype == ORC_F32) chroma_apply_f32((const float *)spec, n_bins, n_frames, fb, norm, (float *)out);
    else chroma_apply_f64((const double *)spec, n_bins, n_frames, fb, norm, (double *)out);
    free(fb);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* Binaural cues (src/binaural.rs). pow_mag :60-83, np_mod :85-87, magphase :106-180.               */
#define ORC_BINAURAL(NAME, REAL, FMA, SQRT, ATAN2, FMOD, LOG10, PI_T)                                                  \
    static REAL NAME##_pow_mag(REAL mag, REAL mag_sq, size_t power) {                                                  \
        switch (power) {                                                                                               \
        case 1: return mag;                                                                                            \
        case 2: return mag_sq;                                                                                         \
        case 3: return mag_sq * mag;                                                                                   \
        case 4: return mag_sq * mag_sq;                                                                                \
        default: {                                                                                                     \
            REAL base = mag, acc = (REAL)1; size_t e = power;                                                          \
            while (e > 0) { if (e & 1) acc *= base; e >>= 1; if (e > 0) base *= base; }                                \
            return acc; }                                                                                              \
        }                                                                                                              \
    }                                                                                                                  \
    static void NAME##_magphase(REAL re, REAL im, size_t power, REAL *m, REAL *pr, REAL *pi) {                         \
        const REAL mag_sq = FMA(re, re, im * im);                        /* c.re.mul_add(c.re, c.im * c.im) :122 */     \
        if (mag_sq == (REAL)0) { *m = (REAL)0; *pr = (REAL)1; *pi = (REAL)0; return; }                                 \
        const REAL mag = SQRT(mag_sq);                                                                                 \
        *m = NAME##_pow_mag(mag, mag_sq, power);                                                                       \
        const REAL inv = (REAL)1 / mag;                                  /* recip :132 */                               \
        *pr = re * inv; *pi = im * inv;                                                                                \
    }                                                                                                                  \
    static REAL NAME##_np_mod(REAL x, REAL m) { return FMOD(FMOD(x, m) + m, m); }                                      \
    static void NAME(int cue, const REAL *left, const REAL *right, size_t n_frames, size_t start_bin, size_t stop_bin, \
                     double bin_width, size_t power, int wrapped, REAL *out) {                                         \
        const REAL pi = PI_T, two_pi = (REAL)2.0 * pi, bw = (REAL)bin_width;                                           \
        for (size_t b = start_bin; b < stop_bin; ++b)                                                                  \
            for (size_t f = 0; f < n_frames; ++f) {                                                                    \
                const REAL *l = left + 2 * (b * n_frames + f), *r = right + 2 * (b * n_frames + f);                    \
                REAL ml, plr, pli, mr, prr, pri, o;                                                                    \
                NAME##_magphase(l[0], l[1], cue == ORC_CUE_ITD ? power : 1, &ml, &plr, &pli);                          \
                NAME##_magphase(r[0], r[1], cue == ORC_CUE_ITD ? power : 1, &mr, &prr, &pri);                          \
                if (cue == ORC_CUE_ITD) {                                    /* :528-545 */                             \
                    o = (REAL)0;                                                                                       \
                    if (ml + mr > (REAL)0) {                                                                           \
                        const REAL diff = ATAN2(pli, plr) - ATAN2(pri, prr);                                           \
                        const REAL w = NAME##_np_mod(diff + pi, two_pi) - pi;                                          \
                        o = w / (two_pi * bw * (REAL)b);                                                               \
                    }                                                                                                  \
                } else if (cue == ORC_CUE_IPD) {                             /* :875-889 */                             \
                    const REAL diff = ATAN2(pli, plr) - ATAN2(pri, prr);                                               \
                    o = wrapped ? NAME##_np_mod(diff + pi, two_pi) - pi : diff;                                        \
                } else {                                                                                               \
                    o = (REAL)NAN;                                           /* from_elem(.., T::nan()) :1212, :1555 */ \
                    if (ml + mr > (REAL)0 && ml > (REAL)0 && mr > (REAL)0) {                                           \
                        const REAL ratio = mr / ml;                                                                    \
                        if (cue == ORC_CUE_ILD) o = (REAL)-20.0 * LOG10(ratio);              /* :1229-1231 */           \
                        else o = ratio < (REAL)1 ? (REAL)1 - ratio : -((REAL)1 - (REAL)1 / ratio);   /* :1572-1580 */   \
                    }                                                                                                  \
                }                                                                                                      \
                out[(b - start_bin) * n_frames + f] = o;                                                               \
            }                                                                                                          \
    }
ORC_BINAURAL(binaural_f32, float, fmaf, sqrtf, atan2f, fmodf, log10f, (float)M_PI)
ORC_BINAURAL(binaural_f64, double, fma, sqrt, atan2, fmod, log10, M_PI)

int orc_binaural_from_stft(int dtype, int cue, const void *left, const void *right, size_t n_bins, size_t n_frames,
                           size_t start_bin, size_t stop_bin, double bin_width, size_t magphase_power, int wrapped, void *out) {
    g_err[0] = 0;
    if (start_bin >= stop_bin || stop_bin > n_bins) { snprintf(g_err, sizeof(g_err), "Frequency range should have at least one bin"); return 1; }
    if (dtype == ORC_F32) binaural_f32(cue, (const float *)left, (const float *)right, n_frames, start_bin, stop_bin, bin_width, magphase_power, wrapped, (float *)out);
    else binaural_f64(cue, (const double *)left, (const double *)right, n_frames, start_bin, stop_bin, bin_width, magphase_power, wrapped, (double *)out);
    return 0;
}

/* inverse path (src/spectrogram.rs:4789-4911) */
int orc_irfft(int dtype, const void *spectrum, size_t n_fft, void *out) {
    g_err[0] = 0;
    if (dtype == ORC_F32) {
        fft_plan_f32 *fp = fft_plan_new_f32(n_fft);
        irfft_f32(fp, (const cpx_f32 *)spectrum, (float *)out);
        fft_plan_free_f32(fp);
    } else {
        fft_plan_f64 *fp = fft_plan_new_f64(n_fft);
        irfft_f64(fp, (const cpx_f64 *)spectrum, (double *)out);
        fft_plan_free_f64(fp);
    }
    return 0;
}

size_t orc_istft(orc_plan *p, const void *stft, size_t n_frames, void *out) {
    if (p->p32) return istft_f32(p->p32->fft, (const cpx_f32 *)stft, n_frames, p->d.hop, p->d.centre, p->p32->window, (float *)out);
    return istft_f64(p->p64->fft, (const cpx_f64 *)stft, n_frames, p->d.hop, p->d.centre, p->p64->window, (double *)out);
}

/* single-frame R2C of <= n_fft samples, zero padded: free fn fft() (:4490-4520) */
int orc_rfft(int dtype, const void *x, size_t n_in, size_t n_fft, void *out) {
    g_err[0] = 0;
    if (n_in > n_fft) { snprintf(g_err, sizeof(g_err), "Input length (%zu) exceeds FFT size (%zu)", n_in, n_fft); return 1; }
    if (dtype == ORC_F32) {
        fft_plan_f32 *fp = fft_plan_new_f32(n_fft);
        float *buf = (float *)calloc(n_fft, sizeof(float));
        memcpy(buf, x, sizeof(float) * n_in);
        rfft_f32(fp, buf, (cpx_f32 *)out);
        free(buf); fft_plan_free_f32(fp);
    } else {
        fft_plan_f64 *fp = fft_plan_new_f64(n_fft);
        double *buf = (double *)calloc(n_fft, sizeof(double));
        memcpy(buf, x, sizeof(double) * n_in);
        rfft_f64(fp, buf, (cpx_f64 *)out);
        free(buf); fft_plan_free_f64(fp);
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* Batch driver: the reference has no batch API (src/lib.rs:228-235) and no threading on this path  */
/* (rayon only in src/binaural.rs); its documented scaling recipe is one plan per worker thread     */
/* (docs/source/guide/performance.rst:208-227). This is that recipe: clips sharded over n_threads,  */
/* each worker owning a private plan. Used for bench.py's cpu_baseline / --impl reference.          */
typedef struct {
    const orc_desc *d; const void *samples; size_t n_samples, clip_stride, c0, c1;
    void *out; size_t out_stride; int mfcc; size_t n_mfcc; int include_c0; size_t lifter; int faithful;
    int rc;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    orc_plan *p = orc_plan_create(j->d);
    if (!p) { j->rc = 1; return NULL; }
    const size_t es = j->d->dtype == ORC_F32 ? 4 : 8;
    const size_t nb = orc_plan_n_bins(p);
    const size_t nf = orc_frame_count(j->n_samples, j->d->n_fft, j->d->hop, j->d->centre);
    void *tmp = j->mfcc ? malloc(es * nb * nf) : NULL;
    for (size_t c = j->c0; c < j->c1; ++c) {
        const char *in = (const char *)j->samples + es * c * j->clip_stride;
        char *o = (char *)j->out + es * c * j->out_stride;
        if (j->mfcc) {
            orc_compute_spectrogram(p, in, j->n_samples, tmp);
            orc_mfcc_from_log_mel(j->d->dtype, tmp, nb, nf, j->n_mfcc, j->include_c0, j->lifter, j->faithful, o);
        } else {
            orc_compute_spectrogram(p, in, j->n_samples, o);
        }
    }
    free(tmp);
    orc_plan_destroy(p);
    return NULL;
}

int orc_compute_batch(const orc_desc *d, const void *samples, size_t n_clips, size_t n_samples, size_t clip_stride,
                      void *out, size_t out_stride, int n_threads,
                      int mfcc, size_t n_mfcc, int include_c0, size_t lifter, int faithful) {
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_clips) n_threads = (int)n_clips;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    batch_job *jobs = (batch_job *)calloc(n_threads, sizeof(batch_job));
    const size_t per = (n_clips + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        batch_job *j = &jobs[t];
        j->d = d; j->samples = samples; j->n_samples = n_samples; j->clip_stride = clip_stride;
        j->c0 = (size_t)t * per; j->c1 = j->c0 + per > n_clips ? n_clips : j->c0 + per;
        if (j->c0 > n_clips) j->c0 = n_clips;
        j->out = out; j->out_stride = out_stride; j->mfcc = mfcc; j->n_mfcc = n_mfcc;
        j->include_c0 = include_c0; j->lifter = lifter; j->faithful = faithful;
        pthread_create(&th[t], NULL, batch_worker, j);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}
