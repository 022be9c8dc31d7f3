/*
 * sgx_b200.h -- C ABI of libsgx_b200.so, the B200-native (sm_100a) engine for the hot path of the
 * `spectrograms` crate (jmg049/Spectrograms v2.1.0): stft() / StftPlan / SpectrogramPlanner plans
 * (linear | mel | ERB | LogHz  x  power | magnitude | dB) / mfcc_from_log_mel, in f32 and f64, and the adjacent
 * components built on the same kernels: chromagram, the binaural cue spectrograms, irfft / istft.
 * (sgx_b200.hpp is the C++ host layer over these entry points.)
 *
 * The reference exposes no C ABI; its seam for this path is the plan API. Every entry point
 * below names the reference item it replaces (paths relative to the reference checkout; a bare :N is
 * src/spectrogram.rs:N). A Rust `gpu` feature would bind these with `extern "C"`.
 *
 * Conventions (mirroring the reference):
 *   - Errors: every call returns sgx_status; the four non-OK codes are the four SpectrogramError variants
 *     (src/error.rs:13-28). sgx_last_error_message() returns the thread-local Display string of the last error.
 *   - Ownership: the caller owns every buffer; nothing is retained after a call returns (for device pointers: after
 *     the work queued on `stream` completes). A plan owns its device tables and scratch.
 *   - Threading: a plan is `&mut self` -- one call at a time per plan; distinct plans are independent. Calls that stage through
 *     plan-owned device scratch (sgx_plan_istft, the unfused route of sgx_plan_compute_binaural) order themselves after the
 *     previous use of that scratch with an event, so consecutive asynchronous calls on different streams do not race.
 *   - Layout: outputs are row-major (rows, n_frames) with frames contiguous (:248, :1433), clips outermost.
 *     Complex values are interleaved (re, im) like num_complex::Complex<T>.
 *   - Pointers may be host or device pointers (detected with cudaPointerGetAttributes). Device pointers run
 *     asynchronously on `stream`; host pointers are copied in chunks (64 MB) to device staging buffers, computed and copied back
 *     over three private streams, and the call returns after the results are in host memory. The copies come straight from
 *     the caller's buffers: they overlap with compute only if the caller's memory is pinned (cudaHostAlloc / cudaHostRegister);
 *     pageable memory works but serialises.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with SGX_BACKEND_ERROR.
 */
#ifndef SGX_B200_H
#define SGX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SpectrogramError (src/error.rs:13-28) */
typedef enum {
    SGX_OK = 0,
    SGX_INVALID_INPUT = 1,       /* InvalidInput(String) */
    SGX_DIMENSION_MISMATCH = 2,  /* DimensionMismatch{expected, got} */
    SGX_BACKEND_ERROR = 3,       /* FftBackendError{backend: "cuda", msg} */
    SGX_INTERNAL_ERROR = 4       /* InternalError(String) */
} sgx_status;

/* Sample (src/sample.rs:23-86): the two sealed implementors */
typedef enum { SGX_F32 = 0, SGX_F64 = 1 } sgx_dtype;

/* WindowType (src/window.rs:19-50) */
typedef enum {
    SGX_WIN_RECTANGULAR = 0, SGX_WIN_HANNING = 1, SGX_WIN_HAMMING = 2, SGX_WIN_BLACKMAN = 3,
    SGX_WIN_KAISER = 4,   /* window_param = beta */
    SGX_WIN_GAUSSIAN = 5, /* window_param = std (samples) */
    SGX_WIN_CUSTOM = 6    /* custom_window[custom_window_len] */
} sgx_window;

/* frequency scales: LinearHz / Mel / Erb / LogHz marker types (MappingKind :1639-1656; Cqt is out of scope).
 * SGX_MAP_CHROMA is the fused chromagram() (src/chroma.rs:487-503): linear magnitude spectrogram -> 12 x bins chroma
 * filterbank (build_chroma_filterbank, src/chroma.rs:279-346) -> per-frame normalisation (:406-453); it requires
 * amp = SGX_AMP_MAGNITUDE, no dB floor, output = SGX_OUT_SPECTROGRAM and yields 12 rows. */
typedef enum { SGX_MAP_LINEAR = 0, SGX_MAP_MEL = 1, SGX_MAP_ERB = 2, SGX_MAP_LOGHZ = 3, SGX_MAP_CHROMA = 4 } sgx_mapping;

/* ChromaNorm (src/chroma.rs:31-45) */
typedef enum { SGX_CHROMANORM_NONE = 0, SGX_CHROMANORM_L1 = 1, SGX_CHROMANORM_L2 = 2, SGX_CHROMANORM_MAX = 3 } sgx_chroma_norm;

/* AmpScaleSpec implementors Power / Magnitude / Decibels (:1986-2037) */
typedef enum { SGX_AMP_POWER = 0, SGX_AMP_MAGNITUDE = 1, SGX_AMP_DECIBELS = 2 } sgx_amp;

/* MelNorm (:3708-3734) and ErbSpacing (src/erb.rs:17-25) */
typedef enum { SGX_MELNORM_NONE = 0, SGX_MELNORM_SLANEY = 1, SGX_MELNORM_L1 = 2, SGX_MELNORM_L2 = 3 } sgx_mel_norm;
typedef enum { SGX_ERB_LINEAR = 0, SGX_ERB_APPLE_TR35 = 1 } sgx_erb_spacing;

/* what a plan produces */
typedef enum {
    SGX_OUT_SPECTROGRAM = 0,  /* SpectrogramPlan::compute  -> (n_bins, n_frames) of T             (:240-294)   */
    SGX_OUT_COMPLEX_STFT = 1, /* StftPlan::compute / stft() -> (n_fft/2+1, n_frames) of Complex<T> (:1424-1458) */
    SGX_OUT_MFCC = 2          /* mfcc(): mel dB plan fused with mfcc_from_log_mel -> (n_mfcc[-1], n_frames)
                                 (src/mfcc.rs:359-379, :224-273) */
} sgx_output;

/*
 * One flat description of a plan = StftParams (:3452-3506) + SpectrogramParams (:4108-4140) + the scale-specific
 * params MelParams (:3744-3920) / ErbParams (src/erb.rs:30-43) / LogHzParams (:3955-3975) + Option<LogParams>
 * (:4052-4100) + MfccParams (src/mfcc.rs:21-39). Zero-initialise, then fill what applies.
 */
typedef struct {
    sgx_dtype dtype;
    size_t n_fft;               /* NonZeroUsize */
    size_t hop_size;            /* NonZeroUsize, <= n_fft (:3485) */
    int centre;                 /* zero padding of n_fft/2 both sides (:1236-1237) */
    sgx_window window;
    double window_param;
    const double *custom_window;
    size_t custom_window_len;   /* must equal n_fft (:3490-3497) */
    double sample_rate_hz;      /* finite, > 0 (:4130) */

    sgx_mapping mapping;
    size_t n_bands;             /* n_mels / n_filters / n_bins; ignored for SGX_MAP_LINEAR */
    double f_min, f_max;
    sgx_mel_norm mel_norm;
    sgx_erb_spacing erb_spacing;

    sgx_amp amp;
    int has_floor_db;           /* Option<&LogParams>: 0 = None. Decibels with None yields raw power (:2052-2058, :2075) */
    double floor_db;

    sgx_output output;
    size_t n_mfcc;              /* SGX_OUT_MFCC only */
    int include_c0;
    size_t lifter;

    int device;                 /* CUDA ordinal; -1 = current device */

    /* SGX_MAP_CHROMA only: ChromaParams (src/chroma.rs:18-30); f_min / f_max above are its frequency range */
    double chroma_tuning;       /* A4 reference in Hz, finite and > 0 (:82-86) */
    sgx_chroma_norm chroma_norm;
} sgx_plan_desc;

typedef struct sgx_plan sgx_plan;

/* Display string of the last error raised on this thread ("Invalid input: ...", "Dimension mismatch: expected N,
 * got M", "cuda -- FFT backend error: ...", "Internal error: ..."), src/error.rs:17-27. Never NULL. */
const char *sgx_last_error_message(void);

/* expected/got of the last SGX_DIMENSION_MISMATCH on this thread (src/error.rs:21) */
void sgx_last_dimension_mismatch(size_t *expected, size_t *got);

/* Library / build information: "sgx_b200 <version> sm_100a". */
const char *sgx_version(void);

/*
 * SpectrogramPlanner::{linear,mel,erb,log_hz}_plan::<A,T> (:893-1102) and StftPlan::<T>::new (:1204-1228).
 * Validates like StftParams::new, SpectrogramParams::new, MelParams/ErbParams/LogHzParams::new, the planner's
 * f_max <= Nyquist checks (:954-959, :1016-1022, :1078-1084) and LogParams::new; builds window, twiddles,
 * filterbank, DCT basis in f64 on the host exactly as the reference does and uploads them once.
 */
sgx_status sgx_plan_create(const sgx_plan_desc *desc, sgx_plan **out_plan);
sgx_status sgx_plan_destroy(sgx_plan *plan);

/* SpectrogramPlan::output_shape (:512-519) / StftPlan::output_shape (:1596-1602) -- frame_count (:1230-1250). */
sgx_status sgx_plan_output_shape(const sgx_plan *plan, size_t n_samples, size_t *n_rows, size_t *n_frames);

/* Axes: FrequencyMapping::frequencies_hz (:1909-1945) into freqs[n_bins]; build_time_axis_seconds (:2128-2139)
 * into times[n_frames]. Either pointer may be NULL. For SGX_OUT_MFCC the frequency axis is the mel axis. */
sgx_status sgx_plan_axes(const sgx_plan *plan, size_t n_frames, double *freqs, double *times);

/* make_window::<T> (:2159-2235): writes n_fft values of the plan's dtype to host memory. */
sgx_status sgx_plan_window(const sgx_plan *plan, void *out_host);

/* Inspection of the frequency mapping: dense row-major (n_bins, n_fft/2+1) f64 copy of the SparseMatrix (:43-118)
 * or ERB response matrix (src/erb.rs:261); *nnz = stored non-zeros (sparse mappings) or n_bins*out_len (dense). */
sgx_status sgx_plan_filterbank(const sgx_plan *plan, double *dense_out_host, size_t *nnz);

/* Name of the CUDA kernel family the plan dispatches to (e.g. "r2c_fused_generic", "r2c_fused_n400_tm"); two-kernel paths are
 * joined with '+' ("r2c_fused_n400_tm+dct2_lifter_tc" for mfcc(), "r2c_fused_pow2+dense_rows_tc" for a dense ERB plan). */
const char *sgx_plan_kernel_name(const sgx_plan *plan);

/* Number of kernel launches issued by the last compute call on this plan. */
size_t sgx_plan_last_launch_count(const sgx_plan *plan);

/* Force the generic kernel family (1) or allow specialised kernels (0, default). Test hook. */
sgx_status sgx_plan_force_generic(sgx_plan *plan, int force);

/* TMEM / tcgen05 kernel variant of a family (r2c_fused_n400_tc: the FFT exchange lives in tensor memory and the filterbank
 * projection -- FrequencyMapping::apply, src/spectrogram.rs:1822-1881 -- runs as 3xTF32 MMAs): -1 = automatic (default:
 * used where it is measured faster, i.e. the dense ERB projection), 0 = never, 1 = whenever the plan supports it.
 * The same switch governs the split path of dense f32 (ERB) plans that the fused kernel does not cover (other n_fft, more than
 * 64 bands): linear power spectrogram by the plan's FFT family, then the filterbank (src/erb.rs:374-402) as row blocks of the
 * tcgen05 GEMM kernel ("...+dense_rows_tc"); 0 keeps the CUDA-core dense rows. Test / measurement hook. */
sgx_status sgx_plan_set_tensor_cores(sgx_plan *plan, int enable);

/* Tensor memory as the exchange medium between the two FFT passes of the n_fft = 400 / hop = 160 f32 family
 * (r2c_fused_n400_tm; the per-frame FFT of SpectrogramPlan::compute, src/spectrogram.rs:240-294, src/fft_backend.rs:423-431):
 * -1 = automatic (default: used wherever the plan supports it -- sparse mel / loghz spectrogram outputs), 0 = never (the
 * shared-memory exchange of r2c_fused_n400), 1 = whenever the plan supports it. Results are bit-identical either way.
 * Test / measurement hook. */
sgx_status sgx_plan_set_tmem_exchange(sgx_plan *plan, int enable);

/*
 * The batched entry point: for c in 0..n_clips { plan.compute_into(&samples[c], &mut out[c]) }  (:414-477,
 * :1548-1580; the reference has no batch API -- batching is a user loop, src/lib.rs:228-235).
 *   samples : [n_clips][clip_stride] of T, n_samples valid per clip (n_samples >= 1, NonEmptySlice)
 *   out     : [n_clips][out_rows][out_cols] of T (Complex<T> for SGX_OUT_COMPLEX_STFT), clip pitch out_clip_stride
 *             elements (0 = out_rows*out_cols)
 * out_rows/out_cols are validated like compute_into: rows first, then columns (:423-434) -> SGX_DIMENSION_MISMATCH.
 * n_clips = 1 with host pointers reproduces the reference call exactly.
 */
sgx_status sgx_plan_compute_batch(sgx_plan *plan, const void *samples, size_t n_clips, size_t n_samples,
                                  size_t clip_stride, void *out, size_t out_rows, size_t out_cols,
                                  size_t out_clip_stride, void *cuda_stream);

/* SpectrogramPlan::compute_frame (:335-372) / StftPlan::compute_frame_simple (:1500-1507): one frame -> out[rows].
 * Like the reference, frame_idx is not range checked: past the end it reads zero padding. */
sgx_status sgx_plan_compute_frame(sgx_plan *plan, const void *samples, size_t n_samples, size_t frame_idx,
                                  void *out, void *cuda_stream);

/*
 * mfcc_from_log_mel::<T> (src/mfcc.rs:224-273): unnormalised DCT-II over n_mels (:278-292), lifter (:297-316),
 * optional c0 drop (:262-267). log_mel: [n_clips][n_mels][n_frames]; out: [n_clips][rows][n_frames] with
 * rows = n_mfcc - (include_c0 || n_mfcc == 1 ? 0 : 1). n_mfcc > n_mels -> SGX_INVALID_INPUT (:231-233).
 */
sgx_status sgx_mfcc_from_log_mel(sgx_dtype dtype, const void *log_mel, size_t n_clips, size_t n_mels,
                                 size_t n_frames, size_t n_mfcc, int include_c0, size_t lifter, void *out,
                                 int device, void *cuda_stream);

/* Interaural cues (src/binaural.rs): ITD seconds, IPD radians, ILD dB, ILR ratio */
typedef enum { SGX_CUE_ITD = 0, SGX_CUE_IPD = 1, SGX_CUE_ILD = 2, SGX_CUE_ILR = 3 } sgx_binaural_cue;

/*
 * The element-wise half of compute_{itd,ipd,ild,ilr}_spectrogram (src/binaural.rs:472-580, :830-917, :1187-1262,
 * :1530-1620) on STFTs the caller already holds (StftPlan::compute of the left and the right channel, i.e.
 * sgx_plan_compute_batch of a SGX_OUT_COMPLEX_STFT plan): left / right are (n_pairs, n_bins, n_frames) of Complex<T>,
 * out is (n_pairs, stop_bin - start_bin, n_frames) of T for the band [start_bin, stop_bin) =
 * [round(start_freq / bin_width), round(end_freq / bin_width)) (:478-481). magphase_power is ITD's weighting exponent
 * (ITDSpectrogramParams, :386-391; the other cues use 1), wrapped is IPD's flag (:755-760). ILD / ILR are NaN where a
 * channel has no energy, ITD is 0 there. Host or device pointers (all three the same kind).
 */
sgx_status sgx_binaural_from_stft(sgx_dtype dtype, sgx_binaural_cue cue, const void *left, const void *right, size_t n_pairs,
                                  size_t n_bins, size_t n_frames, size_t start_bin, size_t stop_bin, double bin_width_hz,
                                  size_t magphase_power, int wrapped, void *out, int device, void *cuda_stream);

/*
 * compute_{itd,ipd,ild,ilr}_spectrogram(audio: [left, right], params, plan: &mut StftPlan) (src/binaural.rs:472, :830,
 * :1187, :1530), batched over stereo pairs: plan must be a SGX_OUT_COMPLEX_STFT plan (the reference's "force reuse"
 * StftPlan argument); left / right are (n_pairs, clip_stride) sample matrices, out is (n_pairs, out_bins, out_frames) with
 * out_bins = round(end_freq / bw) - round(start_freq / bw), bw = sample_rate / n_fft (:476-481) and out_frames the plan's
 * frame count (SGX_DIMENSION_MISMATCH otherwise). Both STFTs stay in device memory; only the cue matrix is written.
 * Host or device pointers (all three the same kind); device calls are asynchronous on cuda_stream.
 */
sgx_status sgx_plan_compute_binaural(sgx_plan *plan, sgx_binaural_cue cue, const void *left, const void *right,
                                     size_t n_pairs, size_t n_samples, size_t clip_stride, double start_freq, double end_freq,
                                     size_t magphase_power, int wrapped, void *out, size_t out_bins, size_t out_frames,
                                     void *cuda_stream);

/*
 * chromagram_from_spectrogram (src/chroma.rs:365-404): spec is (n_clips, n_bins, n_frames) of T with
 * n_bins == n_fft/2 + 1 (DimensionMismatch otherwise, :376-379), any amplitude scale; out is (n_clips, 12, n_frames).
 * The filterbank is built in f64 and cast per element (T::from_f64), rows accumulate in ascending bin order, then the
 * per-frame normalisation. Host or device pointers (both the same kind).
 */
sgx_status sgx_chroma_from_spectrogram(sgx_dtype dtype, const void *spec, size_t n_clips, size_t n_bins, size_t n_frames,
                                       double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max,
                                       sgx_chroma_norm norm, void *out, int device, void *cuda_stream);

/* build_chroma_filterbank (src/chroma.rs:279-346) into dense_out[12][n_fft/2 + 1] (f64, host). */
sgx_status sgx_chroma_filterbank(double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max,
                                 double *dense_out);

/* free fn fft() / rfft helpers (:4490-4520): one unnormalised R2C of n_in <= n_fft samples (zero padded), no window.
 * out: n_fft/2+1 Complex<T>. n_in > n_fft -> SGX_INVALID_INPUT "Input length (..) exceeds FFT size (..)". */
sgx_status sgx_rfft(sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft, void *out, int device,
                    void *cuda_stream);

/*
 * istft<T>(stft_matrix, n_fft, hop_size, window, center) (src/spectrogram.rs:4813-4911), batched: the plan supplies n_fft,
 * hop_size, the window and centre (any output kind; a SGX_OUT_COMPLEX_STFT plan that produced the matrix is the natural
 * one). stft is (n_clips, n_fft/2 + 1, n_frames) of Complex<T>; out is (n_clips, out_len) of T with
 * out_len = (n_frames - 1) * hop + n_fft, minus 2 * (n_fft / 2) when centred and that leaves anything (:4836-4841,
 * :4893-4902) -- query it with out = NULL: *out_len_io is set and nothing runs. Per frame: C2R inverse FFT scaled by 1/n_fft
 * (src/fft_backend.rs:536-566; like realfft, the imaginary parts of the DC / Nyquist bins are ignored), synthesis window,
 * overlap-add, division by the accumulated squared window where it exceeds 1e-10. Host or device pointers.
 */
sgx_status sgx_plan_istft(sgx_plan *plan, const void *stft, size_t n_clips, size_t n_frames, void *out, size_t *out_len_io,
                          void *cuda_stream);

/* irfft<T>(spectrum, n_fft) (src/spectrogram.rs:4789-4811): spectrum_len must be n_fft/2 + 1 (SGX_DIMENSION_MISMATCH
 * otherwise); out receives n_fft samples. Host or device pointers. */
sgx_status sgx_irfft(sgx_dtype dtype, const void *spectrum, size_t spectrum_len, size_t n_fft, void *out, int device,
                     void *cuda_stream);

/* ---- FftPlanner (src/spectrogram.rs:4977-5235): a plan cache for repeated single-frame transforms. The reference object wraps
 * its FFT backend's planner (plans cached by size); here the cache holds complete plans (device tables included) keyed by
 * dtype, n_fft, window and output, so the second call of a size pays no table construction or upload. One planner = one
 * device; like the reference's `&mut self` methods a planner serves one caller at a time. Host or device pointers. */
typedef struct sgx_fft_planner sgx_fft_planner;
sgx_status sgx_fft_planner_create(int device /* -1 = current */, sgx_fft_planner **out);      /* FftPlanner::new :4988 */
sgx_status sgx_fft_planner_destroy(sgx_fft_planner *planner);
size_t sgx_fft_planner_cached_plans(const sgx_fft_planner *planner);
/* FftPlanner::fft :5028-5061: unnormalised R2C of n_in <= n_fft samples (zero padded) -> n_fft/2+1 Complex<T> */
sgx_status sgx_fft_planner_rfft(sgx_fft_planner *planner, sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft, void *out,
                                void *cuda_stream);
/* FftPlanner::irfft :5113-5135: spectrum_len must be n_fft/2+1 (SGX_DIMENSION_MISMATCH otherwise) -> n_fft samples */
sgx_status sgx_fft_planner_irfft(sgx_fft_planner *planner, sgx_dtype dtype, const void *spectrum, size_t spectrum_len, size_t n_fft,
                                 void *out, void *cuda_stream);
/* FftPlanner::power_spectrum :5163-5197 / magnitude_spectrum :5224-5235: zero pad, optional window (SGX_WIN_RECTANGULAR = none),
 * |X|^2 (magnitude = 0) or |X| (magnitude = 1) -> n_fft/2+1 values of T */
sgx_status sgx_fft_planner_power_spectrum(sgx_fft_planner *planner, sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft,
                                          sgx_window window, double window_param, int magnitude, void *out, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SGX_B200_H */
