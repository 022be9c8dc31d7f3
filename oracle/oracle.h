/*
 * oracle.h -- C interface of the CPU oracle (test infrastructure; see oracle.c header).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_F32 = 0, ORC_F64 = 1 };
enum { ORC_MAP_IDENTITY = 0, ORC_MAP_MEL = 1, ORC_MAP_ERB = 2, ORC_MAP_LOGHZ = 3 };
enum { ORC_AMP_POWER = 0, ORC_AMP_MAGNITUDE = 1, ORC_AMP_DECIBELS = 2 };
enum { ORC_WIN_RECT = 0, ORC_WIN_HANN = 1, ORC_WIN_HAMMING = 2, ORC_WIN_BLACKMAN = 3, ORC_WIN_KAISER = 4,
       ORC_WIN_GAUSSIAN = 5, ORC_WIN_CUSTOM = 6 };
enum { ORC_MELNORM_NONE = 0, ORC_MELNORM_SLANEY = 1, ORC_MELNORM_L1 = 2, ORC_MELNORM_L2 = 3 };
enum { ORC_ERB_LINEAR = 0, ORC_ERB_APPLE_TR35 = 1 };

typedef struct {
    int dtype;
    size_t n_fft, hop;
    int centre;
    int window_kind;
    double window_param;            /* Kaiser beta / Gaussian std */
    const double *custom_window;
    size_t custom_window_len;
    double sample_rate;
    int mapping;
    size_t n_bands;                 /* n_mels / n_filters / n_bins */
    double f_min, f_max;
    int mel_norm;
    int erb_spacing;
    int amp;
    int has_floor_db;
    double floor_db;
} orc_desc;

typedef struct {
    size_t out_len, n_bins, nnz;
    double *window;
    size_t *row_ptr, *col;
    double *val;
    double *dense;
    double *freq_axis;
} orc_tables;

typedef struct orc_plan orc_plan;

size_t orc_frame_count(size_t n_samples, size_t n_fft, size_t hop, int centre);
int orc_build_tables(const orc_desc *d, orc_tables *t, char *err, size_t errlen);
void orc_free_tables(orc_tables *t);

const char *orc_last_error(void);
orc_plan *orc_plan_create(const orc_desc *d);
void orc_plan_destroy(orc_plan *p);
size_t orc_plan_n_bins(const orc_plan *p);
size_t orc_plan_out_len(const orc_plan *p);
void orc_plan_window(const orc_plan *p, void *out);
void orc_plan_freq_axis(const orc_plan *p, double *out);
size_t orc_plan_filterbank_nnz(const orc_plan *p);
void orc_plan_filterbank_dense(const orc_plan *p, double *out);
void orc_compute_spectrogram(orc_plan *p, const void *samples, size_t n_samples, void *out);
void orc_compute_stft(orc_plan *p, const void *samples, size_t n_samples, void *out);
void orc_compute_frame(orc_plan *p, const void *samples, size_t n_samples, size_t frame_idx, void *out);
int orc_mfcc_from_log_mel(int dtype, const void *log_mel, size_t n_mels, size_t n_frames, size_t n_mfcc,
                          int include_c0, size_t lifter, int faithful, void *out);
enum { ORC_CHROMANORM_NONE = 0, ORC_CHROMANORM_L1 = 1, ORC_CHROMANORM_L2 = 2, ORC_CHROMANORM_MAX = 3 };
/* src/chroma.rs: build_chroma_filterbank (:279-346) into out[12][n_fft/2+1]; chromagram_from_spectrogram (:365-404)
 * with apply_chroma_normalization (:406-453): spec (n_bins, n_frames) of dtype -> out (12, n_frames). 0 = ok. */
int orc_chroma_filterbank(double sample_rate, size_t n_fft, double tuning, double f_min, double f_max, double *out);
int orc_chroma_from_spectrogram(int dtype, const void *spec, size_t n_bins, size_t n_frames, double sample_rate,
                                size_t n_fft, double tuning, double f_min, double f_max, int norm, void *out);
enum { ORC_CUE_ITD = 0, ORC_CUE_IPD = 1, ORC_CUE_ILD = 2, ORC_CUE_ILR = 3 };
/* src/binaural.rs: the element-wise half of compute_{itd,ipd,ild,ilr}_spectrogram (:472-580, :830-917, :1187-1262,
 * :1530-1620) incl. magphase (:106-180): left/right complex (n_bins, n_frames) -> out (stop_bin-start_bin, n_frames). */
int orc_binaural_from_stft(int dtype, int cue, const void *left, const void *right, size_t n_bins, size_t n_frames,
                           size_t start_bin, size_t stop_bin, double bin_width, size_t magphase_power, int wrapped, void *out);
/* irfft (src/spectrogram.rs:4789-4811): spectrum[n_fft/2+1] -> out[n_fft]; istft (:4813-4911) with the plan's n_fft, hop,
 * window and centre: stft (n_fft/2+1, n_frames) complex -> out; returns the output length (out == NULL: query only). */
int orc_irfft(int dtype, const void *spectrum, size_t n_fft, void *out);
size_t orc_istft(orc_plan *p, const void *stft, size_t n_frames, void *out);
int orc_rfft(int dtype, const void *x, size_t n_in, size_t n_fft, void *out);
int orc_compute_batch(const orc_desc *d, const void *samples, size_t n_clips, size_t n_samples, size_t clip_stride,
                      void *out, size_t out_stride, int n_threads,
                      int mfcc, size_t n_mfcc, int include_c0, size_t lifter, int faithful);

#ifdef __cplusplus
}
#endif
#endif
