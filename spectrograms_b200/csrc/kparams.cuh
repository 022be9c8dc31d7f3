// kparams.cuh -- kernel parameter block and small device helpers shared by all kernel families.
#pragma once

#include <cuda_runtime.h>

#include "../../include/sgx_b200.h"
#include "fastdiv.hpp"

namespace sgx {

constexpr int kMaxStages = 24;

// Everything a fused STFT -> power -> mapping -> scaling (-> DCT) kernel needs. Passed by value (__grid_constant__).
struct KParams {
    // ---- input: [n_clips][clip_stride] samples, n_samples valid per clip
    const void *samples;
    long long n_samples;
    long long clip_stride;
    int n_clips;
    // ---- which frames: [frame_begin, frame_begin + frames_todo), tiled FT frames per CTA
    long long frame_begin;
    long long frames_todo;
    int tiles_per_clip;
    int FT;
    // ---- output: element (clip, row, frame) at out[clip*out_clip_stride + row*out_row_stride + (frame - frame_begin_out)]
    void *out;
    long long out_row_stride;
    long long out_clip_stride;
    long long out_frame_origin;
    // ---- STFT (StftPlan fields, src/spectrogram.rs:1173-1187)
    int n_fft, hop, pad, out_len;
    int L;        // complex FFT length: n_fft/2 (even n_fft, packed real input) or n_fft (odd)
    int even;
    int n_stages;
    int radix[kMaxStages];
    const void *window;   // T[n_fft]            make_window cast to T (:2232)
    const void *tw;       // complex T[L]        W_L^k
    const void *post;     // complex T[L+1]      W_{n_fft}^k (even n_fft only)
    // ---- frequency mapping (MappingKind :1639-1656)
    int mapping;          // sgx_mapping
    int n_bins;
    const int *row_ptr;   // CSR (mel / loghz)
    const int *col;
    const void *val;      // T[nnz]              T::from_f64(value) (:113)
    const void *dense;    // T[n_bins][out_len]  (erb)
    int rows_contig;      // every CSR row's columns are consecutive (mel triangles, loghz pairs): col[e] = col[e0] + (e - e0)
    const int4 *row_desc; // rows_contig only: per row {first entry e0, count, first column, 0} -- one 16-byte load instead of three dependent ones
    const int4 *lane_rows;    // r2c_fused_pow2 rows epilogue: per lane slot {row or -1, first column, count, weight block offset}
    const void *lane_w;       // T: weights, lane-major per warp block: [block offset + i * 32 + lane]
    int n_lane_slots;         // multiple of 32 (0: table absent)
    // ---- fused stereo-pair binaural cues (r2c_fused_pow2, complex-STFT plans): a tile holds FT/2 frames of each channel
    const void *samples_b;    // right channel [n_clips][clip_stride]; null = ordinary launch
    int cue;                  // sgx_binaural_cue
    int cue_start_bin, cue_band;
    unsigned cue_power;
    int cue_wrapped;
    double cue_bin_width;
    const void *sched;    // r2c_fused_n400: host-built quad schedule of the sparse mapping (see sgx_api.cu), else null
    const void *dense_t;  // T[out_len][n_bins]  the dense matrix transposed (chroma; null otherwise)
    int dense_c0, dense_c1;   // columns outside [dense_c0, dense_c1) of the dense matrix are exactly zero in every row
    int chroma_norm;      // sgx_chroma_norm (SGX_MAP_CHROMA: dense holds the 12 x out_len chroma filterbank)
    // ---- amplitude scaling (AmplitudeScaling :2043-2081)
    int amp;              // sgx_amp
    int apply_db;         // amp == Decibels && db_floor.is_some()
    double eps;           // 10^(floor_db/10), cast to T in the kernel (:2028)
    // ---- output kind
    int output;           // sgx_output
    int n_mfcc;           // DCT rows computed
    int mfcc_row0;        // 1 when c0 is dropped (!include_c0 && n_mfcc > 1), else 0
    const void *dct;      // T[n_mfcc][n_bins]
    const void *lifter;   // T[n_mfcc]
    const void *dct_folded;   // T[tasks][n_bins/2][4]: even/odd-symmetric half basis, 4 coefficients per task (n_bins even), or null
    int dct_tasks;            // number of (parity, 4-coefficient group) tasks
    // ---- shared-memory geometry chosen by the host
    int buf_elems;        // complex elements per ping-pong buffer (generic family)
    int vec_ok;           // input base / strides allow the 8- or 16-byte vector (async) load path
    // ---- fast divisors of the generic family's index math (all uniform per launch)
    FastDiv fd_L, fd_out_len, fd_FT, fd_r0, fd_stage_B[kFdStages], fd_stage_cur[kFdStages];   // fd_r0: radix of stage 0 (the only stage that can be a cofactor stage)
    int frame_stride;     // complex elements between frames in a buffer (>= L+1)
    int tile_stride;      // T elements between frames in the power / mel tile
    // ---- r2c_fused_pow2, rows-per-thread epilogue: the lane-major weights staged into shared memory by cp.async at kernel start
    int lane_w_smem;      // byte offset of the staged copy in dynamic shared memory (0: read the weights from global memory)
    int lane_w_bytes;     // its size (a multiple of 16)
    int l2_ahead;         // r2c_fused_pow2: prefetch the samples of CTA blockIdx.x + l2_ahead into L2 at kernel start (0: off)
};

template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

template <typename T> __device__ __forceinline__ typename Cplx<T>::type mk(T a, T b) {
    typename Cplx<T>::type r; r.x = a; r.y = b; return r;
}
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }

__device__ __forceinline__ float t_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double t_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float t_log10(float v) { return log10f(v); }
__device__ __forceinline__ double t_log10(double v) { return log10(v); }
// 10*log10(v) for the dB epilogue. f32: one MUFU.LG2 + one multiply (absolute error of lg2.approx is ~2e-7, i.e.
// < 1e-6 dB, far inside the 1e-3 dB budget; arguments are clamped to eps > 0 so no denormal / zero cases arise).
__device__ __forceinline__ float t_ten_log10(float v) { return 3.01029995663981195f * __log2f(v); }
__device__ __forceinline__ double t_ten_log10(double v) { return 10.0 * log10(v); }
__device__ __forceinline__ float t_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double t_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double t_fma(double a, double b, double c) { return fma(a, b, c); }
// products/sums that must not be contracted into an FMA (the reference writes `acc += w * x`, :113)
__device__ __forceinline__ float t_mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double t_mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float t_add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double t_add_rn(double a, double b) { return __dadd_rn(a, b); }

// AmpScale::apply_from_power + apply_db_in_place (:1986-2037, :2068-2080)
template <typename T> __device__ __forceinline__ T amp_scale(T v, int amp, int apply_db, T eps) {
    if (amp == 1) v = t_sqrt(v);
    if (apply_db) v = t_ten_log10(t_max(v, eps));
    return v;
}

// The same for loops that scale many values: when the dB floor eps is a normal f32 (all but floors below about -379 dB), max(v, eps)
// is normal too and lg2.approx.ftz returns exactly what __log2f returns -- without its per-call denormal fix-up (a scale, a
// select and an add around the MUFU). fast_db is that launch-uniform condition; f64 has no such short cut.
__device__ __forceinline__ float amp_scale_fast(float v, int amp, int apply_db, float eps, bool fast_db) {
    if (amp == 1) v = sqrtf(v);
    if (apply_db) {
        if (fast_db) {
            float r;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(v, eps)));
            v = 3.01029995663981195f * r;
        } else {
            v = t_ten_log10(fmaxf(v, eps));
        }
    }
    return v;
}
__device__ __forceinline__ double amp_scale_fast(double v, int amp, int apply_db, double eps, bool) { return amp_scale<double>(v, amp, apply_db, eps); }

// apply_chroma_normalization (src/chroma.rs:406-453) on one frame's 12 pitch classes, in the reference's fold order
template <typename T> __device__ __forceinline__ void chroma_normalise(T (&c)[12], int norm) {
    T d = T(0);
    if (norm == SGX_CHROMANORM_L1) {
#pragma unroll
        for (int i = 0; i < 12; ++i) d = t_add_rn(d, c[i]);
    } else if (norm == SGX_CHROMANORM_L2) {
#pragma unroll
        for (int i = 0; i < 12; ++i) d = t_add_rn(d, t_mul_rn(c[i], c[i]));
        d = t_sqrt(d);
    } else if (norm == SGX_CHROMANORM_MAX) {
#pragma unroll
        for (int i = 0; i < 12; ++i) d = t_max(d, c[i]);
    } else {
        return;
    }
    if (d > T(0)) {
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = c[i] / d;
    }
}

}  // namespace sgx
