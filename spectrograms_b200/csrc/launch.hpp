// launch.hpp -- host-callable launchers of the kernel families (defined in the kernel_*.cu files).
#pragma once

#include <algorithm>

#include <cuda_runtime.h>

#include "../../include/sgx_b200.h"
#include "kparams.cuh"

namespace sgx {

// r2c_fused_generic: any n_fft (kernel_generic.cu)
cudaError_t launch_generic(const KParams &p, bool f64, size_t smem_bytes, cudaStream_t stream);

// r2c_fused_n400: n_fft = 400, hop = 160, f32, any spectrogram / MFCC output (kernel_fast400.cu).
// p.vec_ok says whether the input allows 8-byte async loads; window_f32 is the host copy of the plan window.
// sparse_table: serve mel / loghz rows from a shared-memory copy of the CSR table (needs rows <= max_sparse_rows and
// nnz <= max_sparse_nnz); sm_count sizes the persistent grid (2 CTAs per SM).
cudaError_t launch_fast400(const KParams &p, const float *window_f32, bool sparse_table, int n_quads, int padded_weights,
                           int sm_count, cudaStream_t stream);
size_t fast400_smem_bytes(int n_quads, int padded_weights);
bool fast400_sparse_fits(int n_quads, int padded_weights);
int fast400_max_scratch_rows();
int fast400_warps();

// r2c_fused_n400_tc: the same family with TMEM as the exchange medium and the filterbank on tcgen05 (kernel_n400_tc.cu).
// p.sched points at the device copy of the step blob built by the host (sgx_api.cu: build_tc_blob):
//   int n_steps, n_rounds, round_start[n_rounds + 1], pad to 4 ints; int4 {A column, D column, byte offset of the step's B
//   tiles, accumulate | N << 8}[n_steps]; float B[b_floats] (per step N rows x 8 bins as K-major core matrices: hi tile,
//   then lo tile; N = 16 or 64).
cudaError_t launch_fast400_tc(const KParams &p, const float *window_f32, int n_steps, int n_rounds, size_t b_floats, int sm_count,
                              cudaStream_t stream);
size_t fast400_tc_smem_bytes(int n_steps, int n_rounds, size_t b_floats);
bool fast400_tc_fits(int n_steps, int n_rounds, size_t b_floats);

// r2c_fused_n400_tm: the same family with TMEM as the exchange medium of the two FFT passes and four independent
// 32-frame groups per SM (kernel_n400_tm.cu). Sparse mel / loghz spectrogram outputs. p.sched points at the quad blob
// built for group_warps (4, 5 or 6) warps per group with every row of a quad zero-padded to the quad's longest row
// (rows may read up to fast400_tm_pad_rows() zero rows behind bin 200).
cudaError_t launch_fast400_tm(const KParams &p, const float *window_f32, int n_quads, int padded_weights, int group_warps, int sm_count,
                              cudaStream_t stream);
size_t fast400_tm_smem_bytes(int n_quads, int padded_weights);
bool fast400_tm_fits(int n_quads, int padded_weights);
int fast400_tm_max_group_warps();
int fast400_tm_pad_rows();

// r2c_fused_pow2: n_fft = 256 .. 8192 (powers of two), f32 / f64 (kernel_pow2.cu). p.FT, p.frame_stride, p.tile_stride and
// p.tiles_per_clip must be set from the helpers below; p.vec_ok is the "vector loads allowed" flag.
bool pow2_supported(size_t n_fft);
int pow2_frames_per_tile(size_t n_fft, bool f64);
int pow2_frame_elems(size_t n_fft, bool f64);
int pow2_min_blocks(size_t n_fft, bool f64);                        // CTAs per SM of the forward kernel's launch bounds
size_t pow2_bulk_stage_bytes(size_t n_fft, size_t hop, bool f64);   // extra smem of the cp.async.bulk staged variant (0: none)
cudaError_t launch_pow2(const KParams &p, bool f64, size_t smem_bytes, cudaStream_t stream);

// r2c_fused_mixed: even n_fft = 2 R1 R2 R3 with small prime factors that no other fast family serves (kernel_mixed.cu):
// 32-frame tiles, lane = frame, in-place register-radix stages. p.tw / p.post / p.window are the generic family's tables.
bool mixed_supported(size_t n_fft, bool f64);
size_t mixed_smem_bytes(size_t n_fft, bool f64);
int mixed_tile_frames();
int mixed_max_scratch_rows(size_t n_fft);
cudaError_t launch_mixed(const KParams &p, bool f64, cudaStream_t stream);

// standalone mfcc_from_log_mel (kernel_mfcc.cu): log_mel [n_clips][n_mels][n_frames] -> out [n_clips][rows][n_frames]
cudaError_t launch_mfcc(bool f64, const void *log_mel, void *out, long long n_clips, int n_mels, long long n_frames,
                        int n_mfcc, int row0, const void *dct, const void *lifter, cudaStream_t stream);

// the same as a tcgen05 GEMM (kernel_mfcc_tc.cu, f32): frames x n_mels times the DCT basis, 3xTF32, FP32 accumulation in tensor
// memory, lifter / c0 drop fused. blob: per 8-mel K step an N x 8 hi tile and an N x 8 lo tile of the basis as K-major core
// matrices (float (n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + k % 4 inside a tile; N = n_mfcc rounded up to 16, mels zero padded
// to a multiple of 8), then the n_mfcc lifter weights.
bool mfcc_tc_supported(int n_mels, int n_mfcc);
int mfcc_tc_padded_mels(int n_mels);
int mfcc_tc_padded_coeffs(int n_mfcc);
size_t mfcc_tc_blob_floats(int n_mels, int n_mfcc);
// log_mel rows are in_row_stride floats apart (>= n_frames), clips n_mels rows apart.
cudaError_t launch_mfcc_tc(const float *log_mel, long long in_row_stride, float *out, long long n_clips, int n_mels, long long n_frames, int n_mfcc,
                           int row0, const float *blob, int sm_count, cudaStream_t stream);

// the same kernel as a dense filterbank row block (ERB): linear power spectrogram [n_clips][n_cols][n_frames] -> rows [row_base,
// row_base + n_rows) of out, amplitude scaling fused; n_rows <= dense_tc_max_rows(n_cols) (0: the column count does not fit).
// blob: build as for dct2_lifter_tc with the block's weights as the basis.
int dense_tc_max_rows(int n_cols);
cudaError_t launch_dense_tc(const float *power, long long in_row_stride, float *out, long long out_clip_stride, long long n_clips, int n_cols,
                            long long n_frames, int row_base, int n_rows, const float *blob, int amp, int apply_db, float eps, int sm_count,
                            cudaStream_t stream);

// standalone chromagram_from_spectrogram (kernel_chroma.cu): spec [n_clips][n_bins][n_frames] -> out [n_clips][12][n_frames];
// w_transposed is the chroma filterbank as T[n_bins][12], non-zero only for bins in [k0, k1)
cudaError_t launch_chroma(bool f64, const void *spec, void *out, long long n_clips, int n_bins, long long n_frames,
                          const void *w_transposed, int norm, int k0, int k1, cudaStream_t stream);

// interaural cue spectrograms from two complex STFTs [n_pairs][n_bins][n_frames] -> [n_pairs][stop_bin - start_bin][n_frames]
// (kernel_binaural.cu); cue is an sgx_binaural_cue
cudaError_t launch_binaural(bool f64, int cue, const void *left, const void *right, void *out, long long n_pairs, int n_bins,
                            long long n_frames, int start_bin, int stop_bin, double bin_width, unsigned power, int wrapped,
                            cudaStream_t stream);

// inverse path (kernel_inverse.cu): Hermitian STFT [n_clips][n_fft/2+1][n_frames] -> (windowed) time frames
// [n_clips][n_frames][n_fft] with the generic family's geometry (p.FT, p.frame_stride, p.buf_elems, p.tiles_per_clip set by
// the caller), then the overlap-add gather into out [n_clips][out_len] (trim = samples dropped at the front)
cudaError_t launch_c2r_frames(const KParams &p, bool f64, size_t smem, const void *stft, void *frames_out, long long n_clips,
                              long long n_frames, int apply_window, cudaStream_t stream);
// the same on the register radix-16 passes of r2c_fused_pow2 (kernel_pow2.cu); p.tiles_per_clip = ceil(n_frames / FT)
cudaError_t launch_c2r_pow2(const KParams &p, bool f64, const void *stft, void *frames_out, long long n_clips,
                            long long n_frames, int apply_window, cudaStream_t stream);
int pow2_c2r_frames_per_tile(size_t n_fft);
// istft fused on a halo tile (kernel_pow2.cu: k_istft_pow2): Hermitian STFT -> out [n_clips][out_len] directly. halo =
// ceil(n_fft / hop) - 1 < pow2_c2r_frames_per_tile(n_fft); p.tiles_per_clip = ceil(ceil(total / hop) / (FT - halo))
cudaError_t launch_istft_pow2(const KParams &p, bool f64, const void *stft, void *out, long long n_clips, long long n_frames, int halo,
                              long long out_len, long long trim, cudaStream_t stream);      // tiles of the inverse kernel: p.tiles_per_clip = ceil(n_frames / this)
cudaError_t launch_ola_gather(bool f64, const void *frames, const void *window, void *out, long long n_clips, long long n_frames,
                              int n_fft, int hop, long long out_len, long long trim, cudaStream_t stream);

}  // namespace sgx
