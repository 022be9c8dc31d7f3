// kernel_chroma.cu -- "chroma_rows": standalone chromagram_from_spectrogram (src/chroma.rs:365-404) for callers that
// already hold a (bins, frames) spectrogram in device memory. (The fused path, SGX_MAP_CHROMA, never materialises the
// magnitude spectrogram.)
//
// One thread = one frame: it streams down the frame's column once (lanes run along frames, so every read is a
// coalesced 128/256-byte run of one spectrogram row), keeping the 12 pitch-class sums in registers. The filterbank is
// passed transposed, wT[bin][12], so the 12 weights of a bin are one warp-uniform 48/96-byte broadcast. Each row
// accumulates `sum += T(w) * x` in ascending bin order without FMA, exactly like the reference loop (:384-394); then the
// per-frame normalisation (:406-453) and twelve coalesced row stores.
#include "launch.hpp"

namespace sgx {
namespace {

template <typename T>
__global__ void __launch_bounds__(128) k_chroma_rows(const T *__restrict__ spec, T *__restrict__ out, int n_bins,
                                                     long long n_frames, const T *__restrict__ wT, int norm, int tiles_per_clip, int k0, int k1) {
    const int clip = blockIdx.x / tiles_per_clip;
    const long long f = static_cast<long long>(blockIdx.x - clip * tiles_per_clip) * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const T *src = spec + static_cast<long long>(clip) * n_bins * n_frames + f;
    T c[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = T(0);
#pragma unroll 2
    for (int k = k0; k < k1; ++k) {                   // bins outside [k0, k1) carry exactly zero weight in every row
        const T x = src[static_cast<long long>(k) * n_frames];
        const T *w = wT + 12 * k;
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = t_add_rn(c[i], t_mul_rn(__ldg(w + i), x));
    }
    chroma_normalise<T>(c, norm);
    T *dst = out + static_cast<long long>(clip) * 12 * n_frames + f;
#pragma unroll
    for (int i = 0; i < 12; ++i) dst[static_cast<long long>(i) * n_frames] = c[i];
}

}  // namespace

cudaError_t launch_chroma(bool f64, const void *spec, void *out, long long n_clips, int n_bins, long long n_frames,
                          const void *w_transposed, int norm, int k0, int k1, cudaStream_t stream) {
    const long long tiles = (n_frames + 127) / 128;
    const long long grid = n_clips * tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    if (f64)
        k_chroma_rows<double><<<static_cast<unsigned>(grid), 128, 0, stream>>>(static_cast<const double *>(spec), static_cast<double *>(out),
                                                                              n_bins, n_frames, static_cast<const double *>(w_transposed),
                                                                              norm, static_cast<int>(tiles), k0, k1);
    else
        k_chroma_rows<float><<<static_cast<unsigned>(grid), 128, 0, stream>>>(static_cast<const float *>(spec), static_cast<float *>(out),
                                                                             n_bins, n_frames, static_cast<const float *>(w_transposed),
                                                                             norm, static_cast<int>(tiles), k0, k1);
    return cudaGetLastError();
}

}  // namespace sgx
