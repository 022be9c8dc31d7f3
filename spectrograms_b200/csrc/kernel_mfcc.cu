// kernel_mfcc.cu -- "dct2_lifter": standalone mfcc_from_log_mel (src/mfcc.rs:224-273) for callers that already
// hold a log-mel spectrogram in device memory. (The fused path, SGX_OUT_MFCC, never materialises the log-mel.)
//
// One CTA = 32 consecutive frames of one clip. The (n_mels x 32) log-mel tile is staged in shared memory with
// coalesced row reads; each thread then owns one (coefficient, frame) pair, accumulating with fma in ascending mel
// order exactly like dct_ii (src/mfcc.rs:282-289); lanes run along frames so reads of the tile are conflict free,
// the basis row is warp-uniform, and stores are 128-byte runs of one output row.
#include "launch.hpp"

namespace sgx {
namespace {

constexpr int kFT = 32;

template <typename T>
__global__ void __launch_bounds__(256) k_dct2_lifter(const T *__restrict__ log_mel, T *__restrict__ out, int n_mels,
                                                     long long n_frames, int n_mfcc, int row0, const T *__restrict__ dct,
                                                     const T *__restrict__ lifter, int tiles_per_clip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);   // [n_mels][kFT + 1]
    const int clip = blockIdx.x / tiles_per_clip;
    const int t = blockIdx.x - clip * tiles_per_clip;
    const long long f0 = static_cast<long long>(t) * kFT;
    const int nf = (n_frames - f0) < kFT ? static_cast<int>(n_frames - f0) : kFT;
    const T *src = log_mel + static_cast<long long>(clip) * n_mels * n_frames + f0;
    for (int idx = threadIdx.x; idx < n_mels * kFT; idx += blockDim.x) {
        const int i = idx / kFT, f = idx - i * kFT;
        tile[i * (kFT + 1) + f] = f < nf ? src[static_cast<long long>(i) * n_frames + f] : T(0);
    }
    __syncthreads();
    const int rows = n_mfcc - row0;
    T *dst = out + static_cast<long long>(clip) * rows * n_frames + f0;
    for (int idx = threadIdx.x; idx < rows * kFT; idx += blockDim.x) {
        const int r = idx / kFT, f = idx - r * kFT;
        if (f >= nf) continue;
        const int c = r + row0;
        const T *b = dct + static_cast<long long>(c) * n_mels;
        T acc = T(0);
        for (int i = 0; i < n_mels; ++i) acc = t_fma(tile[i * (kFT + 1) + f], __ldg(b + i), acc);
        dst[static_cast<long long>(r) * n_frames + f] = acc * __ldg(lifter + c);
    }
}

}  // namespace

cudaError_t launch_mfcc(bool f64, const void *log_mel, void *out, long long n_clips, int n_mels, long long n_frames,
                        int n_mfcc, int row0, const void *dct, const void *lifter, cudaStream_t stream) {
    const long long tiles = (n_frames + kFT - 1) / kFT;
    const long long grid = n_clips * tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const size_t smem = static_cast<size_t>(n_mels) * (kFT + 1) * (f64 ? 8 : 4);
    cudaError_t e;
    if (f64) {
        e = cudaFuncSetAttribute(k_dct2_lifter<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_dct2_lifter<double><<<static_cast<unsigned>(grid), 256, smem, stream>>>(
            static_cast<const double *>(log_mel), static_cast<double *>(out), n_mels, n_frames, n_mfcc, row0,
            static_cast<const double *>(dct), static_cast<const double *>(lifter), static_cast<int>(tiles));
    } else {
        e = cudaFuncSetAttribute(k_dct2_lifter<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_dct2_lifter<float><<<static_cast<unsigned>(grid), 256, smem, stream>>>(
            static_cast<const float *>(log_mel), static_cast<float *>(out), n_mels, n_frames, n_mfcc, row0,
            static_cast<const float *>(dct), static_cast<const float *>(lifter), static_cast<int>(tiles));
    }
    return cudaGetLastError();
}

}  // namespace sgx
