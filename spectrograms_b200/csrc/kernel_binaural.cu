// kernel_binaural.cu -- "binaural_cues": interaural ITD / IPD / ILD / ILR spectrograms from a pair of complex STFTs
// (src/binaural.rs: magphase :106-180, compute_itd_spectrogram :472-580, compute_ipd_spectrogram :830-917,
// compute_ild_spectrogram :1187-1262, compute_ilr_spectrogram :1530-1620).
//
// Purely element-wise on the band [start_bin, stop_bin): one thread = one (bin, frame) of one stereo pair, frames
// fastest across lanes so that both complex reads and the store are coalesced runs of one STFT row. The arithmetic
// follows the reference expression by expression in T (mul_add for |c|^2, recip for the phase normalisation, atan2 of
// the unit phasor, np_mod wrap), so that the only differences left are the last-ulp ones of the math library.
#include "binaural.cuh"
#include "launch.hpp"

namespace sgx {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_binaural_cues(const typename Cplx<T>::type *__restrict__ left,
                                                       const typename Cplx<T>::type *__restrict__ right, T *__restrict__ out,
                                                       int cue, int n_bins, long long n_frames, int start_bin, int band,
                                                       T bin_width, unsigned power, int wrapped, long long total) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long per_pair = static_cast<long long>(band) * n_frames;
    const long long pair = idx / per_pair, rem = idx - pair * per_pair;
    const int b = static_cast<int>(rem / n_frames);
    const long long f = rem - static_cast<long long>(b) * n_frames;
    const long long src = (pair * n_bins + start_bin + b) * n_frames + f;
    const typename Cplx<T>::type l = left[src], r = right[src];
    out[idx] = binaural_cue<T>(cue, l.x, l.y, r.x, r.y, start_bin + b, bin_width, power, wrapped);
}

}  // namespace

cudaError_t launch_binaural(bool f64, int cue, const void *left, const void *right, void *out, long long n_pairs, int n_bins,
                            long long n_frames, int start_bin, int stop_bin, double bin_width, unsigned power, int wrapped,
                            cudaStream_t stream) {
    const int band = stop_bin - start_bin;
    const long long total = n_pairs * band * n_frames;
    if (total <= 0) return cudaSuccess;
    const long long grid = (total + 255) / 256;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    if (f64)
        k_binaural_cues<double><<<static_cast<unsigned>(grid), 256, 0, stream>>>(
            static_cast<const double2 *>(left), static_cast<const double2 *>(right), static_cast<double *>(out), cue, n_bins, n_frames,
            start_bin, band, bin_width, power, wrapped, total);
    else
        k_binaural_cues<float><<<static_cast<unsigned>(grid), 256, 0, stream>>>(
            static_cast<const float2 *>(left), static_cast<const float2 *>(right), static_cast<float *>(out), cue, n_bins, n_frames,
            start_bin, band, static_cast<float>(bin_width), power, wrapped, total);
    return cudaGetLastError();
}

}  // namespace sgx
