// kernel_binaural.cu -- "binaural_cues": interaural ITD / IPD / ILD / ILR spectrograms from a pair of complex STFTs
// (src/binaural.rs: magphase :106-180, compute_itd_spectrogram :472-580, compute_ipd_spectrogram :830-917,
// compute_ild_spectrogram :1187-1262, compute_ilr_spectrogram :1530-1620).
//
// Purely element-wise on the band [start_bin, stop_bin): one thread = one (bin, frame) of one stereo pair, frames
// fastest across lanes so that both complex reads and the store are coalesced runs of one STFT row. The arithmetic
// follows the reference expression by expression in T (mul_add for |c|^2, recip for the phase normalisation, atan2 of
// the unit phasor, np_mod wrap), so that the only differences left are the last-ulp ones of the math library.
#include "launch.hpp"

namespace sgx {
namespace {

__device__ __forceinline__ float t_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double t_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float t_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double t_fmod(double a, double b) { return fmod(a, b); }
template <typename T> __device__ __forceinline__ T t_pi();
template <> __device__ __forceinline__ float t_pi<float>() { return 3.14159274101257324f; }       // f32::consts::PI
template <> __device__ __forceinline__ double t_pi<double>() { return 3.141592653589793; }       // f64::consts::PI

// pow_mag (src/binaural.rs:60-83)
template <typename T> __device__ __forceinline__ T pow_mag(T mag, T mag_sq, unsigned power) {
    switch (power) {
        case 1: return mag;
        case 2: return mag_sq;
        case 3: return mag_sq * mag;
        case 4: return mag_sq * mag_sq;
        default: {
            T base = mag, acc = T(1);
            unsigned e = power;
            while (e > 0) {
                if (e & 1u) acc *= base;
                e >>= 1;
                if (e > 0) base *= base;
            }
            return acc;
        }
    }
}

// magphase (:106-180): |c|^power and the unit phasor; (0, 0) -> magnitude 0, phasor (1, 0)
template <typename T> __device__ __forceinline__ void magphase(T re, T im, unsigned power, T &m, T &pr, T &pi) {
    const T mag_sq = t_fma(re, re, im * im);
    if (mag_sq == T(0)) { m = T(0); pr = T(1); pi = T(0); return; }
    const T mag = t_sqrt(mag_sq);
    m = pow_mag(mag, mag_sq, power);
    const T inv = T(1) / mag;
    pr = re * inv;
    pi = im * inv;
}

template <typename T> __device__ __forceinline__ T np_mod(T x, T m) { return t_fmod(t_fmod(x, m) + m, m); }   // :85-87

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
    T ml, plr, pli, mr, prr, pri;
    magphase<T>(l.x, l.y, (cue == SGX_CUE_ITD) ? power : 1u, ml, plr, pli);
    magphase<T>(r.x, r.y, (cue == SGX_CUE_ITD) ? power : 1u, mr, prr, pri);
    const T pi = t_pi<T>(), two_pi = T(2) * pi;
    T o;
    if (cue == SGX_CUE_ITD) {                                   // :528-545
        o = T(0);
        if (ml + mr > T(0)) {
            const T diff = t_atan2(pli, plr) - t_atan2(pri, prr);
            const T w = np_mod<T>(diff + pi, two_pi) - pi;
            o = w / (two_pi * bin_width * static_cast<T>(start_bin + b));
        }
    } else if (cue == SGX_CUE_IPD) {                            // :875-889
        const T diff = t_atan2(pli, plr) - t_atan2(pri, prr);
        o = wrapped ? np_mod<T>(diff + pi, two_pi) - pi : diff;
    } else {
        o = static_cast<T>(NAN);                                // Array2::from_elem(.., T::nan()) :1212, :1555
        if (ml + mr > T(0) && ml > T(0) && mr > T(0)) {
            const T ratio = mr / ml;
            if (cue == SGX_CUE_ILD) o = T(-20) * t_log10(ratio);                      // :1229-1231
            else o = ratio < T(1) ? T(1) - ratio : -(T(1) - T(1) / ratio);            // :1572-1580
        }
    }
    out[idx] = o;
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
