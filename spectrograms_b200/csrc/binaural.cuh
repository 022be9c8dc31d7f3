// binaural.cuh -- the per-element cue arithmetic of src/binaural.rs (pow_mag :60-83, np_mod :85-87, magphase :106-180, the
// bodies of compute_{itd,ipd,ild,ilr}_spectrogram :528-545, :875-889, :1229-1231, :1572-1580), shared by the standalone
// binaural_cues kernel and the fused stereo-pair epilogue of r2c_fused_pow2. Expression by expression in T.
#pragma once

#include "kparams.cuh"

namespace sgx {

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

// one (bin, frame) element: l / r are the two channels' STFT values, bin is the absolute FFT bin
template <typename T>
__device__ __forceinline__ T binaural_cue(int cue, T lre, T lim, T rre, T rim, int bin, T bin_width, unsigned power, int wrapped) {
    T ml, plr, pli, mr, prr, pri;
    magphase<T>(lre, lim, (cue == SGX_CUE_ITD) ? power : 1u, ml, plr, pli);
    magphase<T>(rre, rim, (cue == SGX_CUE_ITD) ? power : 1u, mr, prr, pri);
    const T pi = t_pi<T>(), two_pi = T(2) * pi;
    if (cue == SGX_CUE_ITD) {                                   // :528-545
        T o = T(0);
        if (ml + mr > T(0)) {
            const T diff = t_atan2(pli, plr) - t_atan2(pri, prr);
            const T w = np_mod<T>(diff + pi, two_pi) - pi;
            o = w / (two_pi * bin_width * static_cast<T>(bin));
        }
        return o;
    }
    if (cue == SGX_CUE_IPD) {                                   // :875-889
        const T diff = t_atan2(pli, plr) - t_atan2(pri, prr);
        return wrapped ? np_mod<T>(diff + pi, two_pi) - pi : diff;
    }
    T o = static_cast<T>(NAN);                                  // Array2::from_elem(.., T::nan()) :1212, :1555
    if (ml + mr > T(0) && ml > T(0) && mr > T(0)) {
        const T ratio = mr / ml;
        if (cue == SGX_CUE_ILD) o = T(-20) * t_log10(ratio);                      // :1229-1231
        else o = ratio < T(1) ? T(1) - ratio : -(T(1) - T(1) / ratio);            // :1572-1580
    }
    return o;
}

}  // namespace sgx
