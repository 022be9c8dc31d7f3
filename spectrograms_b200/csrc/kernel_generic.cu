// kernel_generic.cu -- "r2c_fused_generic": the any-n_fft kernel family (f32 and f64).
//
// One CTA = one tile of FT consecutive frames of one clip. Per tile, entirely in shared memory:
//   1. gather + zero pad + window      (src/spectrogram.rs:1301-1320; packs even n_fft as N/2 complex samples)
//   2. Stockham autosort FFT stages    (radix 4/2/3/5 butterflies, one butterfly per thread, plus at most one
//                                       "cofactor" stage that evaluates an r-point DFT directly, one output per
//                                       thread -- this keeps the plan total for every n_fft, e.g. primes)
//   3. real-input post pass            (split of the packed spectrum, realfft-style) -> power or complex tile
//   4. fused epilogue                  (epilogue.cuh)
// The complex spectrum never leaves the SM unless the caller asked for it (SGX_OUT_COMPLEX_STFT).
//
// This family is the correctness backbone; the specialised families (kernel_fast_*.cu) reuse its epilogue and are
// checked against it as well as against the CPU oracle.
#include "epilogue.cuh"
#include "launch.hpp"
#include "stockham.cuh"

namespace sgx {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_r2c_fused_generic(const __grid_constant__ KParams p) {
    using C = typename Cplx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *bufA = reinterpret_cast<C *>(smem_raw);
    C *bufB = bufA + p.buf_elems;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const long long f0 = p.frame_begin + static_cast<long long>(tile) * p.FT;
    const long long rem = p.frame_begin + p.frames_todo - f0;
    const int nf = rem < p.FT ? static_cast<int>(rem) : p.FT;
    const int L = p.L, FS = p.frame_stride;

    const T *x = static_cast<const T *>(p.samples) + static_cast<long long>(clip) * p.clip_stride;
    const T *win = static_cast<const T *>(p.window);

    // ---- 1. gather / pad / window / pack
    for (int idx = tid; idx < nf * L; idx += nthr) {
        const int f = fd_div(idx, p.fd_L);
        const int n = idx - f * L;
        const long long base = (f0 + f) * p.hop - p.pad;
        C z;
        if (p.even) {
            const long long s0 = base + 2 * n, s1 = s0 + 1;
            const T a = (s0 >= 0 && s0 < p.n_samples) ? x[s0] : T(0);
            const T b = (s1 >= 0 && s1 < p.n_samples) ? x[s1] : T(0);
            z.x = a * __ldg(win + 2 * n);
            z.y = b * __ldg(win + 2 * n + 1);
        } else {
            const long long s0 = base + n;
            const T a = (s0 >= 0 && s0 < p.n_samples) ? x[s0] : T(0);
            z.x = a * __ldg(win + n);
            z.y = T(0);
        }
        bufA[f * FS + n] = z;
    }
    __syncthreads();

    // ---- 2. Stockham stages (stockham.cuh)
    C *outb = nullptr;
    C *in = stockham_stages<T>(p, bufA, bufB, nf, &outb);
    // `in` now holds Z (packed spectrum, even n_fft) or the full spectrum (odd n_fft); `outb` is free

    // ---- 3. post pass
    const C *Z = in;
    if (p.output == SGX_OUT_COMPLEX_STFT) {
        C *S = outb;
        const C *post = static_cast<const C *>(p.post);
        for (int idx = tid; idx < nf * p.out_len; idx += nthr) {
            const int f = fd_div(idx, p.fd_out_len);
            const int k = idx - f * p.out_len;
            const C *zf = Z + f * FS;
            C X;
            if (!p.even) {
                X = zf[k];
            } else if (k == 0) {
                X = mk<T>(zf[0].x + zf[0].y, T(0));
            } else if (k == L) {
                X = mk<T>(zf[0].x - zf[0].y, T(0));
            } else {
                const C a = zf[k], b = zf[L - k];
                const C ev = mk<T>(T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y));
                const C od = mk<T>(T(0.5) * (a.y + b.y), T(0.5) * (b.x - a.x));
                X = cadd(ev, cmul(od, post[k]));
            }
            S[f * FS + k] = X;
        }
        __syncthreads();
        epilogue_complex<T>(p, S, clip, f0, nf);
        return;
    }

    T *P = reinterpret_cast<T *>(outb);
    {
        const C *post = static_cast<const C *>(p.post);
        const int ts = p.tile_stride;
        for (int idx = tid; idx < nf * p.out_len; idx += nthr) {
            const int f = fd_div(idx, p.fd_out_len);
            const int k = idx - f * p.out_len;
            const C *zf = Z + f * FS;
            C X;
            if (!p.even) {
                X = zf[k];
            } else if (k == 0) {
                X = mk<T>(zf[0].x + zf[0].y, T(0));
            } else if (k == L) {
                X = mk<T>(zf[0].x - zf[0].y, T(0));
            } else {
                const C a = zf[k], b = zf[L - k];
                const C ev = mk<T>(T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y));
                const C od = mk<T>(T(0.5) * (a.y + b.y), T(0.5) * (b.x - a.x));
                X = cadd(ev, cmul(od, post[k]));
            }
            P[f * ts + k] = X.x * X.x + X.y * X.y;   // norm_sqr (:1332-1334)
        }
    }
    __syncthreads();
    // the buffer that held Z is free now -> scratch for the log-mel tile
    epilogue_from_power<T>(p, P, reinterpret_cast<T *>(in), clip, f0, nf);
}

}  // namespace

cudaError_t launch_generic(const KParams &p, bool f64, size_t smem_bytes, cudaStream_t stream) {
    const long long grid = static_cast<long long>(p.n_clips) * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (f64) {
        e = cudaFuncSetAttribute(k_r2c_fused_generic<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
        k_r2c_fused_generic<double><<<static_cast<unsigned>(grid), 256, smem_bytes, stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_r2c_fused_generic<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
        if (e != cudaSuccess) return e;
        k_r2c_fused_generic<float><<<static_cast<unsigned>(grid), 256, smem_bytes, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace sgx
