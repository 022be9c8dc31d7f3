// kernel_inverse.cu -- the step on the other side of the path: irfft / istft
// (src/spectrogram.rs:4789-4811, :4813-4911; C2R plan src/fft_backend.rs:509-567).
//
//   c2r_frames   one CTA = a tile of frames of one clip: Hermitian spectrum -> time frame, times the synthesis window.
//                Even n_fft runs the packed inverse: E = (X[k] + conj(X[M-k])) / 2, O = (X[k] - conj(X[M-k])) / 2 * W_n^{-k},
//                Z = E + i O, z = IFFT_M(Z), x[2m] = Re z[m], x[2m+1] = Im z[m]; the inverse FFT is the forward Stockham
//                transform of the generic family on conj(Z) (stockham.cuh), conjugated and scaled by 1/M (the reference
//                scales realfft's unnormalised C2R by 1/n_fft, :559-563 -- the same true inverse). Odd n_fft transforms
//                the full Hermitian extension. Like realfft, the imaginary parts of the DC (and Nyquist) bins are ignored.
//   ola_gather   one thread = one output sample: adds the windowed frames that cover it in ascending frame order and
//                divides by the accumulated squared window where it exceeds 1e-10 -- the reference's overlap-add
//                (:4868-4886) restated as a gather, so it needs no atomics and sums in the reference's order.
// The windowed frames make one round trip through a scratch buffer in HBM (n_fft / hop times the signal size); a tile
// kernel with a frame halo would avoid it and is left for the next round.
#include "launch.hpp"
#include "stockham.cuh"

namespace sgx {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_c2r_frames(const __grid_constant__ KParams p, const typename Cplx<T>::type *__restrict__ stft,
                                                    T *__restrict__ frames_out, long long n_frames, int apply_window) {
    using C = typename Cplx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *bufA = reinterpret_cast<C *>(smem_raw);
    C *bufB = bufA + p.buf_elems;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const long long f0 = static_cast<long long>(tile) * p.FT;
    const long long rem = n_frames - f0;
    const int nf = rem < p.FT ? static_cast<int>(rem) : p.FT;
    const int L = p.L, FS = p.frame_stride, n = p.n_fft, bins = p.out_len;
    const C *X = stft + static_cast<long long>(clip) * bins * n_frames + f0;      // X[k * n_frames + f]
    const C *post = static_cast<const C *>(p.post);
    // ---- pre pass: conj of the sequence whose forward FFT is the conjugated inverse; frames fastest for coalescing
    for (int idx = tid; idx < nf * L; idx += nthr) {
        const int k = idx / nf, f = idx - k * nf;
        C v;
        if (p.even) {                                   // L = M = n / 2
            C a = X[static_cast<long long>(k) * n_frames + f];
            C b = X[static_cast<long long>(L - k) * n_frames + f];
            if (k == 0) { a.y = T(0); b.y = T(0); }     // DC and Nyquist are real (realfft zeroes their imaginary parts)
            const C e = mk<T>(T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y));          // (a + conj b) / 2
            const C d = mk<T>(T(0.5) * (a.x - b.x), T(0.5) * (a.y + b.y));          // (a - conj b) / 2
            const C w = post[k];                                                      // W_n^k; W_n^{-k} is its conjugate
            const C o = mk<T>(d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y);
            v = mk<T>(e.x - o.y, -(e.y + o.x));                                       // conj(E + i O)
        } else {                                        // L = n: Hermitian extension
            if (k <= n / 2) {
                const C a = X[static_cast<long long>(k) * n_frames + f];
                v = mk<T>(a.x, k == 0 ? T(0) : -a.y);
            } else {
                const C a = X[static_cast<long long>(n - k) * n_frames + f];
                v = mk<T>(a.x, a.y);                                                  // conj(conj(X[n-k]))
            }
        }
        bufA[f * FS + k] = v;
    }
    __syncthreads();
    C *other = nullptr;
    const C *Y = stockham_stages<T>(p, bufA, bufB, nf, &other);
    // ---- conj, scale, (window), store time frames [clip][frame][i]
    const T scale = T(1) / static_cast<T>(L);
    const T *win = static_cast<const T *>(p.window);
    T *dst = frames_out + (static_cast<long long>(clip) * n_frames + f0) * n;
    for (int idx = tid; idx < nf * n; idx += nthr) {
        const int f = idx / n, i = idx - f * n;
        T x;
        if (p.even) {
            const C y = Y[f * FS + (i >> 1)];
            x = ((i & 1) ? -y.y : y.x) * scale;
        } else {
            x = Y[f * FS + i].x * scale;
        }
        if (apply_window) x = x * __ldg(win + i);                                     // time_frame[i] *= window[i] (:4868-4870)
        dst[static_cast<long long>(f) * n + i] = x;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_ola_gather(const T *__restrict__ frames, const T *__restrict__ win, T *__restrict__ out,
                                                    long long n_frames, int n, int hop, long long out_len, long long trim,
                                                    long long total_len, long long n_clips) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n_clips * out_len) return;
    const long long clip = idx / out_len, o = idx - clip * out_len;
    const long long pos = o + trim;                       // position in the untrimmed overlap-add buffer
    long long f_lo = pos >= n ? (pos - n) / hop + 1 : 0;  // first frame with f * hop + n > pos
    long long f_hi = pos / hop;
    if (f_hi > n_frames - 1) f_hi = n_frames - 1;
    const T *fr = frames + clip * n_frames * n;
    T acc = T(0), norm = T(0);
    for (long long f = f_lo; f <= f_hi; ++f) {            // ascending frames: the reference's accumulation order (:4874-4881)
        const int i = static_cast<int>(pos - f * hop);
        const T w = __ldg(win + i);
        acc = t_add_rn(acc, fr[f * n + i]);
        norm = t_add_rn(norm, t_mul_rn(w, w));
    }
    if (norm > static_cast<T>(1e-10)) acc = acc / norm;   // :4885-4890
    out[idx] = acc;
    (void)total_len;
}

}  // namespace

cudaError_t launch_c2r_frames(const KParams &p, bool f64, size_t smem, const void *stft, void *frames_out, long long n_clips,
                              long long n_frames, int apply_window, cudaStream_t stream) {
    const long long grid = n_clips * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e;
    if (f64) {
        e = cudaFuncSetAttribute(k_c2r_frames<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_c2r_frames<double><<<static_cast<unsigned>(grid), 256, smem, stream>>>(p, static_cast<const double2 *>(stft),
                                                                                 static_cast<double *>(frames_out), n_frames, apply_window);
    } else {
        e = cudaFuncSetAttribute(k_c2r_frames<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_c2r_frames<float><<<static_cast<unsigned>(grid), 256, smem, stream>>>(p, static_cast<const float2 *>(stft),
                                                                                static_cast<float *>(frames_out), n_frames, apply_window);
    }
    return cudaGetLastError();
}

cudaError_t launch_ola_gather(bool f64, const void *frames, const void *window, void *out, long long n_clips, long long n_frames,
                              int n_fft, int hop, long long out_len, long long trim, cudaStream_t stream) {
    const long long total = n_clips * out_len;
    if (total <= 0) return cudaSuccess;
    const long long grid = (total + 255) / 256;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const long long full = (n_frames - 1) * hop + n_fft;
    if (f64)
        k_ola_gather<double><<<static_cast<unsigned>(grid), 256, 0, stream>>>(static_cast<const double *>(frames), static_cast<const double *>(window),
                                                                              static_cast<double *>(out), n_frames, n_fft, hop, out_len, trim, full, n_clips);
    else
        k_ola_gather<float><<<static_cast<unsigned>(grid), 256, 0, stream>>>(static_cast<const float *>(frames), static_cast<const float *>(window),
                                                                             static_cast<float *>(out), n_frames, n_fft, hop, out_len, trim, full, n_clips);
    return cudaGetLastError();
}

}  // namespace sgx
