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
    const C *tw = static_cast<const C *>(p.tw);

    // ---- 1. gather / pad / window / pack
    for (int idx = tid; idx < nf * L; idx += nthr) {
        const int f = idx / L;
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

    // ---- 2. Stockham stages: in viewed [r][m][cur], out viewed [m][r][cur]
    C *in = bufA, *outb = bufB;
    int cur = 1;
    for (int s = 0; s < p.n_stages; ++s) {
        const int r = p.radix[s];
        const int B = L / r;           // butterflies per frame
        const int m = B / cur;
        if (r <= 5) {
            for (int idx = tid; idx < nf * B; idx += nthr) {
                const int f = idx / B;
                const int b = idx - f * B;
                const int i = b / cur;
                const int q = b - i * cur;
                const C *src = in + f * FS + b;
                C *dst = outb + f * FS + i * r * cur + q;
                const int tstep = q * m;   // twiddle index step: W_{cur*r}^{j q} = W_L^{j q m}
                if (r == 2) {
                    const C a = src[0];
                    const C bb = cmul(src[B], tw[tstep]);
                    dst[0] = cadd(a, bb);
                    dst[cur] = csub(a, bb);
                } else if (r == 4) {
                    const C a = src[0];
                    const C bb = cmul(src[B], tw[tstep]);
                    const C c = cmul(src[2 * B], tw[2 * tstep]);
                    const C d = cmul(src[3 * B], tw[3 * tstep]);
                    const C t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(bb, d), t3 = csub(bb, d);
                    dst[0] = cadd(t0, t2);
                    dst[cur] = mk<T>(t1.x + t3.y, t1.y - t3.x);
                    dst[2 * cur] = csub(t0, t2);
                    dst[3 * cur] = mk<T>(t1.x - t3.y, t1.y + t3.x);
                } else if (r == 3) {
                    const T s3 = T(0.86602540378443864676372317075294);
                    const C a = src[0];
                    const C bb = cmul(src[B], tw[tstep]);
                    const C c = cmul(src[2 * B], tw[2 * tstep]);
                    const C t1 = cadd(bb, c);
                    const C t2 = mk<T>(a.x - T(0.5) * t1.x, a.y - T(0.5) * t1.y);
                    const C d = csub(bb, c);
                    dst[0] = cadd(a, t1);
                    dst[cur] = mk<T>(t2.x + s3 * d.y, t2.y - s3 * d.x);
                    dst[2 * cur] = mk<T>(t2.x - s3 * d.y, t2.y + s3 * d.x);
                } else {   // r == 5
                    const T c1 = T(0.30901699437494742410229341718282), c2 = T(-0.80901699437494742410229341718282);
                    const T s1 = T(0.95105651629515357211643933337938), s2 = T(0.58778525229247312916870595463907);
                    const C a0 = src[0];
                    const C a1 = cmul(src[B], tw[tstep]);
                    const C a2 = cmul(src[2 * B], tw[2 * tstep]);
                    const C a3 = cmul(src[3 * B], tw[3 * tstep]);
                    const C a4 = cmul(src[4 * B], tw[4 * tstep]);
                    const C p1 = cadd(a1, a4), m1 = csub(a1, a4), p2 = cadd(a2, a3), m2 = csub(a2, a3);
                    dst[0] = mk<T>(a0.x + p1.x + p2.x, a0.y + p1.y + p2.y);
                    const C e1 = mk<T>(a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y);
                    const C e2 = mk<T>(a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y);
                    const C u1 = mk<T>(s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y);
                    const C u2 = mk<T>(s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y);
                    // X1 = e1 - i u1 ; X4 = e1 + i u1 ; X2 = e2 - i u2 ; X3 = e2 + i u2
                    dst[cur] = mk<T>(e1.x + u1.y, e1.y - u1.x);
                    dst[4 * cur] = mk<T>(e1.x - u1.y, e1.y + u1.x);
                    dst[2 * cur] = mk<T>(e2.x + u2.y, e2.y - u2.x);
                    dst[3 * cur] = mk<T>(e2.x - u2.y, e2.y + u2.x);
                }
            }
        } else {
            // cofactor stage: out[i][k][q] = sum_j in[j][i][q] * W_L^{ j * (q*m + k*(L/r)) }, one output per thread
            for (int idx = tid; idx < nf * L; idx += nthr) {
                const int f = idx / L;
                const int o = idx - f * L;        // o = (i*r + k)*cur + q
                const int q = o % cur;
                const int ik = o / cur;
                const int k = ik % r;
                const int i = ik / r;
                const int b = i * cur + q;
                const C *src = in + f * FS + b;
                const int step = static_cast<int>((static_cast<long long>(q) * m + static_cast<long long>(k) * B) % L);
                int e = 0;
                C acc = mk<T>(T(0), T(0));
                for (int j = 0; j < r; ++j) {
                    const C v = src[j * B];
                    const C w = tw[e];
                    acc.x += v.x * w.x - v.y * w.y;
                    acc.y += v.x * w.y + v.y * w.x;
                    e += step;
                    if (e >= L) e -= L;
                }
                outb[f * FS + o] = acc;
            }
        }
        __syncthreads();
        C *t = in; in = outb; outb = t;
        cur *= r;
    }
    // `in` now holds Z (packed spectrum, even n_fft) or the full spectrum (odd n_fft); `outb` is free

    // ---- 3. post pass
    const C *Z = in;
    if (p.output == SGX_OUT_COMPLEX_STFT) {
        C *S = outb;
        const C *post = static_cast<const C *>(p.post);
        for (int idx = tid; idx < nf * p.out_len; idx += nthr) {
            const int f = idx / p.out_len;
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
            const int f = idx / p.out_len;
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
