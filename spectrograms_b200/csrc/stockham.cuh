// stockham.cuh -- the shared-memory Stockham autosort FFT of the generic kernel family, usable from any kernel that holds
// `nf` frames of L complex points each at bufA[f * FS + n] (forward transform, e^{-2 pi i jk/L}; an inverse is obtained
// by conjugating input and output). Radix 4/2/3/5 butterflies, one butterfly per thread, plus at most one "cofactor"
// stage that evaluates an r-point DFT directly, one output per thread -- this keeps the plan total for every length.
// Stage view: in [r][m][cur], out [m][r][cur]. Ends with a CTA barrier; returns the buffer that holds the result and
// leaves the other one free (*other).
#pragma once

#include "kparams.cuh"

namespace sgx {

template <typename T>
__device__ __forceinline__ typename Cplx<T>::type *stockham_stages(const KParams &p, typename Cplx<T>::type *bufA,
                                                                   typename Cplx<T>::type *bufB, int nf,
                                                                   typename Cplx<T>::type **other) {
    using C = typename Cplx<T>::type;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int L = p.L, FS = p.frame_stride;
    const C *tw = static_cast<const C *>(p.tw);
    C *in = bufA, *outb = bufB;
    int cur = 1;
    for (int s = 0; s < p.n_stages; ++s) {
        const int r = p.radix[s];
        const int B = L / r;           // butterflies per frame
        const int m = B / cur;                         // uniform: once per stage
        const bool fast = s < kFdStages;
        const FastDiv fdB = p.fd_stage_B[fast ? s : 0], fdc = p.fd_stage_cur[fast ? s : 0];
        if (r <= 5) {
            for (int idx = tid; idx < nf * B; idx += nthr) {
                const int f = fast ? fd_div(idx, fdB) : idx / B;
                const int b = idx - f * B;
                const int i = fast ? fd_div(b, fdc) : b / cur;
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
                const int f = fd_div(idx, p.fd_L);
                const int o = idx - f * L;        // o = (i*r + k)*cur + q
                const int ik = fd_div(o, fdc);            // stage 0 always has a precomputed divisor
                const int q = o - ik * cur;
                const int i = fd_div(ik, p.fd_r0);       // a cofactor stage is always stage 0
                const int k = ik - i * r;
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
    *other = outb;
    return in;
}

}  // namespace sgx
