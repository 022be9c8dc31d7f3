// fft400_core.cuh -- register-level core of the n_fft = 400 / hop = 160 f32 kernel family ("r2c_fused_n400").
//
// Real-input DFT of N = 400 = 20 x 20 in two register passes with ONE shared-memory exchange and no separate
// real-FFT post pass:
//
//   n = 20*n1 + n2,  k = k1 + 20*k2
//   pass 1 (over n1, for each n2):  Y[k1][n2] = sum_n1 x[20 n1 + n2] W20^(n1 k1)          real input -> only k1 = 0..10
//                                   two columns n2 = 2t, 2t+1 share one complex 20-point DFT (2-for-1 inside one frame,
//                                   so no cross-frame mixing of rounding noise)
//   pass 2 (over n2, for each k1):  X[k1 + 20 k2] = sum_n2 (Y[k1][n2] W400^(n2 k1)) W20^(n2 k2),   k1 = 0..10
//                                   outputs k2 >= 10 are the conjugates of bins 400 - k: every output is a wanted bin,
//                                   so bins 0..200 come out of 11 butterflies with nothing left to untangle.
//
// The 20-point DFT is a Good-Thomas 4 x 5 prime-factor butterfly: no internal twiddles, pure register renaming, and --
// written with Blackwell's packed FP32x2 instructions -- about 112 issue slots instead of 224. Everything here is __host__ __device__ so tests/test_fft400_core.py can run the exact task
// functions on the CPU (the build container has no GPU) against the oracle before a kernel ever launches.
#pragma once

#include <cuda_runtime.h>

#ifndef SGX_HD
#define SGX_HD __host__ __device__ __forceinline__
#endif

namespace sgx {
namespace f400 {

constexpr int kN = 400, kHop = 160, kBins = 201;
constexpr int kFT = 32;                 // frames per tile = lanes of a warp
constexpr int kWarps = 11;              // pass 1 uses 10 (t = 0..9), pass 2 uses 11 (k1 = 0..10)
constexpr int kThreads = kWarps * 32;

// shared-memory layouts (in 4-byte words)
//   signal tile : sample u of the tile (u = 0 <-> sample f0*160 - pad) lives at u + 2*(u/160). The 2-word pad per hop
//                 makes the frame stride 162 == 2 (mod 32): lanes = frames read 8-byte pairs from 16 distinct bank pairs.
//   Y exchange  : Y[f][row][n2] complex at 404 f + 40 row + 2 n2, row = k1 for k1 = 1..9; rows k1 = 0 and k1 = 10 are purely
//                 real and share row 0 as (Y[0][n2], Y[10][n2]). 404/4 odd -> 16-byte accesses with lanes = frames are
//                 conflict free both when pass 1 writes and when pass 2 reads.
//   power tile  : P[bin][frame_col(f)] at 32 bin + frame_col(f): frames fastest (conflict free for lanes = frames, and
//                 already the (rows, frames) orientation of the output), with the frame order permuted inside a row so
//                 that one 16-byte read returns frames j, j+8, j+16, j+24 (what one epilogue lane owns).
constexpr int kSigBlocks = 34;                         // ceil((31*160 + 400) / 160)
constexpr int kSigBlockStride = 162;
constexpr int kSigWords = kSigBlocks * kSigBlockStride;   // 5508
constexpr int kTileSamples = (kFT - 1) * kHop + kN;       // 5360
constexpr int kYFrameStride = 404;
constexpr int kYWords = kFT * kYFrameStride;              // 12928
constexpr int kPWords = kBins * kFT;                      // 6432

// per-plan constants, passed by value as a kernel parameter (constant bank; indices are warp-uniform)
struct Consts {
    float win[kN];          // window cast to f32 (make_window, src/spectrogram.rs:2232)
    float2 tw2[11][20];     // s(k1) * W400^(n2 k1), s = 1 for k1 in {0, 10}, 0.5 otherwise (folds the 2-for-1 halving)
};

SGX_HD int sig_word(int u) { return u + 2 * (u / kHop); }
SGX_HD int frame_col(int f) { return ((f & 7) << 2) | (f >> 3); }

// Packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): one instruction works on a (re, im) register pair, and
// the operand swizzle / per-half negate of the packed forms absorbs the +-i rotations of the butterflies, so a complex
// add, a complex-by-real scale and a "t1 + (-i) t3" each cost ONE issue slot. Host builds (the CPU emulation test) get
// the plain scalar definitions; both round every operation to nearest exactly once, so the results are identical.
#ifdef __CUDA_ARCH__
SGX_HD unsigned long long pk2(float2 v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
SGX_HD float2 upk2(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
SGX_HD float2 cadd(float2 a, float2 b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
SGX_HD float2 csub(float2 a, float2 b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
SGX_HD float2 cmul2(float2 a, float2 b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
SGX_HD float2 cfma2(float2 a, float2 b, float2 c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c))); return upk2(r); }
#else
SGX_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SGX_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SGX_HD float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
SGX_HD float2 cfma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
SGX_HD float2 bc2(float s) { return make_float2(s, s); }
SGX_HD float2 rot_mi(float2 v) { return make_float2(v.y, -v.x); }   // multiply by -i

// in-place 4-point DFT (forward): (a,b,c,d) <- (X0,X1,X2,X3)
SGX_HD void dft4(float2 &a, float2 &b, float2 &c, float2 &d) {
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = rot_mi(csub(b, d));
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, t3);
    d = csub(t1, t3);
}

// in-place 5-point DFT (forward): (a0..a4) <- (X0..X4)
SGX_HD void dft5(float2 &a0, float2 &a1, float2 &a2, float2 &a3, float2 &a4) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 p1 = cadd(a1, a4), m1 = csub(a1, a4), p2 = cadd(a2, a3), m2 = csub(a2, a3);
    const float2 e1 = cfma2(bc2(c2), p2, cfma2(bc2(c1), p1, a0));
    const float2 e2 = cfma2(bc2(c1), p2, cfma2(bc2(c2), p1, a0));
    const float2 u1 = rot_mi(cfma2(bc2(s2), m2, cmul2(bc2(s1), m1)));    // -i (s1 m1 + s2 m2)
    const float2 u2 = rot_mi(cfma2(bc2(-s1), m2, cmul2(bc2(s2), m1)));   // -i (s2 m1 - s1 m2)
    a0 = cadd(cadd(a0, p1), p2);
    a1 = cadd(e1, u1);
    a4 = csub(e1, u1);
    a2 = cadd(e2, u2);
    a3 = csub(e2, u2);
}

// Good-Thomas 20 = 4 x 5: input sample n sits in v[n]; afterwards output bin k sits in v[reg_of_bin(k)].
SGX_HD constexpr int reg_of_bin(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }

SGX_HD void dft20(float2 (&v)[20]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)   // 5-point DFTs over b on n = 5a + 4b (mod 20)
        dft5(v[(5 * a) % 20], v[(5 * a + 4) % 20], v[(5 * a + 8) % 20], v[(5 * a + 12) % 20], v[(5 * a + 16) % 20]);
#pragma unroll
    for (int d = 0; d < 5; ++d)   // 4-point DFTs over a on the same registers
        dft4(v[(4 * d) % 20], v[(5 + 4 * d) % 20], v[(10 + 4 * d) % 20], v[(15 + 4 * d) % 20]);
}

// ---- pass 1, one task = (frame f of the tile, column pair t): columns n2 = 2t and 2t+1
SGX_HD void pass1_task(const float *__restrict__ sig, float *__restrict__ ybuf, const float *__restrict__ win, int f, int t) {
    float2 v[20];
    const float *s = sig + kSigBlockStride * f + 2 * t;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const float2 x = *reinterpret_cast<const float2 *>(s + 20 * n1 + 2 * (n1 / 8));
        const float2 w = *reinterpret_cast<const float2 *>(win + 20 * n1 + 2 * t);
        v[n1] = cmul2(x, w);                          // sample * window[i] (src/spectrogram.rs:1319)
    }
    dft20(v);
    float *y = ybuf + kYFrameStride * f + 4 * t;      // complex index 2t -> word 4t
    // k1 = 0 and k1 = 10: both column spectra are real there; they share row 0 as (Y[0], Y[10]) pairs
    {
        const float2 z0 = v[reg_of_bin(0)], z10 = v[reg_of_bin(10)];
        *reinterpret_cast<float4 *>(y) = make_float4(z0.x, z10.x, z0.y, z10.y);
    }
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float2 A = v[reg_of_bin(k1)], B = v[reg_of_bin(20 - k1)];
        // 2*Ya = A + conj(B) ; 2*Yb = (A - conj(B)) / i   (the 1/2 lives in tw2)
        const float2 sa = cadd(A, make_float2(B.x, -B.y));                      // A + conj(B)
        const float2 sb = cadd(make_float2(A.y, -A.x), make_float2(B.y, B.x));  // (A - conj(B)) / i
        *reinterpret_cast<float4 *>(y + 40 * k1) = make_float4(sa.x, sa.y, sb.x, sb.y);
    }
}

// ---- pass 2, one task = (frame f, k1), in two halves so that a barrier can sit between "every Y value has been
// read" and "the power tile (which may alias the Y buffer) is written".
// tw2: this k1's 20 twiddles (constant bank or shared memory). k1 is warp-uniform in the kernel, so the three cases
// below are branches, not selects.
SGX_HD void pass2_load(const float *__restrict__ ybuf, const float2 *__restrict__ tw2, int f, int k1, float2 (&v)[20]) {
    if (k1 == 0) {                       // real row packed in the .x slots of row 0, twiddle = 1
        const float *y = ybuf + kYFrameStride * f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float4 q = *reinterpret_cast<const float4 *>(y + 4 * j);
            v[2 * j] = make_float2(q.x, 0.f);
            v[2 * j + 1] = make_float2(q.z, 0.f);
        }
    } else if (k1 == 10) {               // real row packed in the .y slots of row 0
        const float *y = ybuf + kYFrameStride * f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float4 q = *reinterpret_cast<const float4 *>(y + 4 * j);
            v[2 * j] = cmul2(bc2(q.y), tw2[2 * j]);
            v[2 * j + 1] = cmul2(bc2(q.w), tw2[2 * j + 1]);
        }
    } else {
        const float *y = ybuf + kYFrameStride * f + 40 * k1;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float4 q = *reinterpret_cast<const float4 *>(y + 4 * j);
            const float2 w0 = tw2[2 * j], w1 = tw2[2 * j + 1];
            // q * w = q.x * (w.x, w.y) + q.y * (-w.y, w.x)
            v[2 * j] = cfma2(bc2(q.y), make_float2(-w0.y, w0.x), cmul2(bc2(q.x), w0));
            v[2 * j + 1] = cfma2(bc2(q.w), make_float2(-w1.y, w1.x), cmul2(bc2(q.z), w1));
        }
    }
}

// norm_sqr (src/spectrogram.rs:1332-1334) as FMUL + FFMA: one rounding fewer than the reference's re*re + im*im (the contraction
// nvcc applies by itself in the generic / mixed / pow2 families), two FMA-pipe cycles per bin instead of three (packed multiply +
// add). SGX_NORM_UNFUSED keeps the two-rounding form for A/B runs.
SGX_HD float norm_sqr(float2 X) {
#ifdef SGX_NORM_UNFUSED
    const float2 sq = cmul2(X, X);
    return sq.x + sq.y;
#else
    return fmaf(X.x, X.x, X.y * X.y);
#endif
}

// writes |X[bin]|^2 for the bins this butterfly owns into P[bin][f]
SGX_HD void pass2_finish(float2 (&v)[20], float *__restrict__ ptile, int f, int k1) {
    dft20(v);
    float *p = ptile + frame_col(f);
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) {
        const float2 X = v[reg_of_bin(k2)];
        const float pw = norm_sqr(X);                 // norm_sqr = re*re + im*im (src/spectrogram.rs:1332-1334)
        if (k2 < 10) {
            p[kFT * (k1 + 20 * k2)] = pw;             // bin k1 + 20 k2
        } else if (k2 == 10) {
            if (k1 != 10) p[kFT * (200 - k1)] = pw;   // bin 400 - (k1 + 200); k1 = 10 would repeat bin 190
        } else {
            if (k1 != 0 && k1 != 10) p[kFT * (400 - k1 - 20 * k2)] = pw;   // k1 = 0 / 10: conjugate duplicates
        }
    }
}

SGX_HD void pass2_task(const float *__restrict__ ybuf, float *__restrict__ ptile, const Consts &c, int f, int k1) {
    float2 v[20];
    pass2_load(ybuf, c.tw2[k1], f, k1, v);
    pass2_finish(v, ptile, f, k1);
}

}  // namespace f400
}  // namespace sgx
