// mixed_dft.cuh -- compile-time Cooley-Tukey DFT butterflies of any size R = 2^a 3^b 5^c held in registers, for the
// mixed-radix kernel family (kernel_mixed.cu). Everything is resolved at compile time: the factorisation, the register
// renaming and the internal twiddles W_R^m (constexpr cos / sin, folded into immediates or constant-bank operands).
// f32 arithmetic uses Blackwell's packed FP32x2 forms (one issue slot per complex add / scale).
#pragma once

#include <cuda_runtime.h>

namespace sgx {
namespace mx {

// ---- compile-time cos / sin of 2 pi num / den (argument reduced to [0, pi/4], Taylor to 1e-17)
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double taylor_sin(double x) {
    double term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x * x / ((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
constexpr double taylor_cos(double x) {
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x * x / ((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2 pi num / den), sin(2 pi num / den) for 0 <= num < den, exact at the multiples of 1/8 turn
constexpr double cos_turn(int num, int den) {
    num %= den;
    const int e = 8 * num;                     // eighths of a turn, scaled by den
    if (e == 0) return 1.0;
    if (e == 2 * den) return 0.0;
    if (e == 4 * den) return -1.0;
    if (e == 6 * den) return 0.0;
    if (e > 4 * den) return cos_turn(den - num, den);                 // cos(2 pi - x) = cos x
    if (e > 2 * den) return -cos_turn(den - 2 * num, 2 * den);         // cos(pi - x) = -cos x   (x = 2 pi (1/2 - num/den))
    if (e > den) return taylor_sin(2.0 * kPi * (0.25 - static_cast<double>(num) / den));   // cos x = sin(pi/2 - x)
    return taylor_cos(2.0 * kPi * static_cast<double>(num) / den);
}
constexpr double sin_turn(int num, int den) {
    num %= den;
    const int e = 8 * num;
    if (e == 0 || e == 4 * den) return 0.0;
    if (e == 2 * den) return 1.0;
    if (e == 6 * den) return -1.0;
    if (e > 4 * den) return -sin_turn(den - num, den);                 // sin(2 pi - x) = -sin x
    if (e > 2 * den) return sin_turn(den - 2 * num, 2 * den);          // sin(pi - x) = sin x
    if (e > den) return taylor_cos(2.0 * kPi * (0.25 - static_cast<double>(num) / den));   // sin x = cos(pi/2 - x)
    return taylor_sin(2.0 * kPi * static_cast<double>(num) / den);
}

// ---- complex value in registers
template <typename T> struct Cx { T x, y; };
__device__ __forceinline__ unsigned long long pk2(Cx<float> v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ Cx<float> upk2(unsigned long long v) { Cx<float> r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ Cx<float> operator+(Cx<float> a, Cx<float> b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> operator-(Cx<float> a, Cx<float> b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> mul2(Cx<float> a, Cx<float> b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> fma2(Cx<float> a, Cx<float> b, Cx<float> c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c))); return upk2(r); }
// a * b = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
__device__ __forceinline__ Cx<float> operator*(Cx<float> a, Cx<float> b) { return fma2(Cx<float>{a.y, a.y}, Cx<float>{-b.y, b.x}, mul2(Cx<float>{a.x, a.x}, b)); }
__device__ __forceinline__ Cx<float> scale(float s, Cx<float> a) { return mul2(Cx<float>{s, s}, a); }
__device__ __forceinline__ Cx<float> axpy(float s, Cx<float> a, Cx<float> b) { return fma2(Cx<float>{s, s}, a, b); }      // s a + b
__device__ __forceinline__ Cx<double> operator+(Cx<double> a, Cx<double> b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ Cx<double> operator-(Cx<double> a, Cx<double> b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ Cx<double> operator*(Cx<double> a, Cx<double> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ Cx<double> scale(double s, Cx<double> a) { return {s * a.x, s * a.y}; }
__device__ __forceinline__ Cx<double> axpy(double s, Cx<double> a, Cx<double> b) { return {fma(s, a.x, b.x), fma(s, a.y, b.y)}; }
template <typename T> __device__ __forceinline__ Cx<T> mul_mi(Cx<T> a) { return {a.y, -a.x}; }   // * (-i)
template <typename T> __device__ __forceinline__ Cx<T> mul_pi(Cx<T> a) { return {-a.y, a.x}; }   // * (+i)

// ---- prime-size kernels (forward, e^{-2 pi i jk/R}), in place
template <typename T> __device__ __forceinline__ void dft2(Cx<T> &a, Cx<T> &b) { const Cx<T> t = a - b; a = a + b; b = t; }
template <typename T> __device__ __forceinline__ void dft3(Cx<T> &a, Cx<T> &b, Cx<T> &c) {
    const T s3 = T(0.86602540378443864676372317075294);
    const Cx<T> t1 = b + c, d = mul_mi(scale(s3, b - c));      // -i s3 (b - c)
    const Cx<T> t2 = axpy(T(-0.5), t1, a);
    a = a + t1;
    b = t2 + d;
    c = t2 - d;
}
template <typename T> __device__ __forceinline__ void dft4(Cx<T> &a, Cx<T> &b, Cx<T> &c, Cx<T> &d) {
    const Cx<T> t0 = a + c, t1 = a - c, t2 = b + d, t3 = mul_mi(b - d);
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}
template <typename T> __device__ __forceinline__ void dft5(Cx<T> &a0, Cx<T> &a1, Cx<T> &a2, Cx<T> &a3, Cx<T> &a4) {
    const T c1 = T(0.30901699437494742410229341718282), c2 = T(-0.80901699437494742410229341718282);
    const T s1 = T(0.95105651629515357211643933337938), s2 = T(0.58778525229247312916870595463907);
    const Cx<T> p1 = a1 + a4, m1 = a1 - a4, p2 = a2 + a3, m2 = a2 - a3;
    const Cx<T> e1 = axpy(c2, p2, axpy(c1, p1, a0));
    const Cx<T> e2 = axpy(c1, p2, axpy(c2, p1, a0));
    const Cx<T> u1 = mul_mi(axpy(s2, m2, scale(s1, m1)));      // -i (s1 m1 + s2 m2)
    const Cx<T> u2 = mul_mi(axpy(-s1, m2, scale(s2, m1)));     // -i (s2 m1 - s1 m2)
    a0 = (a0 + p1) + p2;
    a1 = e1 + u1;
    a4 = e1 - u1;
    a2 = e2 + u2;
    a3 = e2 - u2;
}

constexpr int first_factor(int r) { return r % 4 == 0 ? 4 : (r % 2 == 0 ? 2 : (r % 3 == 0 ? 3 : (r % 5 == 0 ? 5 : r))); }
constexpr bool smooth235(int r) {
    while (r % 2 == 0) r /= 2;
    while (r % 3 == 0) r /= 3;
    while (r % 5 == 0) r /= 5;
    return r == 1;
}

// multiply by W_R^m = (cos, -sin)(2 pi m / R), m compile time; the trivial rotations cost nothing
template <typename T, int R, int M> __device__ __forceinline__ Cx<T> twiddle_const(Cx<T> a) {
    constexpr int m = M % R;
    if (m == 0) return a;
    if (4 * m == R) return mul_mi(a);
    if (2 * m == R) return Cx<T>{-a.x, -a.y};
    if (4 * m == 3 * R) return mul_pi(a);
    constexpr double c = cos_turn(m, R), s = sin_turn(m, R);
    const Cx<T> w = {static_cast<T>(c), static_cast<T>(-s)};
    return a * w;
}

// forward DFT of v[0..R-1], natural order in and out (decimation in time: n = A m + a, k = k2 + B k1, R = A B)
template <typename T, int R> struct Dft {
    static_assert(smooth235(R), "register butterflies exist for sizes 2^a 3^b 5^c");
    static constexpr int A = first_factor(R), B = R / A;
    template <int a, int k2> static __device__ __forceinline__ void tw_one(Cx<T> (&y)[A][B]) {
        y[a][k2] = twiddle_const<T, R, a * k2>(y[a][k2]);
        if constexpr (k2 + 1 < B) tw_one<a, k2 + 1>(y);
    }
    template <int a> static __device__ __forceinline__ void tw_row(Cx<T> (&y)[A][B]) {
        tw_one<a, 1>(y);
        if constexpr (a + 1 < A) tw_row<a + 1>(y);
    }
    static __device__ __forceinline__ void run(Cx<T> *v) {
        if constexpr (R == 1) {
        } else if constexpr (R == 2) {
            dft2(v[0], v[1]);
        } else if constexpr (R == 3) {
            dft3(v[0], v[1], v[2]);
        } else if constexpr (R == 4) {
            dft4(v[0], v[1], v[2], v[3]);
        } else if constexpr (R == 5) {
            dft5(v[0], v[1], v[2], v[3], v[4]);
        } else {
            Cx<T> y[A][B];
#pragma unroll
            for (int a = 0; a < A; ++a) {
#pragma unroll
                for (int m = 0; m < B; ++m) y[a][m] = v[A * m + a];
                Dft<T, B>::run(y[a]);
            }
            if constexpr (B > 1) tw_row<1>(y);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) {
                Cx<T> t[A];
#pragma unroll
                for (int a = 0; a < A; ++a) t[a] = y[a][k2];
                Dft<T, A>::run(t);
#pragma unroll
                for (int k1 = 0; k1 < A; ++k1) v[k2 + B * k1] = t[k1];
            }
        }
    }
};

}  // namespace mx
}  // namespace sgx
