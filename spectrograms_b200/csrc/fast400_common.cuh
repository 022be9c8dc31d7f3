// fast400_common.cuh -- pieces shared by the two kernels of the n_fft = 400 / hop = 160 f32 family
// (kernel_fast400.cu: shared-memory exchange + CUDA-core filterbank; kernel_n400_tc.cu: TMEM exchange + tcgen05 filterbank).
#pragma once

#include <cmath>

#include "fft400_core.cuh"
#include "kparams.cuh"

namespace sgx {

struct F400Params {
    KParams k;
    f400::Consts c;
};

// per-plan constants of both kernels: the f32 window and the pass-2 twiddles s(k1) W400^(n2 k1) (extended precision -> f32)
inline void fast400_fill_consts(f400::Consts &c, const float *window_f32) {
    for (int i = 0; i < f400::kN; ++i) c.win[i] = window_f32[i];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int n2 = 0; n2 < 20; ++n2) {
            const long double a = -2.0L * pi * static_cast<long double>((n2 * k1) % 400) / 400.0L;
            const double s = (k1 == 0 || k1 == 10) ? 1.0 : 0.5;
            c.tw2[k1][n2] = make_float2(static_cast<float>(s * static_cast<double>(cosl(a))),
                                        static_cast<float>(s * static_cast<double>(sinl(a))));
        }
}

namespace f400 {

__device__ __forceinline__ void cp_async8(float *dst_smem, const float *src, int src_bytes) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst_smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// stage one tile's samples [s0, s0 + 5360) of a clip into a padded signal buffer
// float2 units j = first, first + step, ... < last of the tile are handled by the calling thread
__device__ __forceinline__ void load_tile(float *sig, const float *x, long long s0, long long n, bool vec_ok, int first, int step,
                                          int last) {
    if (vec_ok && s0 >= 0 && s0 + kTileSamples <= n) {
        // interior tile (all but the first / last tile of a clip): no bounds logic at all
        const float *src = x + s0;
#pragma unroll 4
        for (int j = first; j < last; j += step) cp_async8(sig + 2 * j + 2 * (j / (kHop / 2)), src + 2 * j, 8);
    } else if (vec_ok) {
        for (int j = first; j < last; j += step) {
            const long long s = s0 + 2 * j;
            const long long avail = n - s;                // samples available from s on
            const int bytes = (s < 0 || avail <= 0) ? 0 : (avail >= 2 ? 8 : 4);
            cp_async8(sig + 2 * j + 2 * (j / (kHop / 2)), bytes ? x + s : x, bytes);
        }
    } else {
        for (int j = first; j < last; j += step) {
            const long long s = s0 + 2 * j;
            float2 v;
            v.x = (s >= 0 && s < n) ? __ldg(x + s) : 0.f;
            v.y = (s + 1 >= 0 && s + 1 < n) ? __ldg(x + s + 1) : 0.f;
            *reinterpret_cast<float2 *>(sig + 2 * j + 2 * (j / (kHop / 2))) = v;
        }
    }
}

// lg2.approx.ftz: the argument is clamped to eps > 0 first, so the denormal fix-up of __log2f is dead weight
__device__ __forceinline__ float fast_lg2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// AMP: 0 power, 1 magnitude, 2 dB (AmpScale::apply_from_power + apply_db_in_place, src/spectrogram.rs:1986-2037, :2068-2080)
template <int AMP>
__device__ __forceinline__ float finish_value(float acc, float eps) {
    if (AMP == 1) acc = sqrtf(acc);
    if (AMP == 2) acc = 3.01029995663981195f * fast_lg2(fmaxf(acc, eps));
    return acc;
}

}  // namespace f400
}  // namespace sgx
