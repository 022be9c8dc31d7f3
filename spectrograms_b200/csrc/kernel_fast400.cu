// kernel_fast400.cu -- "r2c_fused_n400": the Whisper-shaped family, n_fft = 400, hop = 160, f32 (BASELINE configs[1]
// and configs[3]). One CTA = one tile of 32 consecutive frames of one clip, 11 warps, lane = frame:
//
//   load    5360 samples (31*160 + 400) -> padded signal tile; zero fill outside the clip gives the centre padding
//           (src/spectrogram.rs:1309-1320) with no per-tap branch
//   pass 1  warps 0..9 : window multiply + 20-point real-pair DFT in registers   (fft400_core.cuh)
//   ----    one shared-memory exchange (Y[k1][n2], 11 x 20 complex per frame)
//   pass 2  warps 0..10: twiddle + 20-point DFT in registers -> |X|^2 straight into the power tile P[bin][frame]
//   epilogue mapping -> sqrt / dB -> (DCT-II + lifter) -> 128-byte row stores     (epilogue.cuh, lane = frame)
//
// Every shared-memory access in the two passes is conflict free by construction of the layouts (fft400_core.cuh);
// window and twiddles come from the constant bank (kernel parameter) with warp-uniform indices.
#include "epilogue.cuh"
#include "fft400_core.cuh"
#include "launch.hpp"

namespace sgx {

struct F400Params {
    KParams k;
    f400::Consts c;
};

namespace {

using namespace f400;

__global__ void __launch_bounds__(kThreads, 2) k_r2c_fused_n400(const __grid_constant__ F400Params P) {
    extern __shared__ __align__(16) float smem[];
    float *sig = smem;
    float *ybuf = sig + kSigWords;
    float *ptile = ybuf + kYWords;
    const KParams &p = P.k;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const long long f0 = p.frame_begin + static_cast<long long>(tile) * kFT;
    const long long rem = p.frame_begin + p.frames_todo - f0;
    const int nf = rem < kFT ? static_cast<int>(rem) : kFT;

    // ---- load the signal tile (8-byte units; p.buf_elems != 0 means the base is 8-byte aligned and the stride even)
    const float *x = static_cast<const float *>(p.samples) + static_cast<long long>(clip) * p.clip_stride;
    const long long s0 = f0 * kHop - p.pad;
    const long long n = p.n_samples;
    for (int j = tid; j < kTileSamples / 2; j += kThreads) {
        const long long s = s0 + 2 * j;
        float2 v;
        if (p.buf_elems && s >= 0 && s + 1 < n) {
            v = __ldg(reinterpret_cast<const float2 *>(x + s));
        } else {
            v.x = (s >= 0 && s < n) ? __ldg(x + s) : 0.f;
            v.y = (s + 1 >= 0 && s + 1 < n) ? __ldg(x + s + 1) : 0.f;
        }
        *reinterpret_cast<float2 *>(sig + 2 * j + 2 * (j / (kHop / 2))) = v;
    }
    __syncthreads();

    if (warp < 10) pass1_task(sig, ybuf, P.c, lane, warp);
    __syncthreads();

    pass2_task(ybuf, ptile, P.c, lane, warp);
    __syncthreads();

    // Y exchange buffer is dead now: scratch for the log-mel tile of the MFCC path (n_bins * 32 <= kYWords checked on host)
    epilogue_lane_frames<float>(p, ptile, ybuf, clip, f0, nf);
}

}  // namespace

size_t fast400_smem_bytes() { return sizeof(float) * (f400::kSigWords + f400::kYWords + f400::kPWords); }
int fast400_max_scratch_rows() { return f400::kYWords / 32; }

cudaError_t launch_fast400(const KParams &p, const float *window_f32, cudaStream_t stream) {
    static_assert(sizeof(F400Params) <= 4096, "kernel parameter block must fit the classic 4 KiB limit");
    F400Params P;
    P.k = p;
    P.k.FT = f400::kFT;
    P.k.tiles_per_clip = static_cast<int>((p.frames_todo + f400::kFT - 1) / f400::kFT);
    for (int i = 0; i < f400::kN; ++i) P.c.win[i] = window_f32[i];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int n2 = 0; n2 < 20; ++n2) {
            const long double a = -2.0L * pi * static_cast<long double>((n2 * k1) % 400) / 400.0L;
            const double s = (k1 == 0 || k1 == 10) ? 1.0 : 0.5;
            P.c.tw2[k1][n2] = make_float2(static_cast<float>(s * static_cast<double>(cosl(a))),
                                          static_cast<float>(s * static_cast<double>(sinl(a))));
        }
    const long long grid = static_cast<long long>(p.n_clips) * P.k.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const size_t smem = fast400_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_n400, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_n400<<<static_cast<unsigned>(grid), f400::kThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace sgx
