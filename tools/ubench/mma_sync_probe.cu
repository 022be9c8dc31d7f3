// Microbenchmark: per-SM rate of the warp-level (legacy) mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 on B200, alone and
// with packed FP32 work on the FMA pipe next to it (does the tensor pipe run beside the FMA pipe from the same warps?).
// Question behind it: could pass 1 of r2c_fused_n400_tm move to the tensor pipe as 3xTF32 warp MMAs?
// compile: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_sync_probe mma_sync_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// MODE 0: 8 independent accumulators, MMAs only. MODE 1: the same plus 8 FFMA2 per 8 MMAs. MODE 2: FFMA2 only (8 per trip).
template <int MODE> __global__ void k(float *out, int iters, float s) {
    float d[8][4];
    unsigned a[4], b[2];
    unsigned long long p[8];
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(threadIdx.x * 0.001f + i), "f"(1.0f * i));
    }
    for (int j = 0; j < 4; ++j) a[j] = __float_as_uint(1.0f + 0.125f * ((threadIdx.x + j) & 7));
    b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
    unsigned long long ps; asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE != 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) mma_tf32(d[i], a, b);
            }
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ps));
            }
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc += d[i][j];
    unsigned long long q = 0; for (int i = 0; i < 8; ++i) q ^= p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)(q & 0xff);
}
template <int MODE> void run(const char *name, int warps_per_sm) {
    const int iters = 2000; float *out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = warps_per_sm * 32;
    k<MODE><<<148, threads>>>(out, 10, 1.0001f);
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;                 // SM cycles at the nominal clock
    const double trips = double(iters) * 4 * 8;                  // MMAs (and / or FFMA2s) per warp
    printf("%-28s warps/SM %2d  %.3f ms  cycles per warp-instruction per sub-partition: %.2f", name, warps_per_sm, ms,
           cycles / (trips * (warps_per_sm / 4.0)));
    if (MODE != 2) printf("  -> %.1f TF32 MMA m16n8k8 / clk / SM, %.1f dense TFLOP/s", trips * warps_per_sm / cycles,
                          2.0 * 16 * 8 * 8 * trips * warps_per_sm * 148 / (ms * 1e-3) / 1e12);
    printf("\n");
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("mma.sync tf32 only", w);
        run<2>("FFMA2 only", w);
        run<1>("mma.sync tf32 + FFMA2 (1:1)", w);
    }
    return 0;
}
