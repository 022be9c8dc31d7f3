// Microbenchmark: per-SM issue rate of scalar vs packed FP32 ops on B200 (compile: nvcc -gencode arch=compute_100a,code=sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE> __global__ void k(float *out, int iters, float s) {
    float a[16]; unsigned long long p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const unsigned long long ps = pk(s, s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {   // 16 independent scalar FFMA
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 1.0f);
            } else if (MODE == 1) {   // 16 independent scalar FADD
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
            } else if (MODE == 2) {   // 8 independent FFMA2 (= 16 lane-ops)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ps));
            } else {                  // 8 independent FADD2
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            }
        }
    }
    float acc = 0; for (int i = 0; i < 16; ++i) acc += a[i];
    unsigned long long q = 0; for (int i = 0; i < 8; ++i) q ^= p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)(q & 0xff);
}
template <int MODE> void run(const char *name, int warps_per_sm) {
    int iters = 2000; float *out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int threads = warps_per_sm * 32;
    k<MODE><<<148, threads>>>(out, 10, 1.0001f);
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * threads * iters * 8 * 16;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-8s warps/SM=%2d  %.3f ms  lane-ops/clk/SM = %.1f (at %d MHz)\n", name, warps_per_sm, ms, lane_ops / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<0>("FFMA", w); run<1>("FADD", w); run<2>("FFMA2", w); run<3>("FADD2", w); }
    return 0;
}
