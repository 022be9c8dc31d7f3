// mma_rate_probe.cu -- latency / throughput of small tcgen05.mma kind::tf32 instructions with A in TMEM (the shapes the
// n400 filterbank uses): cycles per MMA for chains accumulating into the same D columns and for independent D columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../spectrograms_b200/csrc -o mma_rate_probe mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace sgx;

template <int N>
__global__ void __launch_bounds__(128, 1) probe(long long *out, int n_mma, int d_stride, int a_stride, int a_tmem, int swz) {
    extern __shared__ __align__(1024) float sb[];          // zeros
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) sb[i] = 0.f;
    if (warp == 0) tc::alloc(&s_tmem, 512);
    if (tid == 0) tc::mbar_init(&bar, 1);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = s_tmem;
    const uint32_t z[16] = {0};
    for (int c = 0; c < 512; c += 16) tc::st16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, z);
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t b0 = tc::smem_addr(sb);
        const uint64_t adesc = tc::smem_desc_kmajor(b0 + 32768, 128, 256);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                // swz: SWIZZLE_128B K-major tiles (rows of 128 bytes = 4 K steps, 8-row groups 1024 bytes apart; layout type 2 in
                // bits 61..63) instead of the SWIZZLE_NONE core matrices -- timing only, the operands are zeros
                const uint64_t bd = swz ? (tc::smem_desc_kmajor(b0 + (i & 3) * 32, 16, 1024) | (2ull << 61))
                                        : tc::smem_desc_kmajor(b0 + (i & 7) * N * 32, 128, 256);
                const uint32_t d = tm + 256 + ((i * d_stride) & 255) % (256 - N + 1);
                if (a_tmem) tc::mma_tf32_ts(d, tm + ((i * a_stride) & 127), bd, idesc, 1u);
                else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
            const long long t1 = clock64();
            tc::commit(&bar);
            tc::mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            out[2 * rep] = t1 - t0;
            out[2 * rep + 1] = t2 - t0;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::dealloc(tm, 512);
}

template <int N>
void run(int n_mma, int d_stride, int a_stride, int a_tmem, const char *what, int swz = 0) {
    long long *d, h[6];
    cudaMalloc(&d, 48);
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe<N><<<1, 128, 100 * 1024>>>(d, n_mma, d_stride, a_stride, a_tmem, swz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("N=%3d %-34s %4d MMAs: issue %6lld cyc, complete %6lld cyc -> %.1f cyc / MMA (A from %s)\n", N, what, n_mma, h[4], h[5],
           static_cast<double>(h[5]) / n_mma, a_tmem ? "TMEM" : "smem");
    cudaFree(d);
}

int main() {
    run<16>(1, 0, 0, 1, "single");
    run<16>(100, 0, 8, 1, "same D (dependent chain)");
    run<16>(100, 16, 8, 1, "rotating D blocks (independent)");
    run<32>(100, 0, 8, 1, "same D");
    run<64>(100, 0, 8, 1, "same D");
    run<64>(100, 64, 8, 1, "rotating D");
    run<128>(100, 0, 8, 1, "same D");
    run<256>(100, 0, 8, 1, "same D");
    run<16>(100, 0, 8, 0, "same D");
    run<64>(100, 0, 8, 0, "same D");
    run<128>(100, 0, 8, 0, "same D");
    run<16>(100, 0, 8, 1, "same D, B SWIZZLE_128B", 1);
    run<48>(100, 0, 8, 1, "same D, B SWIZZLE_128B", 1);
    run<64>(100, 64, 8, 1, "rotating D, B SWIZZLE_128B", 1);
    run<128>(100, 0, 8, 1, "same D, B SWIZZLE_128B", 1);
    run<256>(100, 0, 8, 1, "same D, B SWIZZLE_128B", 1);
    return 0;
}
