// tmem_probe.cu -- stand-alone check of the tcgen05 / TMEM building blocks the n400 kernel uses, run on the B200 box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
// (1) A operand written to TMEM from registers (tcgen05.st 32x32b: thread = lane = GEMM row), B operand from shared
//     memory through a K-major SWIZZLE_NONE descriptor, D accumulated in TMEM, read back with tcgen05.ld:
//     D[128 x N] = A[128 x K] B[K x N] in kind::tf32, single pass and 3xTF32 (hi*hi + hi*lo + lo*hi), against f64.
// (2) D written at an arbitrary column offset with N = 16 (the banded filterbank blocks).
// (3) throughput of tcgen05.ld / tcgen05.st per SM with 4, 8, 12, 16 warps (cycles per 32x32b.x32 instruction).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory descriptor: 8-row groups SBO bytes apart, the two 16-byte K chunks LBO bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;          // descriptor version 1 (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);      // a_major = b_major = K (0), no negate
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

#define TMEM_ST16(addr, v)                                                                                                     \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr), \
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),   \
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])                                          \
                 : "memory")
#define TMEM_LD16(addr, v)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"      \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
                 : "r"(addr)                                                                                                   \
                 : "memory")
#define TMEM_LD32(addr, v)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),        \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),       \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                     \
                 : "r"(addr)                                                                                                   \
                 : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int M = 128, K = 16;

// B in shared memory, per K step of 8: [n/8][k/4][n%8][k%4] floats -> LBO = 128 B, SBO = 256 B
__host__ __device__ inline int b_index(int N, int n, int k) { return (k / 8) * (N * 8) + (n / 8) * 64 + ((k % 8) / 4) * 32 + (n % 8) * 4 + (k % 4); }

template <int N, int PASSES>
__global__ void __launch_bounds__(128, 1) gemm_probe(const float *A, const float *B, float *D, int dcol) {
    __shared__ __align__(1024) float sb_hi[K * 64];
    __shared__ __align__(1024) float sb_lo[K * 64];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) mbar_init(&s_bar, 1);
    for (int i = tid; i < K * N; i += 128) {
        const int n = i / K, k = i % K;
        const float b = B[k * N + n];
        const float hi = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
        sb_hi[b_index(N, n, k)] = hi;
        sb_lo[b_index(N, n, k)] = b - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the MMA unit
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    const uint32_t lane_base = tm + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t hi[K], lo[K];
    for (int k = 0; k < K; ++k) {
        const float a = A[tid * K + k];
        const float h = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
        hi[k] = __float_as_uint(h);
        lo[k] = __float_as_uint(a - h);
    }
    TMEM_ST16(lane_base + 0, hi);
    TMEM_ST16(lane_base + 16, lo);
    tmem_wait_st();
    fence_before();
    __syncthreads();
    if (tid == 0) {
        fence_after();
        const uint32_t idesc = make_idesc(M, N);
        const uint32_t d = tm + dcol;
        for (int j = 0; j < K / 8; ++j) {
            const uint64_t bh = make_desc(smem_u32(sb_hi) + j * N * 32, 128, 256);
            const uint64_t bl = make_desc(smem_u32(sb_lo) + j * N * 32, 128, 256);
            umma_ts(d, tm + 8 * j, bh, idesc, j > 0);
            if (PASSES == 3) {
                umma_ts(d, tm + 8 * j, bl, idesc, 1);
                umma_ts(d, tm + 16 + 8 * j, bh, idesc, 1);
            }
        }
        umma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    fence_after();
    uint32_t v[16];
    for (int c = 0; c < N; c += 16) {
        TMEM_LD16(lane_base + dcol + c, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) D[tid * N + c + i] = __uint_as_float(v[i]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

// throughput: every warp hammers its own quadrant; report cycles per instruction per SM
template <bool LOAD>
__global__ void __launch_bounds__(1024, 1) rate_probe(long long *cycles, int iters, float *sink) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t lane_base = s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = tid + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = lane_base + ((it * 32) & 255);
        if (LOAD) {
            TMEM_LD32(a, v);
            tmem_wait_ld();
            acc += v[it & 31];
        } else {
            TMEM_ST16(a, v);
            TMEM_ST16(a + 16, (v + 16));
            tmem_wait_st();
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    if (acc == 0x12345678u) *sink = 1.f;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(s_tmem, 512);
}

template <int N, int PASSES>
double run_gemm(int dcol) {
    std::vector<float> A(M * K), B(K * N), D(M * N);
    srand(1234 + N + PASSES);
    for (auto &a : A) a = static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f;
    for (auto &b : B) b = static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    gemm_probe<N, PASSES><<<1, 128>>>(dA, dB, dD, dcol);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0, worst = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += static_cast<double>(A[m * K + k]) * B[k * N + n];
            const double e = D[m * N + n] - r;
            num += e * e; den += r * r; worst = fmax(worst, fabs(e));
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    printf("gemm M=128 N=%d K=%d passes=%d dcol=%d : rel-L2 %.3e  max-abs %.3e\n", N, K, PASSES, dcol, sqrt(num / den), worst);
    return sqrt(num / den);
}

int main() {
    run_gemm<32, 1>(64);
    run_gemm<32, 3>(64);
    run_gemm<16, 3>(72);
    run_gemm<16, 3>(100);
    run_gemm<64, 3>(128);
    long long *dc; float *sink;
    CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&sink, 4));
    for (int warps : {4, 8, 12, 16, 24}) {
        long long c;
        rate_probe<true><<<1, warps * 32>>>(dc, 2000, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        printf("tcgen05.ld 32x32b.x32 + wait: %2d warps: %.1f cycles / instruction / warp, %.2f cycles per instruction per SM (4 KB each)\n", warps,
               static_cast<double>(c) / 2000, static_cast<double>(c) / 2000 / warps);
        rate_probe<false><<<1, warps * 32>>>(dc, 2000, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        printf("tcgen05.st 2 x 32x32b.x16 + wait: %2d warps: %.1f cycles / 4 KB / warp, %.2f cycles per 4 KB per SM\n", warps,
               static_cast<double>(c) / 2000, static_cast<double>(c) / 2000 / warps);
    }
    return 0;
}
