// kernel_mfcc_tc.cu -- "dct2_lifter_tc": mfcc_from_log_mel (src/mfcc.rs:224-273) as a tensor-core GEMM, f32.
//
// The DCT-II of a log-mel spectrogram is the dense contraction MFCC[frames x n_mfcc] = LogMel[frames x n_mels] . B[n_mels x
// n_mfcc] (dct_ii, src/mfcc.rs:278-292): 5 120 multiply-adds per frame at 128 mels / 40 coefficients, as much FP32 work as
// the FFT of the frame. Here it runs on the tcgen05 tensor cores with a 3xTF32 split (A_hi B_hi + A_hi B_lo + A_lo B_hi,
// relative error 5e-7: tools/ubench/tmem_probe.cu), FP32 accumulation in tensor memory, lifter and c0 drop fused on the way
// out (src/mfcc.rs:294-316).
//
//   one persistent CTA per SM, 8 warps; a tile = 128 consecutive frames of one clip = the 128 rows (TMEM lanes) of the MMA.
//   Warps q and q + 4 share SM sub-partition q (TMEM lanes 32q .. 32q+31, lane = frame) and split the mels between them.
//   load     per 64-mel half a thread reads 32 values of its frame straight from global memory (a warp instruction covers 32
//            consecutive frames of one mel row = 128 bytes); the loads of the NEXT half are issued before the current one is
//            converted, on two register sets, so the memory latency hides behind the conversion and the MMAs
//   split    each value into a TF32-exact hi part and the f32 remainder, written into tensor memory (tcgen05.st) as the A
//            operand: two A buffers (one per register set), so half h+1 is converted while the MMAs of half h run
//   mma      a ninth warp's lane 0 waits for a converted half (mbarrier) and issues tcgen05.mma.cta_group::1.kind::tf32, A from TMEM, B = the DCT basis from shared memory
//            (K-major SWIZZLE_NONE core matrices, hi and lo tiles per 8-mel K step, built on the host), M = 128,
//            N = n_mfcc rounded up to 16, two MMAs per K step: A_hi [B_hi | B_lo] (N doubled) and A_lo B_hi; completion through tcgen05.commit -> mbarrier
//   out      four more warps (one per sub-partition) wait for a tile's MMAs, tcgen05.ld the frame's coefficients (two D buffers:
//            tile t-1 is drained while tile t multiplies), add the two column halves, apply the lifter and store one 128-byte
//            run of frames per coefficient row; the conversion warps never wait for an MMA to complete
#include <type_traits>

#include "kparams.cuh"
#include "launch.hpp"
#include "tcgen05.cuh"

namespace sgx {
namespace {

constexpr int kTile = 128;                 // frames per tile = MMA rows
constexpr int kThreads = 256;              // 8 conversion warps: two per SM sub-partition
constexpr int kAllThreads = kThreads + 32 + 128; // + one warp that issues the MMAs + four warps (one per sub-partition) that drain D
constexpr int kHalf = 64;                  // mels per A buffer
constexpr int kPart = 32;                  // mels per thread and half
constexpr uint32_t kColA = 0;              // A buffers: [buf][hi 64 | lo 64]
constexpr uint32_t kColD = 256;            // D buffers: [buf][128]: columns [0, N) += A_hi B_hi + A_lo B_hi, columns [N, 2N) += A_hi B_lo
constexpr uint32_t kTmemCols = 512;

struct DctParams {
    const float *log_mel;      // [n_clips][n_mels][in_row_stride]
    float *out;                // [n_clips][rows][n_frames]
    const float *blob;         // per 8-mel K step: N x 8 hi tile then lo tile (K-major core matrices), then lifter[n_mfcc]
    long long n_frames, in_row_stride;
    long long out_clip_stride;  // floats between clips of the output
    int n_clips, n_mels, kp, n_mfcc, row0, N, tiles_per_clip;
    int out_row_base;           // output row of coefficient row0 (dense filterbank passes write row blocks of a wider output)
    int mode;                   // 0: coefficient * lifter (MFCC); 1: amplitude scaling of a filterbank row (amp / apply_db / eps)
    int amp, apply_db;
    float eps;
};

__global__ void __launch_bounds__(kAllThreads, 1) k_dct2_lifter_tc(const __grid_constant__ DctParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int n_steps = p.kp / 8;
    const int b_floats = n_steps * 2 * p.N * 8;
    float *sB = reinterpret_cast<float *>(smem_raw);                                  // basis tiles
    float *sLift = sB + b_floats;                                                     // [64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sLift + 64);                        // abar[2] (A buffer free), dbar[2] (D complete), rbar[2] (A buffer ready), fbar[2] (D buffer drained)
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hp = warp >> 2;                    // SM sub-partition / TMEM lane quarter; which 32 mels of a half
    const int frame_in_tile = 32 * q + lane;

    for (int i = tid; i < b_floats; i += kAllThreads) sB[i] = __ldg(p.blob + i);
    for (int i = tid; i < 64; i += kAllThreads) sLift[i] = i < p.n_mfcc ? __ldg(p.blob + b_floats + i) : 0.f;
    if (warp == 0) tc::alloc(tmem_ptr, kTmemCols);
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
        for (int i = 4; i < 6; ++i) tc::mbar_init(&bars[i], kThreads / 32);      // one arrival per conversion warp
        for (int i = 6; i < 8; ++i) tc::mbar_init(&bars[i], 4);                  // one arrival per drain warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    // The CTA allocates all 512 columns, so the allocation starts at lane 0, column 0. Using the constant (checked here) keeps
    // every MMA operand provably warp-uniform: the compiler then issues UTCHMMA straight from uniform registers instead of a
    // per-lane elect / broadcast loop around each instruction (~125 cycles per MMA on the issuing thread).
    if (*tmem_ptr != 0u) __trap();
    constexpr uint32_t tm = 0u;
    const uint32_t lane_base = tm + (static_cast<uint32_t>(32 * q) << 16);
    const uint32_t b_base = tc::smem_addr(sB);
    const uint32_t idesc = tc::idesc_tf32(128, p.N), idesc2 = tc::idesc_tf32(128, 2 * p.N);
    const int n_halves = (p.kp + kHalf - 1) / kHalf;

    const long long total = static_cast<long long>(p.n_clips) * p.tiles_per_clip;
    const long long my_tiles = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_jobs = my_tiles * n_halves;              // job g = (tile g / n_halves of this CTA, half g % n_halves)

    // Job g of this CTA = (its tile g / n_halves, half g % n_halves). The conversion warps walk the jobs in order and keep the job
    // coordinates as counters -- (tile, clip) of the tile advance by carry, no division per job -- once for the job being loaded
    // (one ahead) and implicitly (g) for the job being converted.
    const int step_clip = static_cast<int>(gridDim.x / p.tiles_per_clip), step_tile = static_cast<int>(gridDim.x % p.tiles_per_clip);
    int ld_h = 0, ld_clip = static_cast<int>(blockIdx.x / p.tiles_per_clip), ld_tile = static_cast<int>(blockIdx.x % p.tiles_per_clip);
    const unsigned row_bytes = static_cast<unsigned>(p.in_row_stride * static_cast<long long>(sizeof(float)));   // < 4 GB (checked by the host)
    // the 32 values of this thread's frame for the next job in order (zeros past the clip's last frame / the last mel)
    auto load_next = [&](float (&v)[kPart]) {
        const long long f = static_cast<long long>(ld_tile) * kTile + frame_in_tile;
        const int m0 = ld_h * kHalf + hp * kPart;
        const bool live = f < p.n_frames;
        const char *src = reinterpret_cast<const char *>(p.log_mel + (static_cast<long long>(ld_clip) * p.n_mels + m0) * p.in_row_stride + (live ? f : 0));
        const int nvalid = live ? p.n_mels - m0 : 0;
        // independent addresses (one 32 x 32 -> 64-bit multiply-add each: every load can issue at once); rows beyond the last get zeros
#pragma unroll
        for (int i = 0; i < kPart; ++i)
            v[i] = i < nvalid ? __ldg(reinterpret_cast<const float *>(src + static_cast<size_t>(static_cast<unsigned>(i) * static_cast<unsigned long long>(row_bytes)))) : 0.f;
        if (++ld_h == n_halves) {
            ld_h = 0;
            ld_clip += step_clip;
            ld_tile += step_tile;
            if (ld_tile >= p.tiles_per_clip) { ld_tile -= p.tiles_per_clip; ++ld_clip; }
        }
    };
    // coefficients of tile `it` of this CTA: D buffer it & 1 -> lifter -> rows of the output (drain warps)
    // (two instantiations, chosen once per kernel: the MFCC loop stays as tight as it was before the kernel learnt the scalings)
    auto drain = [&](auto mode_tag, long long it, uint32_t parity) {
        constexpr int MODE = decltype(mode_tag)::value;
        const int pb = static_cast<int>(it & 1);
        tc::mbar_wait(&bars[2 + pb], parity);
        tc::fence_after_sync();
        const long long t = blockIdx.x + it * gridDim.x;
        const long long clip = t / p.tiles_per_clip, tile = t - clip * p.tiles_per_clip;
        const long long f = tile * kTile + frame_in_tile;
        float *o = p.out + clip * p.out_clip_stride + static_cast<long long>(p.out_row_base) * p.n_frames + f;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
            uint32_t v[16];
            uint32_t u[16];
            tc::ld16(lane_base + kColD + 128u * pb + c0, v);
            tc::ld16(lane_base + kColD + 128u * pb + p.N + c0, u);
            tc::wait_ld();
            if (f < p.n_frames) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = c0 + i;
                    if (c >= p.row0 && c < p.n_mfcc) {
                        const float acc = __uint_as_float(v[i]) + __uint_as_float(u[i]);
                        o[static_cast<long long>(c - p.row0) * p.n_frames] = MODE == 0 ? acc * sLift[c] : amp_scale_fast(acc, p.amp, p.apply_db, p.eps, p.eps >= 1.17549435e-38f);
                    }
                }
            }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[6 + pb]);         // this sub-partition's quarter of the D buffer is free again
    };
    uint32_t a_uses[2] = {0, 0};          // MMA groups committed on each A buffer so far (mbarrier parity)
    // one job: job g + 1 is requested into `nxt`, job g is converted from `cur` into A buffer g & 1 and multiplied
    auto step = [&](float (&cur)[kPart], float (&nxt)[kPart], long long g) {
        const int b = static_cast<int>(g & 1);
        if (g + 1 < n_jobs) load_next(nxt);
        if (a_uses[b] > 0) {                                   // the MMAs that last read this A buffer are done
            tc::mbar_wait(&bars[b], (a_uses[b] - 1) & 1);
            tc::fence_after_sync();
        }
        const uint32_t acol = kColA + 128u * b + kPart * hp;
#pragma unroll
        for (int k = 0; k < kPart; k += 8) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float vh = tc::tf32_hi(cur[k + i]);
                hi[i] = __float_as_uint(vh);
                lo[i] = __float_as_uint(cur[k + i] - vh);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(lane_base + acol + k), "r"(hi[0]),
                         "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7])
                         : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(lane_base + acol + 64 + k), "r"(lo[0]),
                         "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                         : "memory");
        }
        tc::wait_st();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[4 + b]);          // this warp's 32 rows x 32 mels of the half are in TMEM
        a_uses[b] += 1;
    };

    if (__shfl_sync(0xffffffffu, warp, 0) == kThreads / 32) {
        // ---- the MMA warp: waits for a converted half, multiplies it, signals the A buffer free (and the tile complete). All 32
        //      lanes run the loop with identical (warp-uniform) operands and one elected lane issues: the operands then live
        //      in uniform registers and every tcgen05.mma is a single UTCHMMA (issued from divergent code, each one is wrapped
        //      in an elect / broadcast loop that costs the issuing thread ~125 cycles -- tools/ubench/mma_rate_probe.cu)
        {
            const uint64_t bdesc0 = tc::smem_desc_kmajor(b_base, 128, 256);
            const uint64_t step_units = static_cast<uint64_t>(4 * p.N), lo_units = static_cast<uint64_t>(2 * p.N);
            auto mma = [&](uint32_t d, uint32_t a, uint64_t bd, uint32_t id, uint32_t acc) {
                asm volatile(
                    "{\n"
                    ".reg .pred p, q;\n"
                    "elect.sync _|q, 0xffffffff;\n"
                    "setp.ne.b32 p, %4, 0;\n"
                    "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
                    "}\n" ::"r"(d),
                    "r"(a), "l"(bd), "r"(id), "r"(acc)
                    : "memory");
            };
            auto commit = [&](uint64_t *bar) {
                asm volatile(
                    "{\n"
                    ".reg .pred q;\n"
                    "elect.sync _|q, 0xffffffff;\n"
                    "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                    "}\n" ::"r"(tc::smem_addr(bar))
                    : "memory");
            };
            long long it = 0;
            int h = 0;
            for (long long g = 0; g < n_jobs; ++g, h = h + 1 == n_halves ? 0 : h + 1, it += h == 0 ? 1 : 0) {
                const int b = static_cast<int>(g & 1);
                tc::mbar_wait(&bars[4 + b], static_cast<uint32_t>((g >> 1) & 1));
                if (h == 0 && it >= 2) tc::mbar_wait(&bars[6 + (it & 1)], static_cast<uint32_t>(((it - 2) >> 1) & 1));   // tile it - 2 has left this D buffer
                tc::fence_after_sync();
                const int k0 = h * kHalf, kn = (p.kp - k0) < kHalf ? (p.kp - k0) : kHalf;      // mels of this half (multiple of 8)
                const uint32_t a0 = tm + kColA + 128u * b, dcol = tm + kColD + 128u * static_cast<uint32_t>(it & 1);
                // descriptors advance by constants: a K step's hi tile is 4 N, its lo tile 2 N sixteen-byte units further on
                uint64_t bhi = bdesc0 + static_cast<uint64_t>(k0 / 8) * step_units;
                // an MMA costs the tensor core ~125 cycles whatever N <= 128 (the A operand's 128 rows x 32 bytes arrive at 32 B/clk), so
                // A_hi meets [B_hi | B_lo] -- the lo tile directly follows the hi tile, one 2N-row operand -- in ONE instruction;
                // A_lo B_hi goes on top of the first N columns; the two column halves are added on the way out
                mma(dcol, a0, bhi, idesc2, h != 0 ? 1u : 0u);
                mma(dcol, a0 + 64, bhi, idesc, 1u);
                for (int k = 8; k < kn; k += 8) {
                    bhi += step_units;
                    mma(dcol, a0 + k, bhi, idesc2, 1u);
                    mma(dcol, a0 + 64 + k, bhi, idesc, 1u);
                }
                commit(&bars[b]);
                if (h == n_halves - 1) commit(&bars[2 + (it & 1)]);
            }
        }
    } else if (warp > kThreads / 32) {
        // ---- drain warps (one per sub-partition: warp_id % 4 selects the TMEM lanes): tile after tile as the MMAs complete
        if (p.mode == 0) for (long long it = 0; it < my_tiles; ++it) drain(std::integral_constant<int, 0>{}, it, static_cast<uint32_t>((it >> 1) & 1));
        else for (long long it = 0; it < my_tiles; ++it) drain(std::integral_constant<int, 1>{}, it, static_cast<uint32_t>((it >> 1) & 1));
    } else {
        float va[kPart], vb[kPart];
        if (n_jobs > 0) load_next(va);
        for (long long g = 0; g < n_jobs; g += 2) {
            step(va, vb, g);
            if (g + 1 < n_jobs) step(vb, va, g + 1);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::dealloc(tm, kTmemCols);
}

}  // namespace

bool mfcc_tc_supported(int n_mels, int n_mfcc) { return n_mels >= 8 && n_mels <= 128 && n_mfcc >= 1 && n_mfcc <= 64; }
int mfcc_tc_padded_mels(int n_mels) { return (n_mels + 7) & ~7; }
int mfcc_tc_padded_coeffs(int n_mfcc) { return (n_mfcc + 15) & ~15; }
// floats of the basis blob: per 8-mel K step an N x 8 hi tile and an N x 8 lo tile (float index (n / 8) * 64 + (k / 4) * 32 +
// (n % 8) * 4 + k % 4 inside a tile), then n_mfcc lifter weights
size_t mfcc_tc_blob_floats(int n_mels, int n_mfcc) {
    return static_cast<size_t>(mfcc_tc_padded_mels(n_mels) / 8) * 2 * mfcc_tc_padded_coeffs(n_mfcc) * 8 + static_cast<size_t>(n_mfcc);
}

namespace {
cudaError_t launch_gemm_tc(DctParams p, int sm_count, cudaStream_t stream) {
    p.kp = mfcc_tc_padded_mels(p.n_mels);
    p.N = mfcc_tc_padded_coeffs(p.n_mfcc);
    p.tiles_per_clip = static_cast<int>((p.n_frames + kTile - 1) / kTile);
    const long long total = static_cast<long long>(p.n_clips) * p.tiles_per_clip;
    if (total <= 0) return cudaSuccess;
    const size_t b_floats = static_cast<size_t>(p.kp / 8) * 2 * p.N * 8;
    size_t smem = sizeof(float) * (b_floats + 64) + sizeof(uint64_t) * 8 + 16;
    if (smem > 227 * 1024 || p.in_row_stride * static_cast<long long>(sizeof(float)) >= (1LL << 32)) return cudaErrorInvalidValue;
    smem = std::max<size_t>(smem, 120 * 1024);          // one CTA per SM: it owns all 512 TMEM columns
    cudaError_t e = cudaFuncSetAttribute(k_dct2_lifter_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long grid = std::min<long long>(total, sm_count);
    k_dct2_lifter_tc<<<static_cast<unsigned>(grid), kAllThreads, smem, stream>>>(p);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_mfcc_tc(const float *log_mel, long long in_row_stride, float *out, long long n_clips, int n_mels, long long n_frames, int n_mfcc,
                           int row0, const float *blob, int sm_count, cudaStream_t stream) {
    DctParams p{};
    p.log_mel = log_mel;
    p.in_row_stride = in_row_stride;
    p.out = out;
    p.blob = blob;
    p.n_frames = n_frames;
    p.n_clips = static_cast<int>(n_clips);
    p.n_mels = n_mels;
    p.n_mfcc = n_mfcc;
    p.row0 = row0;
    p.out_clip_stride = static_cast<long long>(n_mfcc - row0) * n_frames;
    p.out_row_base = 0;
    p.mode = 0;
    return launch_gemm_tc(p, sm_count, stream);
}

// One row block of a dense filterbank (ERB, src/erb.rs:374-402) applied to a linear power spectrogram as the same GEMM:
// power [n_clips][n_cols][n_frames] (rows in_row_stride apart) -> rows [row_base, row_base + n_rows) of out [n_clips][.][n_frames]
// (clips out_clip_stride floats apart), followed by the amplitude scaling. n_rows <= dense_tc_max_rows(n_cols); blob as for
// dct2_lifter_tc with the block's weights as the "basis" (no lifter section needed).
int dense_tc_max_rows(int n_cols) {
    const size_t steps = static_cast<size_t>(mfcc_tc_padded_mels(n_cols) / 8);
    const size_t budget = (size_t(200) * 1024) / (steps * 2 * 8 * sizeof(float));      // rows whose hi + lo tiles fit beside the barriers
    const int rows = static_cast<int>(std::min<size_t>(64, budget)) & ~15;
    return rows;
}
cudaError_t launch_dense_tc(const float *power, long long in_row_stride, float *out, long long out_clip_stride, long long n_clips, int n_cols,
                            long long n_frames, int row_base, int n_rows, const float *blob, int amp, int apply_db, float eps, int sm_count,
                            cudaStream_t stream) {
    DctParams p{};
    p.log_mel = power;
    p.in_row_stride = in_row_stride;
    p.out = out;
    p.blob = blob;
    p.n_frames = n_frames;
    p.n_clips = static_cast<int>(n_clips);
    p.n_mels = n_cols;
    p.n_mfcc = n_rows;
    p.row0 = 0;
    p.out_clip_stride = out_clip_stride;
    p.out_row_base = row_base;
    p.mode = 1;
    p.amp = amp;
    p.apply_db = apply_db;
    p.eps = eps;
    return launch_gemm_tc(p, sm_count, stream);
}

}  // namespace sgx
