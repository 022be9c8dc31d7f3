// kernel_n400_tc.cu -- "r2c_fused_n400_tc": n_fft = 400, hop = 160, f32 (BASELINE configs[1] / configs[3]) with the
// Blackwell tensor memory as the exchange medium and the filterbank projection on the tcgen05 tensor cores.
//
// Why: the shared-memory form of this family (kernel_fast400.cu) is bound by the L1 / shared-memory data pipe -- 103
// wavefronts per frame against 85 cycles per frame at the 60 %-of-HBM target -- and half of those wavefronts are the Y
// exchange between the two FFT passes and the power-tile round trip of the filterbank. Here both leave shared memory:
//
//   one persistent CTA per SM = 4 groups x 4 warps (+ 1 MMA-issue warp). Group q = the warps with warp_id % 4 == q, i.e.
//   the warps of SM sub-partition q, which own TMEM lanes 32q .. 32q+31. Each group walks its own tiles of 32 consecutive
//   frames (lane = frame = TMEM lane); the 4 tiles of the 4 groups form the 128 rows of one MMA.
//
//   pass 1     window + 20-point real-pair DFT in registers (fft400_core.cuh) -> Y[k1][n2] written to TMEM columns
//              0..399 of the thread's own lane with tcgen05.st                 (no shared-memory exchange buffer)
//   pass 2     tcgen05.ld of one Y row -> twiddle + 20-point DFT -> |X|^2 in registers; once every Y row of the group has
//              been read, the power values go back into the same TMEM columns, split for 3xTF32:
//              P_hi[bin] at column bin, P_lo[bin] at column 208 + bin           (no shared-memory power tile)
//   filterbank D[128 frames x 64 rows] (TMEM columns 416..479) = P[128 x 208] W^T, issued by ONE thread as tcgen05.mma
//              kind::tf32 with A = P from TMEM and B = the filterbank from shared memory, three MMAs per step
//              (P_hi W_hi + P_hi W_lo + P_lo W_hi), only over the 8-bin K steps where a 16-row block of the filterbank is
//              non-zero (mel / loghz rows are banded: ~34 steps instead of 208), 64 output rows per round
//   epilogue   tcgen05.ld of the thread's frame (16 rows per warp and round) -> sqrt / dB -> one 128-byte store per row
//
// The reference's arithmetic (SparseMatrix::multiply_vec, src/spectrogram.rs:102-117: acc += T(w) * x in ascending
// columns) is reproduced to 3xTF32 accuracy (relative 5e-7, measured by tools/ubench/tmem_probe.cu), inside the f32
// tolerance of 1e-5 relative L2 / 1e-3 dB; bins outside a block's K range contribute exact zeros.
#include "fast400_common.cuh"
#include "launch.hpp"
#include "tcgen05.cuh"

namespace sgx {
namespace {

using namespace f400;

constexpr int kGroupWarps = 4;                       // warps per group (one group per SM sub-partition)
constexpr int kComputeWarps = 4 * kGroupWarps;       // 16
constexpr int kTcThreads = kComputeWarps * 32;       // warp 0 also issues the MMAs (a 17th warp would cap the registers at 96)
constexpr int kGroupThreads = kGroupWarps * 32;
constexpr uint32_t kColPhi = 0, kColPlo = 208, kColD = 416, kTmemCols = 512;
constexpr int kRoundRows = 64;                       // output rows per MMA round (D columns)

struct TcSmem {
    float *sig;          // [4][kSigWords]
    float *win;          // [400]
    float *bsteps;       // per step: N rows x 8 bins as K-major core matrices, hi tile then lo tile (N = 16 or 64)
    int4 *steps;         // [n_steps] {a column, d column, byte offset of the step's B tiles, accumulate flag | N << 8}
    int *round_start;    // [n_rounds + 1]
    uint64_t *bars;      // pready, d[0], d[1], dfree
    uint32_t *tmem_ptr;
};

__device__ __forceinline__ TcSmem carve(unsigned char *base, int n_steps, int n_rounds, int b_floats) {
    TcSmem s;
    size_t o = 0;
    s.bsteps = reinterpret_cast<float *>(base + o);      o += sizeof(float) * static_cast<size_t>(b_floats);
    s.sig = reinterpret_cast<float *>(base + o);         o += sizeof(float) * 4 * kSigWords;
    s.win = reinterpret_cast<float *>(base + o);         o += sizeof(float) * kN;
    s.steps = reinterpret_cast<int4 *>(base + o);        o += sizeof(int4) * static_cast<size_t>(n_steps);
    s.bars = reinterpret_cast<uint64_t *>(base + o);     o += sizeof(uint64_t) * 4;
    s.round_start = reinterpret_cast<int *>(base + o);   o += sizeof(int) * static_cast<size_t>((n_rounds + 1 + 3) & ~3);
    s.tmem_ptr = reinterpret_cast<uint32_t *>(base + o);
    return s;
}

// ---- pass 1, one task = (frame = lane, column pair t): as f400::pass1_task, with Y going to the thread's TMEM lane.
// Y layout (columns): row 0 = (Y[0][n2], Y[10][n2]) pairs, rows 1..9 = Y[k1][n2] complex; column 40 * row + 2 * n2 (+1).
__device__ __forceinline__ void pass1_tc(const float *__restrict__ sig, const float *__restrict__ win, int f, int t, uint32_t ybase) {
    float2 v[20];
    const float *s = sig + kSigBlockStride * f + 2 * t;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const float2 x = *reinterpret_cast<const float2 *>(s + 20 * n1 + 2 * (n1 / 8));
        const float2 w = *reinterpret_cast<const float2 *>(win + 20 * n1 + 2 * t);
        v[n1] = cmul2(x, w);                          // sample * window[i] (src/spectrogram.rs:1319)
    }
    dft20(v);
    const uint32_t y = ybase + 4 * t;
    {
        const float2 z0 = v[reg_of_bin(0)], z10 = v[reg_of_bin(10)];
        tc::st4(y, __float_as_uint(z0.x), __float_as_uint(z10.x), __float_as_uint(z0.y), __float_as_uint(z10.y));
    }
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float2 A = v[reg_of_bin(k1)], B = v[reg_of_bin(20 - k1)];
        const float2 sa = cadd(A, make_float2(B.x, -B.y));                      // A + conj(B)
        const float2 sb = cadd(make_float2(A.y, -A.x), make_float2(B.y, B.x));  // (A - conj(B)) / i
        tc::st4(y + 40 * k1, __float_as_uint(sa.x), __float_as_uint(sa.y), __float_as_uint(sb.x), __float_as_uint(sb.y));
    }
}

// ---- pass 2, one task = (frame = lane, k1): Y row from TMEM -> twiddle -> DFT20 -> |X|^2 of the 20 outputs in pw[k2]
__device__ __forceinline__ void pass2_tc(uint32_t ybase, const float2 *__restrict__ tw2, int k1, float (&pw)[20]) {
    uint32_t q[40];
    const uint32_t row = ybase + ((k1 == 0 || k1 == 10) ? 0 : 40 * k1);
    tc::ld32(row, q);
    tc::ld8(row + 32, q + 32);
    tc::wait_ld();
    float2 v[20];
    if (k1 == 0) {
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            v[2 * j] = make_float2(__uint_as_float(q[4 * j]), 0.f);
            v[2 * j + 1] = make_float2(__uint_as_float(q[4 * j + 2]), 0.f);
        }
    } else if (k1 == 10) {
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            v[2 * j] = cmul2(bc2(__uint_as_float(q[4 * j + 1])), tw2[2 * j]);
            v[2 * j + 1] = cmul2(bc2(__uint_as_float(q[4 * j + 3])), tw2[2 * j + 1]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float2 w0 = tw2[2 * j], w1 = tw2[2 * j + 1];
            v[2 * j] = cfma2(bc2(__uint_as_float(q[4 * j + 1])), make_float2(-w0.y, w0.x), cmul2(bc2(__uint_as_float(q[4 * j])), w0));
            v[2 * j + 1] = cfma2(bc2(__uint_as_float(q[4 * j + 3])), make_float2(-w1.y, w1.x), cmul2(bc2(__uint_as_float(q[4 * j + 2])), w1));
        }
    }
    dft20(v);
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) {
        const float2 X = v[reg_of_bin(k2)];
        pw[k2] = norm_sqr(X);                         // norm_sqr (src/spectrogram.rs:1332-1334)
    }
}

// the bins a pass-2 task owns (f400::pass2_finish): P_hi -> column bin, P_lo -> column 208 + bin of the thread's lane
__device__ __forceinline__ void store_power_tc(uint32_t lane_base, int k1, const float (&pw)[20]) {
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) {
        int bin;
        bool ok = true;
        if (k2 < 10) bin = k1 + 20 * k2;
        else if (k2 == 10) { bin = 200 - k1; ok = k1 != 10; }
        else { bin = 400 - k1 - 20 * k2; ok = k1 != 0 && k1 != 10; }
        if (ok) {
            const float hi = tc::tf32_hi(pw[k2]);
            tc::st1(lane_base + kColPhi + bin, __float_as_uint(hi));
            tc::st1(lane_base + kColPlo + bin, __float_as_uint(pw[k2] - hi));
        }
    }
}

template <int AMP>
__device__ __forceinline__ void store_rows(const uint32_t (&v)[16], float *__restrict__ o, long long ors, int row0, int n_rows, float eps) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (row0 + i < n_rows) o[static_cast<long long>(row0 + i) * ors] = finish_value<AMP>(__uint_as_float(v[i]), eps);
}

__global__ void __launch_bounds__(kTcThreads, 1) k_r2c_fused_n400_tc(const __grid_constant__ F400Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const KParams &p = P.k;
    const int *blob = reinterpret_cast<const int *>(p.sched);
    const int n_steps = __ldg(blob), n_rounds = __ldg(blob + 1), b_floats = p.buf_elems;
    const TcSmem S = carve(smem_raw, n_steps, n_rounds, b_floats);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: tables -> shared memory, TMEM allocation, mbarriers
    {
        const int hdr = (2 + n_rounds + 1 + 3) & ~3;
        for (int i = tid; i <= n_rounds; i += kTcThreads) S.round_start[i] = __ldg(blob + 2 + i);
        const int4 *gsteps = reinterpret_cast<const int4 *>(blob + hdr);
        for (int i = tid; i < n_steps; i += kTcThreads) S.steps[i] = __ldg(gsteps + i);
        const float4 *gb = reinterpret_cast<const float4 *>(gsteps + n_steps);
        float4 *sb = reinterpret_cast<float4 *>(S.bsteps);
        for (int i = tid; i < b_floats / 4; i += kTcThreads) sb[i] = __ldg(gb + i);
        for (int i = tid; i < kN; i += kTcThreads) S.win[i] = P.c.win[i];
        if (warp == 0) tc::alloc(S.tmem_ptr, kTmemCols);
        if (tid == 32) {
            tc::mbar_init(&S.bars[0], kComputeWarps);      // pready: every compute warp has stored its power values
            tc::mbar_init(&S.bars[1], 1);                  // d[0], d[1]: the MMAs of a round have completed
            tc::mbar_init(&S.bars[2], 1);
            tc::mbar_init(&S.bars[3], kComputeWarps);      // dfree: every compute warp has read the round's D columns
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t tm = *S.tmem_ptr;

    const int tpc = p.tiles_per_clip;
    const long long total_tiles = static_cast<long long>(p.n_clips) * tpc;
    const long long n_super = (total_tiles + 3) / 4;

    {
        // ================================================================= compute warps
        const int q = warp & 3, wl = warp >> 2;            // group = SM sub-partition = TMEM lane quarter; warp within the group
        const int gt = wl * 32 + lane;                     // thread within the group
        const uint32_t lane_base = tm + (static_cast<uint32_t>(32 * q) << 16);
        float *sig = S.sig + q * kSigWords;
        const float *xbase = static_cast<const float *>(p.samples);
        const bool vec_ok = p.vec_ok != 0;
        const float eps = static_cast<float>(p.eps);
        const int amp = p.apply_db ? 2 : (p.amp == SGX_AMP_MAGNITUDE ? 1 : 0);

        if (wl == 0) {                                     // columns 400..415 are read by the MMA (zero weights) but never written
            const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            tc::st16(lane_base + 400, z);
            tc::wait_st();
        }
        long long g = 4 * static_cast<long long>(blockIdx.x) + q;      // this group's global tile index
        if (g < total_tiles) {
            const long long clip = g / tpc, tile = g - clip * tpc;
            load_tile(sig, xbase + clip * p.clip_stride, (p.frame_begin + tile * kFT) * kHop - p.pad, p.n_samples, vec_ok, gt,
                      kGroupThreads, kTileSamples / 2);
        }
        const bool issuer = warp == 0;                     // this warp's lane 0 issues the MMAs of the whole CTA
        const uint32_t b_base = tc::smem_addr(S.bsteps);
        auto issue_round = [&](int round, uint32_t Rr) {
            if (lane == 0) {
                const int i1 = S.round_start[round + 1];
                for (int i = S.round_start[round]; i < i1; ++i) {
                    const int4 st = S.steps[i];
                    const uint32_t N = static_cast<uint32_t>(st.w) >> 8;
                    const uint32_t idesc = tc::idesc_tf32(128, static_cast<int>(N));
                    const uint64_t bhi = tc::smem_desc_kmajor(b_base + static_cast<uint32_t>(st.z), 128, 256);
                    const uint64_t blo = tc::smem_desc_kmajor(b_base + static_cast<uint32_t>(st.z) + 32 * N, 128, 256);
                    const uint32_t d = tm + kColD + static_cast<uint32_t>(st.y);
                    tc::mma_tf32_ts(d, tm + kColPhi + static_cast<uint32_t>(st.x), bhi, idesc, static_cast<uint32_t>(st.w) & 1u);
                    tc::mma_tf32_ts(d, tm + kColPhi + static_cast<uint32_t>(st.x), blo, idesc, 1u);
                    tc::mma_tf32_ts(d, tm + kColPlo + static_cast<uint32_t>(st.x), bhi, idesc, 1u);
                }
                tc::commit(&S.bars[1 + (Rr & 1)]);
            }
            __syncwarp();
        };
        uint32_t R = 0, it = 0;
        for (long long s = blockIdx.x; s < n_super; s += gridDim.x, g += 4LL * gridDim.x, ++it) {
            const bool live = g < total_tiles;
            const long long clip = live ? g / tpc : 0, tile = live ? g - clip * tpc : 0;
            const long long f0 = p.frame_begin + tile * kFT;
            const long long rem = p.frame_begin + p.frames_todo - f0;
            const int nf = rem < kFT ? static_cast<int>(rem) : kFT;

            cp_async_commit_wait_all();
            tc::bar_sync(1 + q, kGroupThreads);            // the tile's samples have landed
            if (live) {
#pragma unroll 1
                for (int t = wl; t < 10; t += kGroupWarps) pass1_tc(sig, S.win, lane, t, lane_base);
            }
            tc::wait_st();
            tc::fence_before_sync();
            tc::bar_sync(1 + q, kGroupThreads);            // every Y column of the group is in TMEM; the samples are dead
            tc::fence_after_sync();
            {
                const long long gn = g + 4LL * gridDim.x;  // prefetch the group's next tile into its (only) signal buffer
                if (s + gridDim.x < n_super && gn < total_tiles) {
                    const long long cn = gn / tpc, tn = gn - cn * tpc;
                    load_tile(sig, xbase + cn * p.clip_stride, (p.frame_begin + tn * kFT) * kHop - p.pad, p.n_samples, vec_ok, gt,
                              kGroupThreads, kTileSamples / 2);
                }
            }
            float pw[3][20];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int k1 = wl + kGroupWarps * r;
                if (k1 <= 10 && live) pass2_tc(lane_base, P.c.tw2[k1], k1, pw[r]);
            }
            tc::fence_before_sync();
            tc::bar_sync(1 + q, kGroupThreads);            // every Y row has been read: the columns may be overwritten
            tc::fence_after_sync();
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int k1 = wl + kGroupWarps * r;
                if (k1 <= 10 && live) store_power_tc(lane_base, k1, pw[r]);
            }
            tc::wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&S.bars[0]);
            if (issuer) {                                  // all 128 rows of P are in TMEM: first round of the filterbank
                tc::mbar_wait(&S.bars[0], it & 1);
                tc::fence_after_sync();
                issue_round(0, R);
            }

            float *o = static_cast<float *>(p.out) + clip * p.out_clip_stride + (f0 - p.out_frame_origin) + lane;
            for (int round = 0; round < n_rounds; ++round, ++R) {
                tc::mbar_wait(&S.bars[1 + (R & 1)], (R >> 1) & 1);
                tc::fence_after_sync();
                uint32_t v[16];
                tc::ld16(lane_base + kColD + 16 * wl, v);
                tc::wait_ld();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&S.bars[3]);
                if (issuer && round + 1 < n_rounds) {      // next round once every warp has read this round's D columns
                    tc::mbar_wait(&S.bars[3], R & 1);
                    tc::fence_after_sync();
                    issue_round(round + 1, R + 1);
                }
                if (live && lane < nf) {
                    const int row0 = round * kRoundRows + 16 * wl;
                    if (amp == 2) store_rows<2>(v, o, p.out_row_stride, row0, p.n_bins, eps);
                    else if (amp == 1) store_rows<1>(v, o, p.out_row_stride, row0, p.n_bins, eps);
                    else store_rows<0>(v, o, p.out_row_stride, row0, p.n_bins, eps);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::dealloc(tm, kTmemCols);
}

}  // namespace

// dynamic shared memory of a launch with n_steps filterbank steps in n_rounds rounds; at least half an SM's worth so that
// one CTA (which owns all 512 TMEM columns) is resident per SM
size_t fast400_tc_smem_bytes(int n_steps, int n_rounds, size_t b_floats) {
    const size_t need = sizeof(float) * b_floats + sizeof(float) * (4 * f400::kSigWords + f400::kN) +
                        sizeof(int4) * static_cast<size_t>(n_steps) + sizeof(uint64_t) * 4 +
                        sizeof(int) * static_cast<size_t>((n_rounds + 1 + 3) & ~3) + 16;
    return std::max<size_t>(need, 120 * 1024);
}
bool fast400_tc_fits(int n_steps, int n_rounds, size_t b_floats) {
    return n_steps > 0 && fast400_tc_smem_bytes(n_steps, n_rounds, b_floats) <= 227 * 1024;
}

cudaError_t launch_fast400_tc(const KParams &p, const float *window_f32, int n_steps, int n_rounds, size_t b_floats, int sm_count,
                              cudaStream_t stream) {
    static_assert(sizeof(F400Params) <= 4096, "kernel parameter block must fit the classic 4 KiB limit");
    F400Params P;
    P.k = p;
    P.k.FT = f400::kFT;
    P.k.fd_FT = make_fastdiv(static_cast<unsigned>(f400::kFT));
    P.k.tiles_per_clip = static_cast<int>((p.frames_todo + f400::kFT - 1) / f400::kFT);
    P.k.buf_elems = static_cast<int>(b_floats);          // floats of filterbank tiles in the step blob
    fast400_fill_consts(P.c, window_f32);
    const long long total = static_cast<long long>(p.n_clips) * P.k.tiles_per_clip;
    if (total <= 0) return cudaSuccess;
    const long long n_super = (total + 3) / 4;
    const long long grid = std::min<long long>(n_super, sm_count);        // persistent: one CTA per SM
    const size_t smem = fast400_tc_smem_bytes(n_steps, n_rounds, b_floats);
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_n400_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_n400_tc<<<static_cast<unsigned>(grid), kTcThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace sgx
