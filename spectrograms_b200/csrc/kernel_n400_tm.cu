// kernel_n400_tm.cu -- "r2c_fused_n400_tm": n_fft = 400, hop = 160, f32 (BASELINE configs[1]) with the Blackwell tensor
// memory (TMEM) as the exchange medium between the two FFT passes and warp-specialised filterbank warps.
//
// Why: the shared-memory form (kernel_fast400.cu) is bound by the L1 / shared-memory data pipe: 3480 wavefronts per 32-frame
// tile, of which 840 are the Y exchange between the passes (profiles/r1_n400_phase_ablation.md). TMEM has its own data
// path (tcgen05.st 794 B/clk, tcgen05.ld 57 B/clk per SM, tools/ubench/tmem_probe.cu), so the exchange leaves the L1 pipe,
// the 52 KB exchange buffer leaves shared memory, and four tiles instead of two are in flight per SM.
//
//   one persistent CTA per SM = 4 groups; group q = the warps with warp_id % 4 == q (SM sub-partition q, TMEM lanes
//   32q .. 32q+31). A group walks its own tiles of 32 consecutive frames (lane = frame = TMEM lane) and never synchronises
//   with the other groups: the four sub-partitions run their FFT-bound and load/store-bound phases out of step.
//
//   FFT warps (4 per group)
//     pass 1   window + 20-point real-pair DFT in registers (fft400_core.cuh) -> Y[k1][n2] into TMEM columns 0..399 of the
//              thread's own lane (tcgen05.st, SASS STTM)
//     pass 2   tcgen05.ld (LDTM) of one Y row -> twiddle + 20-point DFT -> |X|^2 into the group's power tile P[bin][frame]
//              in shared memory
//   filterbank warps (EW per group; EW = 0: the FFT warps do this themselves after a group barrier)
//     sparse mel / loghz rows from the power tile (the quad schedule of kernel_fast400.cu) -> sqrt / dB -> row stores.
//     They hand the tile back through a pair of named barriers (full / free), so the FFT warps start pass 1 of the next
//     tile while the rows of the previous one are still being written: the FP32 pipe of the sub-partition does not idle
//     during the load/store-bound epilogue.
//
// Arithmetic is the arithmetic of kernel_fast400.cu (same task functions, same epilogue); only where Y travels differs.
#include "fast400_common.cuh"
#include "launch.hpp"
#include "tcgen05.cuh"

namespace sgx {
namespace {

using namespace f400;

constexpr int kGroups = 4;
constexpr int kFftWarpsPerGroup = 4;
constexpr int kFftWarps = kGroups * kFftWarpsPerGroup;      // 16
constexpr int kFftGroupThreads = kFftWarpsPerGroup * 32;    // 128
constexpr uint32_t kTmemCols = 512;                         // Y needs 400 columns; allocations are powers of two

struct TmSmem {
    float *sig;        // [4][kSigWords]   one signal tile per group
    float *ptile;      // [4][kPWords]     one power tile per group
    float *win;        // [400]
    int4 *quads;       // [4 * n_quads]    {byte offset of P[c0], cnt, weights address, row}
    float *w;          // padded weights
    uint32_t *tmem_ptr;
};

__host__ __device__ inline size_t tm_smem_bytes(int n_quads, int padded_weights) {
    return sizeof(float) * (kGroups * (kSigWords + kPWords) + kN) + sizeof(int4) * 4 * static_cast<size_t>(n_quads) +
           sizeof(float) * static_cast<size_t>(padded_weights + 8) + 16;
}

__device__ __forceinline__ TmSmem carve(unsigned char *base, int n_quads, int padded_weights) {
    TmSmem s;
    size_t o = 0;
    s.sig = reinterpret_cast<float *>(base + o);        o += sizeof(float) * kGroups * kSigWords;
    s.ptile = reinterpret_cast<float *>(base + o);      o += sizeof(float) * kGroups * kPWords;
    s.quads = reinterpret_cast<int4 *>(base + o);       o += sizeof(int4) * 4 * static_cast<size_t>(n_quads);
    s.win = reinterpret_cast<float *>(base + o);        o += sizeof(float) * kN;
    s.w = reinterpret_cast<float *>(base + o);          o += sizeof(float) * static_cast<size_t>(padded_weights + 8);
    s.tmem_ptr = reinterpret_cast<uint32_t *>(base + o);
    return s;
}

__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---- pass 1, one task = (frame = lane, column pair t): f400::pass1_task with Y going to the thread's TMEM lane.
// Y layout (columns): row 0 = (Y[0][n2], Y[10][n2]) pairs, rows 1..9 = Y[k1][n2] complex; column 40 * row + 2 * n2 (+1).
__device__ __forceinline__ void pass1_tm(const float *__restrict__ sig, const float *__restrict__ win, int f, int t, uint32_t ybase) {
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

// ---- pass 2, one task = (frame = lane, k1): Y row from TMEM -> twiddle -> DFT20 -> |X|^2 into P[bin][frame]
__device__ __forceinline__ void pass2_tm(uint32_t ybase, const float2 *__restrict__ tw2, float *__restrict__ ptile, int f, int k1) {
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
    pass2_finish(v, ptile, f, k1);
}

__device__ __forceinline__ void rows_epilogue(const KParams &p, const float *ptile, const int4 *quads, int q0, int nq, int qstep, float *ocf,
                                              int nf, int lane) {
    if (p.apply_db) sparse_quads_epilogue<2, false>(p, ptile, quads, q0, nq, qstep, ocf, nullptr, nf, lane);
    else if (p.amp == SGX_AMP_MAGNITUDE) sparse_quads_epilogue<1, false>(p, ptile, quads, q0, nq, qstep, ocf, nullptr, nf, lane);
    else sparse_quads_epilogue<0, false>(p, ptile, quads, q0, nq, qstep, ocf, nullptr, nf, lane);
}

template <int EW>
__global__ void __launch_bounds__(32 * (kFftWarps + kGroups * EW), 1) k_r2c_fused_n400_tm(const __grid_constant__ F400Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kThreadsAll = 32 * (kFftWarps + kGroups * EW);
    constexpr int kHandshake = kFftGroupThreads + 32 * EW;      // threads on the full / free barriers of a group
    const KParams &p = P.k;
    const int *blob = reinterpret_cast<const int *>(p.sched);   // int n_quads; int qrange[2]; int maxcnt[n_quads]; pad; int4 quads[4 n]
    const int nq = __ldg(blob);
    const int padded_weights = p.buf_elems;
    const TmSmem S = carve(smem_raw, nq, padded_weights);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: tables -> shared memory, TMEM allocation
    {
        const int hdr = (1 + 2 + nq + 3) & ~3;
        const int4 *gq = reinterpret_cast<const int4 *>(blob + hdr);
        const float *val = static_cast<const float *>(p.val);
        const unsigned wbase = smem_u32(S.w);
        for (int i = tid; i < kN; i += kThreadsAll) S.win[i] = P.c.win[i];
        for (int i = tid; i < 4 * nq; i += kThreadsAll) {
            const int4 e = __ldg(gq + i);
            S.quads[i] = make_int4(e.x * (kFT * 4), e.y, static_cast<int>(wbase + 4u * e.z), e.w);
            if (e.w >= 0) {
                const int e0 = __ldg(p.row_ptr + e.w);
                for (int k = 0; k < ((e.y + 3) & ~3); ++k) S.w[e.z + k] = k < e.y ? __ldg(val + e0 + k) : 0.f;
            }
        }
        if (warp == 0) tc::alloc(S.tmem_ptr, kTmemCols);
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t tm = *S.tmem_ptr;

    const int tpc = p.tiles_per_clip;
    const long long total_tiles = static_cast<long long>(p.n_clips) * tpc;
    const long long gstep = 4LL * gridDim.x;
    const int q = warp & 3;                                  // group = SM sub-partition = TMEM lane quarter
    float *ptile = S.ptile + q * kPWords;

    if (warp < kFftWarps) {
        // ================================================================= FFT warps
        const int wl = warp >> 2;                            // warp within the group
        const int gt = wl * 32 + lane;                       // thread within the group
        const uint32_t lane_base = tm + (static_cast<uint32_t>(32 * q) << 16);
        float *sig = S.sig + q * kSigWords;
        const float *xbase = static_cast<const float *>(p.samples);
        const bool vec_ok = p.vec_ok != 0;

        long long g = 4LL * blockIdx.x + q;                  // this group's global tile index
        if (g < total_tiles) {
            const long long clip = g / tpc, tile = g - clip * tpc;
            load_tile(sig, xbase + clip * p.clip_stride, (p.frame_begin + tile * kFT) * kHop - p.pad, p.n_samples, vec_ok, gt,
                      kFftGroupThreads, kTileSamples / 2);
        }
        for (int it = 0; g < total_tiles; g += gstep, ++it) {
            cp_async_commit_wait_all();
            tc::fence_before_sync();
            tc::bar_sync(1 + q, kFftGroupThreads);           // the tile's samples have landed; every Y row of the previous tile has been read
            tc::fence_after_sync();
#pragma unroll 1
            for (int t = wl; t < 10; t += kFftWarpsPerGroup) pass1_tm(sig, S.win, lane, t, lane_base);
            tc::wait_st();
            tc::fence_before_sync();
            tc::bar_sync(1 + q, kFftGroupThreads);           // every Y column of the group is in TMEM; the samples are dead
            tc::fence_after_sync();
            {
                const long long gn = g + gstep;              // prefetch the group's next tile into its (only) signal buffer
                if (gn < total_tiles) {
                    const long long cn = gn / tpc, tn = gn - cn * tpc;
                    load_tile(sig, xbase + cn * p.clip_stride, (p.frame_begin + tn * kFT) * kHop - p.pad, p.n_samples, vec_ok, gt,
                              kFftGroupThreads, kTileSamples / 2);
                }
            }
            if (EW > 0 && it > 0) tc::bar_sync(9 + q, kHandshake);      // the filterbank warps are done with the previous power tile
#pragma unroll 1
            for (int k1 = wl; k1 <= 10; k1 += kFftWarpsPerGroup) pass2_tm(lane_base, P.c.tw2[k1], ptile, lane, k1);
            if (EW > 0) {
                __threadfence_block();
                bar_arrive(5 + q, kHandshake);               // power tile full
            } else {
                tc::bar_sync(1 + q, kFftGroupThreads);
                const long long clip = g / tpc, tile = g - clip * tpc;
                const long long f0 = p.frame_begin + tile * kFT;
                const long long rem = p.frame_begin + p.frames_todo - f0;
                const int nf = rem < kFT ? static_cast<int>(rem) : kFT;
                float *ocf = static_cast<float *>(p.out) + clip * p.out_clip_stride + (f0 - p.out_frame_origin);
                rows_epilogue(p, ptile, S.quads, wl, nq, kFftWarpsPerGroup, ocf, nf, lane);
            }
        }
    } else if (EW > 0) {
        // ================================================================= filterbank warps
        const int el = (warp - kFftWarps) >> 2;              // filterbank warp within the group
        for (long long g = 4LL * blockIdx.x + q; g < total_tiles; g += gstep) {
            const long long clip = g / tpc, tile = g - clip * tpc;
            const long long f0 = p.frame_begin + tile * kFT;
            const long long rem = p.frame_begin + p.frames_todo - f0;
            const int nf = rem < kFT ? static_cast<int>(rem) : kFT;
            float *ocf = static_cast<float *>(p.out) + clip * p.out_clip_stride + (f0 - p.out_frame_origin);
            tc::bar_sync(5 + q, kHandshake);                 // power tile full
            rows_epilogue(p, ptile, S.quads, el, nq, EW > 0 ? EW : 1, ocf, nf, lane);
            if (g + gstep < total_tiles) bar_arrive(9 + q, kHandshake);   // power tile free again
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::dealloc(tm, kTmemCols);
}

template <int EW>
cudaError_t launch_tm(const F400Params &P, long long grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_n400_tm<EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_n400_tm<EW><<<static_cast<unsigned>(grid), 32 * (kFftWarps + kGroups * EW), smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace

// at least half an SM's worth of shared memory so that one CTA (which owns all 512 TMEM columns) is resident per SM
size_t fast400_tm_smem_bytes(int n_quads, int padded_weights) { return std::max<size_t>(tm_smem_bytes(n_quads, padded_weights), 120 * 1024); }
bool fast400_tm_fits(int n_quads, int padded_weights) { return n_quads > 0 && fast400_tm_smem_bytes(n_quads, padded_weights) <= 227 * 1024; }

cudaError_t launch_fast400_tm(const KParams &p, const float *window_f32, int n_quads, int padded_weights, int epilogue_warps,
                              int sm_count, cudaStream_t stream) {
    F400Params P;
    P.k = p;
    P.k.FT = f400::kFT;
    P.k.fd_FT = make_fastdiv(static_cast<unsigned>(f400::kFT));
    P.k.tiles_per_clip = static_cast<int>((p.frames_todo + f400::kFT - 1) / f400::kFT);
    P.k.buf_elems = padded_weights;
    fast400_fill_consts(P.c, window_f32);
    const long long total = static_cast<long long>(p.n_clips) * P.k.tiles_per_clip;
    if (total <= 0) return cudaSuccess;
    const long long grid = std::min<long long>((total + 3) / 4, sm_count);        // persistent: one CTA per SM
    const size_t smem = fast400_tm_smem_bytes(n_quads, padded_weights);
    switch (epilogue_warps) {
        case 0: return launch_tm<0>(P, grid, smem, stream);
        case 1: return launch_tm<1>(P, grid, smem, stream);
        default: return launch_tm<2>(P, grid, smem, stream);
    }
}

}  // namespace sgx
