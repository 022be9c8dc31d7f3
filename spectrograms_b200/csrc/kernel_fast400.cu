// kernel_fast400.cu -- "r2c_fused_n400": the Whisper-shaped family, n_fft = 400, hop = 160, f32 (BASELINE configs[1]
// and configs[3]). Persistent CTAs (2 per SM), 11 warps, lane = frame; each CTA walks tiles of 32 consecutive frames:
//
//   prefetch  cp.async (LDGSTS, 8-byte, zero-fill) of the NEXT tile's 5360 samples into the other signal buffer, issued
//             by warp 10 (which has no pass-1 role) -- overlaps everything below; out-of-clip bytes are zero filled, which is the reference's centre padding
//             (src/spectrogram.rs:1309-1320) with no per-tap branch
//   pass 1    warps 0..9 : window multiply + 20-point real-pair DFT in registers        (fft400_core.cuh)
//   ----      one shared-memory exchange (Y[k1][n2], 10 rows x 20 complex per frame: the two real rows share one)
//   pass 2    warps 0..10: twiddle + 20-point DFT in registers -> |X|^2 into the power tile P[bin][frame], which reuses
//             the tile's own signal buffer (its samples are dead after pass 1; the prefetch targets the other buffer)
//   epilogue  sparse filterbank rows from a shared-memory table -> sqrt / dB -> 128-byte row stores; other mappings and
//             the fused DCT-II take the general lane = frame epilogue (epilogue.cuh)
//
// Every shared-memory access of the two passes is conflict free by construction of the layouts (fft400_core.cuh);
// the window is copied once per CTA into shared memory (warp-uniform broadcast reads); the pass-2 twiddles are read from
// the constant bank (kernel parameter) with warp-uniform indices.
#include "epilogue.cuh"
#include "fast400_common.cuh"
#include "launch.hpp"

namespace sgx {

namespace {

using namespace f400;

// shared memory: two signal buffers, each big enough to be reused as the power tile of its tile once pass 1 has
// consumed the samples (kPWords >= kSigWords), the Y exchange buffer, then the plan's sparse schedule (sized per plan).
constexpr int kBufWords = kPWords > kSigWords ? kPWords : kSigWords;
constexpr size_t kFixedSmemBytes = sizeof(float) * (2 * kBufWords + kYWords);
constexpr int kPrefetchSplit = 40 * 32;       // float2 units of a tile (2680) prefetched by warp 10
constexpr size_t kSmemBudget = (233472 - 2 * 1024) / 2;      // two CTAs per SM: 228 KB per SM, 1 KB reserved per CTA

// Interior-tile prefetch by ONE warp, hop block by hop block: block b (160 samples = 80 float2 units) goes to word 162 b,
// so source and destination advance by constants and each block costs three cp.async per lane (the third on 16 lanes).
__device__ __forceinline__ void prefetch_tile_by_warp(float *sig, const float *src, int lane) {
    const float *s = src + 2 * lane;
    float *d = sig + 2 * lane;
#pragma unroll 3
    for (int b = 0; b < kTileSamples / kHop; ++b, s += kHop, d += kSigBlockStride) {      // 33 full blocks
        cp_async8(d, s, 8);
        cp_async8(d + 64, s + 64, 8);
        if (lane < 16) cp_async8(d + 128, s + 128, 8);
    }
    cp_async8(d, s, 8);                                                                   // last block: 80 samples
    if (lane < 8) cp_async8(d + 64, s + 64, 8);
}

// Fused DCT-II + lifter on the log-mel tile mtile[n][32] (mfcc_from_log_mel, src/mfcc.rs:224-273) using the basis
// symmetry B[c][n-1-i] = (-1)^c B[c][i]: fold the tile in place into E[i] = m[i] + m[n-1-i] (kept in row i) and
// O[i] = m[i] - m[n-1-i] (kept in row n-1-i), then every coefficient needs n/2 instead of n multiply-adds. One warp
// task = 4 coefficients of one parity x 32 frames (lane = frame); `basis` is the half basis [task][n/2][4], staged by the
// caller into shared memory (cp.async, overlapped with the mel rows), read as one warp-uniform 16-byte broadcast per step. Requires n even (host falls back to the general epilogue otherwise).
__device__ __forceinline__ void folded_dct_epilogue(const KParams &p, float *mtile, const float4 *basis, float *out_clip_frame, int nf,
                                                    int warp, int lane, int tid) {
    const int n = p.n_bins, half = n >> 1;
    for (int idx = tid; idx < half * kFT; idx += kThreads) {
        const int i = idx >> 5, c = idx & 31;
        const float a = mtile[i * kFT + c], b = mtile[(n - 1 - i) * kFT + c];
        mtile[i * kFT + c] = a + b;
        mtile[(n - 1 - i) * kFT + c] = a - b;
    }
    __syncthreads();
    const float *lift = static_cast<const float *>(p.lifter);
    const int col = frame_col(lane);
    const bool live = lane < nf;
    const int groups_even = ((p.n_mfcc + 1) / 2 + 3) / 4;        // tasks [0, groups_even) are even coefficients
#pragma unroll 1
    for (int task = warp; task < p.dct_tasks; task += kWarps) {
        const bool odd = task >= groups_even;
        const int g = odd ? task - groups_even : task;
        const float *m = mtile + col + (odd ? (n - 1) * kFT : 0);
        const int mstep = odd ? -kFT : kFT;
        const float4 *b = basis + static_cast<long long>(task) * half;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int i = 0; i < half; ++i) {
            const float4 w = b[i];
            const float x = m[i * mstep];
            a0 = fmaf(x, w.x, a0);
            a1 = fmaf(x, w.y, a1);
            a2 = fmaf(x, w.z, a2);
            a3 = fmaf(x, w.w, a3);
        }
        const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (odd ? 1 : 0) + 2 * (4 * g + k);
            if (c < p.n_mfcc && c >= p.mfcc_row0 && live)
                out_clip_frame[static_cast<long long>(c - p.mfcc_row0) * p.out_row_stride + lane] = acc[k] * __ldg(lift + c);
        }
    }
}

// SPARSE: mel / loghz rows served from the shared-memory table; otherwise the general epilogue.
template <bool SPARSE>
__global__ void __launch_bounds__(kThreads, 2) k_r2c_fused_n400(const __grid_constant__ F400Params P) {
    extern __shared__ __align__(16) float smem[];
    float *sig0 = smem;
    float *sig1 = sig0 + kBufWords;
    float *ybuf = sig1 + kBufWords;
    float *scratch = ybuf;                     // log-mel tile of the MFCC path: Y is dead once pass 2 is done
    const int nq_smem = SPARSE ? __ldg(reinterpret_cast<const int *>(P.k.sched)) : 0;
    int4 *s_quads = reinterpret_cast<int4 *>(ybuf + kYWords);        // [4 * n_quads] {byte offset of P[c0], cnt, weights address, row}
    int *s_qinfo = reinterpret_cast<int *>(s_quads + 4 * nq_smem);   // [kWarps + 1] quad ranges per warp, then [n_quads] max cnt
    // the window is read with warp-uniform addresses in pass 1: a shared-memory broadcast measurably beats the indexed
    // constant-bank load there (-3.6 %); the pass-2 twiddles stay in the constant bank (shared memory was slower for them)
    float *s_win = reinterpret_cast<float *>(s_qinfo + ((kWarps + 1 + nq_smem + 3) & ~3));
    float *s_w = s_win + kN;                                         // weights, rows padded to multiples of 4
    const KParams &p = P.k;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool vec_ok = p.vec_ok != 0;

    for (int i = tid; i < kN; i += kThreads) s_win[i] = P.c.win[i];
    if (SPARSE) {
        // p.sched is the host-built schedule blob: int n_quads; int qrange[kWarps + 1]; int maxcnt[n_quads];
        // (16-byte aligned) int4 {c0, cnt, padded weight offset, row or -1}[4 * n_quads]
        const int *blob = reinterpret_cast<const int *>(p.sched);
        const int nq = __ldg(blob);
        const int hdr = (1 + kWarps + 1 + nq + 3) & ~3;
        const int4 *quads = reinterpret_cast<const int4 *>(blob + hdr);
        const float *val = static_cast<const float *>(p.val);
        const unsigned wbase = smem_u32(s_w);
        for (int i = tid; i < kWarps + 1 + nq; i += kThreads) s_qinfo[i] = __ldg(blob + 1 + i);
        for (int i = tid; i < 4 * nq; i += kThreads) {
            const int4 e = __ldg(quads + i);
            s_quads[i] = make_int4(e.x * (kFT * 4), e.y, static_cast<int>(wbase + 4u * e.z), e.w);
            if (e.w >= 0) {
                const int e0 = __ldg(p.row_ptr + e.w);
                for (int k = 0; k < ((e.y + 3) & ~3); ++k) s_w[e.z + k] = k < e.y ? __ldg(val + e0 + k) : 0.f;
            }
        }
    }

    // persistent walk over a CONTIGUOUS run of tiles [t_cur, t_end) per CTA (the slower first / last tiles of the clips spread evenly
    // over the CTAs; a grid-strided walk can resonate with the tiles per clip); (clip, tile) advance by carry instead of dividing
    const int tpc = p.tiles_per_clip;
    const long long total_tiles = static_cast<long long>(p.n_clips) * tpc;
    long long t_cur = blockIdx.x * total_tiles / gridDim.x;
    const long long t_end = (blockIdx.x + 1LL) * total_tiles / gridDim.x;
    int clip = t_cur < t_end ? static_cast<int>(t_cur / tpc) : p.n_clips, tile = static_cast<int>(t_cur % tpc);
    const float *xbase = static_cast<const float *>(p.samples);
    int buf = 0;
    if (clip < p.n_clips)
        load_tile(sig0, xbase + static_cast<long long>(clip) * p.clip_stride,
                  (p.frame_begin + static_cast<long long>(tile) * kFT) * kHop - p.pad, p.n_samples, vec_ok, tid, kThreads, kTileSamples / 2);
    for (; clip < p.n_clips; buf ^= 1) {
        const long long f0 = p.frame_begin + static_cast<long long>(tile) * kFT;
        const long long rem = p.frame_begin + p.frames_todo - f0;
        const int nf = rem < kFT ? static_cast<int>(rem) : kFT;
        float *sig = buf ? sig1 : sig0;
        float *ptile = sig;                    // the signal buffer becomes this tile's power tile after pass 1
        const int cur_clip = clip;

        cp_async_commit_wait_all();
        __syncthreads();                       // tile t has landed; everyone is done with the previous tile's P

        ++tile;                                // next tile of this CTA
        if (tile >= tpc) { tile = 0; ++clip; }
        if (++t_cur >= t_end) clip = p.n_clips;
        // Prefetch of the next tile into the other buffer, overlapped with pass 1: warp 10 has no pass-1 role and issues
        // the whole tile (about as many instructions as one pass-1 task). Edge tiles (first / last of a clip, or unaligned
        // input) take the general path, shared by all warps.
        const bool more = clip < p.n_clips;
        const float *xn = xbase + static_cast<long long>(more ? clip : 0) * p.clip_stride;
        const long long sn = (p.frame_begin + static_cast<long long>(tile) * kFT) * kHop - p.pad;
        const bool interior = vec_ok && sn >= 0 && sn + kTileSamples <= p.n_samples;
        if (warp < 10) {
            pass1_task(sig, ybuf, s_win, lane, warp);
            if (more && !interior) load_tile(buf ? sig0 : sig1, xn, sn, p.n_samples, vec_ok, kPrefetchSplit + tid, 320, kTileSamples / 2);
        } else if (more) {
            if (interior) prefetch_tile_by_warp(buf ? sig0 : sig1, xn + sn, lane);
            else load_tile(buf ? sig0 : sig1, xn, sn, p.n_samples, vec_ok, lane, 32, kPrefetchSplit);
        }
        __syncthreads();

        {
            float2 v[20];
            pass2_load(ybuf, P.c.tw2[warp], lane, warp, v);
            pass2_finish(v, ptile, lane, warp);       // samples are consumed: P overwrites the signal buffer
        }
        __syncthreads();

        float *ocf = static_cast<float *>(p.out) + static_cast<long long>(cur_clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
        if (SPARSE && p.output != SGX_OUT_MFCC) {
            if (p.apply_db) sparse_quads_epilogue<2, false>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, nullptr, nf, lane);
            else if (p.amp == SGX_AMP_MAGNITUDE) sparse_quads_epilogue<1, false>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, nullptr, nf, lane);
            else sparse_quads_epilogue<0, false>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, nullptr, nf, lane);
        } else if (SPARSE && p.dct_folded != nullptr) {
            // fused mfcc(): mel -> dB (always has a floor here, src/mfcc.rs:371) -> folded DCT-II -> lifter
            // stage the half basis behind the log-mel tile (the Y buffer is dead after pass 2) while the mel rows run
            float4 *s_basis = reinterpret_cast<float4 *>(scratch + p.n_bins * kFT);
            const int n_basis = p.dct_tasks * (p.n_bins >> 1);
            for (int i = tid; i < n_basis; i += kThreads) {
                const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(s_basis + i));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(static_cast<const float4 *>(p.dct_folded) + i) : "memory");
            }
            if (p.apply_db) sparse_quads_epilogue<2, true>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, scratch, nf, lane);
            else if (p.amp == SGX_AMP_MAGNITUDE) sparse_quads_epilogue<1, true>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, scratch, nf, lane);
            else sparse_quads_epilogue<0, true>(p, ptile, s_quads, s_qinfo[warp], s_qinfo[warp + 1], 1, ocf, scratch, nf, lane);
            cp_async_commit_wait_all();
            __syncthreads();
            folded_dct_epilogue(p, scratch, s_basis, ocf, nf, warp, lane, tid);
        } else {
            epilogue_lane_frames<float>(p, ptile, scratch, cur_clip, f0, nf, frame_col(lane));
        }
    }
}

}  // namespace

// dynamic shared memory of a launch whose sparse schedule has n_quads quads and padded_weights weights (0, 0: none)
size_t fast400_smem_bytes(int n_quads, int padded_weights) {
    if (n_quads == 0) return kFixedSmemBytes + sizeof(float) * (f400::kN + 16);
    return kFixedSmemBytes + sizeof(int4) * 4 * n_quads + sizeof(int) * ((f400::kWarps + 1 + n_quads + 3) & ~3) +
           sizeof(float) * (padded_weights + 8 + f400::kN);
}
// the sparse schedule is used only while two CTAs still fit on an SM
bool fast400_sparse_fits(int n_quads, int padded_weights) { return fast400_smem_bytes(n_quads, padded_weights) <= kSmemBudget; }
int fast400_max_scratch_rows() { return f400::kYWords / 32; }
int fast400_warps() { return f400::kWarps; }


cudaError_t launch_fast400(const KParams &p, const float *window_f32, bool sparse_table, int n_quads, int padded_weights,
                           int sm_count, cudaStream_t stream) {
    static_assert(sizeof(F400Params) <= 4096, "kernel parameter block must fit the classic 4 KiB limit");
    F400Params P;
    P.k = p;
    P.k.FT = f400::kFT;
    P.k.fd_FT = make_fastdiv(static_cast<unsigned>(f400::kFT));
    P.k.tiles_per_clip = static_cast<int>((p.frames_todo + f400::kFT - 1) / f400::kFT);
    fast400_fill_consts(P.c, window_f32);
    const long long total = static_cast<long long>(p.n_clips) * P.k.tiles_per_clip;
    if (total <= 0) return cudaSuccess;
    const long long grid = std::min<long long>(total, 2LL * sm_count);     // persistent: 2 CTAs per SM
    const size_t smem = sparse_table ? fast400_smem_bytes(n_quads, padded_weights) : fast400_smem_bytes(0, 0);
    cudaError_t e;
    if (sparse_table) {
        e = cudaFuncSetAttribute(k_r2c_fused_n400<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_r2c_fused_n400<true><<<static_cast<unsigned>(grid), f400::kThreads, smem, stream>>>(P);
    } else {
        e = cudaFuncSetAttribute(k_r2c_fused_n400<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_r2c_fused_n400<false><<<static_cast<unsigned>(grid), f400::kThreads, smem, stream>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace sgx
