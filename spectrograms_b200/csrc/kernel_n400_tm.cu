// kernel_n400_tm.cu -- "r2c_fused_n400_tm": n_fft = 400, hop = 160, f32 (BASELINE configs[1]) with the Blackwell tensor
// memory (TMEM) as the exchange medium between the two FFT passes and warp-specialised filterbank warps.
//
// Why: the shared-memory form (kernel_fast400.cu) is bound by the L1 / shared-memory data pipe: 3480 wavefronts per 32-frame
// tile, of which 840 are the Y exchange between the passes (profiles/r1_n400_phase_ablation.md). TMEM has its own data
// path (tcgen05.st 794 B/clk, tcgen05.ld 57 B/clk per SM, tools/ubench/tmem_probe.cu), so the exchange leaves the L1 pipe,
// the 52 KB exchange buffer leaves shared memory, and four tiles instead of two are in flight per SM.
//
//   one persistent CTA per SM = 4 groups x 4 warps; group q = the warps with warp_id % 4 == q (SM sub-partition q, TMEM
//   lanes 32q .. 32q+31). A group walks its own tiles of 32 consecutive frames (lane = frame = TMEM lane) and never
//   synchronises with the other groups: the four sub-partitions run their FFT-bound and load/store-bound phases out of step.
//   Two group barriers per tile:
//
//   phase B   prefetch (cp.async) of the next tile's samples, then pass 2 of tile t: tcgen05.ld (LDTM) of one Y row ->
//             twiddle + 20-point DFT -> |X|^2 into the group's power tile P[bin][frame] in shared memory
//   phase A   the filterbank rows of tile t (sparse mel / loghz rows from the power tile, the quad schedule of
//             kernel_fast400.cu -> sqrt / dB -> row stores) AND pass 1 of tile t+1 (window + 20-point real-pair DFT in
//             registers -> Y[k1][n2] into TMEM columns 0..399 of the thread's own lane, tcgen05.st / STTM) share one
//             phase with no barrier between them: every warp runs its rows, then its FFT tasks (the warps drift apart by the
//             different lengths of their row lists, which is what mixes load/store-bound and FP32-bound work on a scheduler).
//
// Arithmetic is the arithmetic of kernel_fast400.cu (same task functions, same epilogue); only where Y travels differs.
// Measured steps of this design (dedicated filterbank warps behind full / free barriers lost: one or two such warps per
// group are latency bound) are in profiles/r2_n400_tm_experiments.md.
#include <cstdlib>
#include <vector>

#include "fast400_common.cuh"
#include "launch.hpp"
#include "tcgen05.cuh"

namespace sgx {
namespace {

using namespace f400;

constexpr int kGroups = 4;
constexpr int kMaxGroupWarps = 8;
constexpr uint32_t kTmemCols = 512;                         // Y needs 400 columns; allocations are powers of two
constexpr int kPadRows = 8;                                 // zero rows behind the power tile (padded quad rows read them)
constexpr int kPWordsTm = (kBins + kPadRows) * kFT;
// Bulk-staged signal tile (SIG = 1): hop block b of the tile lives at word 164 b (4 pad words = 16 bytes per hop), so every
// block is a legal cp.async.bulk destination and lanes = frames read 16-byte sample quads conflict free (164 / 4 odd).
constexpr int kSigStrideB = 164;
constexpr int kSigWordsTm = kSigBlocks * kSigStrideB;      // 5576 >= kSigWords: both layouts fit
constexpr uint32_t kTileBytes = kTileSamples * 4u;

// kernel parameter block: the window comes from the plan's device copy (it is staged into shared memory once per CTA anyway),
// which leaves room for the pass-2 twiddles in both forms inside the classic 4 KiB
struct TmParams {
    KParams k;
    float2 tw2[10][20];     // s(k1) * W400^(n2 k1) (f400::Consts::tw2) for k1 = 1 .. 10 (row 0 is all ones)
    float2 tw2r[9][20];     // i * tw2 = (-tw2.y, tw2.x) for k1 = 1 .. 9 (row 10 multiplies real values)
};
static_assert(sizeof(TmParams) <= 4096, "kernel parameter block must fit the classic 4 KiB limit");

struct TmSmem {
    float *sig;        // [4][kSigWordsTm] one signal tile per group
    float *ptile;      // [4][kPWordsTm]   one power tile per group (+ zero rows)
    float *win;        // [400]
    int4 *quads;       // [4 * n_quads]    {byte offset of P[c0], padded cnt, weights address, row}
    int *qrange;       // [group warps + 1] quad range of every warp of a group
    float *w;          // padded weights
    uint64_t *bars;    // [4] one mbarrier per group: bulk-copy completion of the group's next signal tile
    uint32_t *tmem_ptr;
};

__host__ __device__ inline size_t tm_smem_bytes(int n_quads, int padded_weights) {
    return sizeof(float) * (kGroups * (kSigWordsTm + kPWordsTm) + kN) + sizeof(int4) * 4 * static_cast<size_t>(n_quads + 3) +
           sizeof(int) * 12 + sizeof(float) * static_cast<size_t>(padded_weights + 8) + 8 + sizeof(uint64_t) * kGroups + 16;
}

__device__ __forceinline__ TmSmem carve(unsigned char *base, int n_quads, int padded_weights) {
    TmSmem s;
    size_t o = 0;
    s.sig = reinterpret_cast<float *>(base + o);        o += sizeof(float) * kGroups * kSigWordsTm;
    s.ptile = reinterpret_cast<float *>(base + o);      o += sizeof(float) * kGroups * kPWordsTm;
    s.quads = reinterpret_cast<int4 *>(base + o);       o += sizeof(int4) * 4 * static_cast<size_t>(n_quads + 3);    // + 3 look-ahead copies
    s.win = reinterpret_cast<float *>(base + o);        o += sizeof(float) * kN;
    s.qrange = reinterpret_cast<int *>(base + o);       o += sizeof(int) * 12;
    s.w = reinterpret_cast<float *>(base + o);          o += sizeof(float) * static_cast<size_t>(padded_weights + 8);
    o = (o + 7) & ~static_cast<size_t>(7);
    s.bars = reinterpret_cast<uint64_t *>(base + o);    o += sizeof(uint64_t) * kGroups;
    s.tmem_ptr = reinterpret_cast<uint32_t *>(base + o);
    return s;
}

// ---- pass 1, one task = (frame = lane, column pair t): f400::pass1_task with Y going to the thread's TMEM lane.
// Y layout (columns): row 0 = (Y[0][n2], Y[10][n2]) pairs, rows 1..9 = Y[k1][n2] complex; column 40 * row + 2 * n2 (+1).
// STRIDE: words between hop blocks of the signal tile (162, or 164 in the bulk-staged layout, where these 8-byte reads are
// two-way bank conflicted -- only the two tasks that do not fill a quad use them there).
__device__ __forceinline__ void pass1_store(const float2 (&v)[20], uint32_t y) {
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
template <int STRIDE = kSigBlockStride>
__device__ __forceinline__ void pass1_tm(const float *__restrict__ sig, const float *__restrict__ win, int f, int t, uint32_t ybase) {
    float2 v[20];
    const float *s = sig + STRIDE * f + 2 * t;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const float2 x = *reinterpret_cast<const float2 *>(s + 20 * n1 + (STRIDE - kHop) * (n1 / 8));
        const float2 w = *reinterpret_cast<const float2 *>(win + 20 * n1 + 2 * t);
        v[n1] = cmul2(x, w);                          // sample * window[i] (src/spectrogram.rs:1319)
    }
    dft20(v);
    pass1_store(v, ybase + 4 * t);
}
// One task = (frame = lane, column quad u): columns n2 = 4u .. 4u+3 = the column pairs 2u and 2u+1 from ONE 16-byte sample read
// and one 16-byte window read per n1 (bulk-staged layout only): half the shared-memory instructions of two pair tasks.
__device__ __forceinline__ void pass1_quad_tm(const float *__restrict__ sig, const float *__restrict__ win, int f, int u, uint32_t ybase) {
    float2 va[20], vb[20];
    const float *s = sig + kSigStrideB * f + 4 * u;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const float4 x = *reinterpret_cast<const float4 *>(s + 20 * n1 + (kSigStrideB - kHop) * (n1 / 8));
        const float4 w = *reinterpret_cast<const float4 *>(win + 20 * n1 + 4 * u);
        va[n1] = cmul2(make_float2(x.x, x.y), make_float2(w.x, w.y));     // sample * window[i] (src/spectrogram.rs:1319)
        vb[n1] = cmul2(make_float2(x.z, x.w), make_float2(w.z, w.w));
    }
    dft20(va);
    dft20(vb);
    const uint32_t y = ybase + 8 * u;
    {
        const float2 a0 = va[reg_of_bin(0)], a10 = va[reg_of_bin(10)], b0 = vb[reg_of_bin(0)], b10 = vb[reg_of_bin(10)];
        tc::st8(y, __float_as_uint(a0.x), __float_as_uint(a10.x), __float_as_uint(a0.y), __float_as_uint(a10.y), __float_as_uint(b0.x),
                __float_as_uint(b10.x), __float_as_uint(b0.y), __float_as_uint(b10.y));
    }
#pragma unroll
    for (int k1 = 1; k1 < 10; ++k1) {
        const float2 A = va[reg_of_bin(k1)], B = va[reg_of_bin(20 - k1)], C = vb[reg_of_bin(k1)], D = vb[reg_of_bin(20 - k1)];
        const float2 sa = cadd(A, make_float2(B.x, -B.y)), sb = cadd(make_float2(A.y, -A.x), make_float2(B.y, B.x));
        const float2 sc = cadd(C, make_float2(D.x, -D.y)), sd = cadd(make_float2(C.y, -C.x), make_float2(D.y, D.x));
        tc::st8(y + 40 * k1, __float_as_uint(sa.x), __float_as_uint(sa.y), __float_as_uint(sb.x), __float_as_uint(sb.y), __float_as_uint(sc.x),
                __float_as_uint(sc.y), __float_as_uint(sd.x), __float_as_uint(sd.y));
    }
}

// The next tile's samples by the TMA unit's 1-D form (cp.async.bulk, SASS UBLKCP): one thread of the group issues one copy per
// hop block (34 per tile) against the group's mbarrier -- no per-lane address arithmetic, no LDGSTS issue slots.
__device__ __forceinline__ void bulk_tile(float *sig, const float *src, uint64_t *bar) {
    // Called by a whole converged warp with warp-uniform operands; one elected lane issues (operands in uniform registers, each
    // copy a single UBLKCP -- issued from divergent code every copy is wrapped in an elect / broadcast loop).
    // No fence.proxy.async before the copies: the tile is only *read* through the generic proxy (pass 1, whose loads have returned
    // before the group barrier the issuing warp has just passed), and generic-proxy writes to it (edge / unaligned tiles) are a
    // whole tile and two group barriers old. A buffer that is refilled after its readers have released it needs no proxy fence
    // (the consumer-release / producer-acquire hand-over of any TMA load pipeline); the fence compiled to a MEMBAR.ALL.CTA per
    // tile on the critical path of the group's phase B and cost 0.9 % (same box, alternating: 0.9974 -> 0.9884 ms; SGX_TM_FENCE
    // puts it back).
    const uint32_t b = tc::smem_addr(bar), d = tc::smem_addr(sig);
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
#ifdef SGX_TM_FENCE
        "@q fence.proxy.async.shared::cta;\n"
#endif
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "}\n" ::"r"(b),
        "r"(kTileBytes)
        : "memory");
#pragma unroll
    for (int blk = 0; blk < kSigBlocks; ++blk) {
        const uint32_t bytes = blk < kSigBlocks - 1 ? kHop * 4u : (kTileSamples - (kSigBlocks - 1) * kHop) * 4u;
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "elect.sync _|q, 0xffffffff;\n"
            "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
            "}\n" ::"r"(d + blk * (kSigStrideB * 4u)),
            "l"(src + blk * kHop), "r"(bytes), "r"(b)
            : "memory");
    }
}
// First / last tiles of a clip (centre zero padding, clip end): the hop blocks [b0, b1) that lie wholly inside the clip still come
// by bulk copy (same calling convention as bulk_tile; src = the tile's first sample, possibly before the clip) ...
__device__ __forceinline__ void bulk_blocks(float *sig, const float *src, uint64_t *bar, int b0, int b1) {
    const uint32_t b = tc::smem_addr(bar), d = tc::smem_addr(sig);
    const uint32_t total = static_cast<uint32_t>(b1 - b0) * (kHop * 4u) - (b1 == kSigBlocks ? (kSigBlocks * kHop - kTileSamples) * 4u : 0u);
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
#ifdef SGX_TM_FENCE
        "@q fence.proxy.async.shared::cta;\n"
#endif
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "}\n" ::"r"(b),
        "r"(total)
        : "memory");
#pragma unroll 1
    for (int blk = b0; blk < b1; ++blk) {
        const uint32_t bytes = blk < kSigBlocks - 1 ? kHop * 4u : (kTileSamples - (kSigBlocks - 1) * kHop) * 4u;
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "elect.sync _|q, 0xffffffff;\n"
            "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
            "}\n" ::"r"(d + blk * (kSigStrideB * 4u)),
            "l"(src + blk * kHop), "r"(bytes), "r"(b)
            : "memory");
    }
}
// ... and the group's threads fill the others: zeros where a block lies wholly outside the clip, sample by sample (zero outside,
// the reference's centre zero padding src/spectrogram.rs:1309-1320) for the at most two blocks that straddle an end of the clip.
__device__ __forceinline__ void edge_blocks(float *sig, const float *x, long long sn, long long n, int b0, int b1, int gt, int threads) {
#pragma unroll 1
    for (int blk = 0; blk < kSigBlocks; ++blk) {
        if (blk == b0 && b0 < b1) { blk = b1 - 1; continue; }
        const long long sb = sn + static_cast<long long>(blk) * kHop;
        const int len = blk < kSigBlocks - 1 ? kHop : kTileSamples - (kSigBlocks - 1) * kHop;
        float *d = sig + blk * kSigStrideB;
        if (sb + len <= 0 || sb >= n) {
            for (int i = gt; i < len / 4; i += threads) reinterpret_cast<float4 *>(d)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = gt; i < len; i += threads) {
                const long long sidx = sb + i;
                d[i] = (sidx >= 0 && sidx < n) ? __ldg(x + sidx) : 0.f;
            }
        }
    }
}

// ---- pass 2, one task = (frame = lane, k1): Y row from TMEM -> twiddle -> DFT20 -> |X|^2 into P[bin][frame].
// The TMEM read port of a sub-partition delivers 14 B/clk (a 40-column row of 32 lanes takes ~360 cycles); the other warps of
// the group cover it. (Requesting the warp's next row while the butterfly runs was measured slower: 1.37 against 1.32 ms.)
__device__ __forceinline__ void y_row_request(uint32_t ybase, int k1, uint32_t (&q)[40]) {
    const uint32_t row = ybase + ((k1 == 0 || k1 == 10) ? 0 : 40 * k1);
    tc::ld32(row, q);
    tc::ld8(row + 32, q + 32);
}
// tcgen05.wait::ld, with the row registers as read-write operands so that no use of them can be scheduled above the wait
__device__ __forceinline__ void y_row_wait(uint32_t (&q)[40]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9]),
                   "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15]), "+r"(q[16]), "+r"(q[17]), "+r"(q[18]),
                   "+r"(q[19])
                 :
                 : "memory");
    asm volatile(""
                 : "+r"(q[20]), "+r"(q[21]), "+r"(q[22]), "+r"(q[23]), "+r"(q[24]), "+r"(q[25]), "+r"(q[26]), "+r"(q[27]), "+r"(q[28]),
                   "+r"(q[29]), "+r"(q[30]), "+r"(q[31]), "+r"(q[32]), "+r"(q[33]), "+r"(q[34]), "+r"(q[35]), "+r"(q[36]), "+r"(q[37]),
                   "+r"(q[38]), "+r"(q[39])
                 :
                 : "memory");
}
// tw2 / tw2r: this k1's twiddles w and i w = (-w.y, w.x) from the constant bank (forming the rotation in registers cost an FADD per
// twiddle; reading both from shared memory cost 200 wavefronts per tile on the L1 pipe, the busiest unit)
__device__ __forceinline__ void pass2_twiddle(const uint32_t (&q)[40], const float2 *__restrict__ tw2, const float2 *__restrict__ tw2r, int k1,
                                              float2 (&v)[20]) {
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
            v[2 * j] = cfma2(bc2(__uint_as_float(q[4 * j + 1])), tw2r[2 * j], cmul2(bc2(__uint_as_float(q[4 * j])), tw2[2 * j]));
            v[2 * j + 1] = cfma2(bc2(__uint_as_float(q[4 * j + 3])), tw2r[2 * j + 1], cmul2(bc2(__uint_as_float(q[4 * j + 2])), tw2[2 * j + 1]));
        }
    }
}

// |X[bin]|^2 of the bins a pass-2 butterfly owns into P[bin][f] (f400::pass2_finish with the address arithmetic hoisted): bins
// k1 + 20 k2 (k2 < 10) going up from k1, bin 200 - k1 (k2 = 10) and the mirrored bins (200 - k1) - 20 (k2 - 10) going down from
// it; k1 = 0 / 10 own no mirrored bins (conjugate duplicates) and k1 = 10 would repeat bin 190.
__device__ __forceinline__ void pass2_finish_tm(float2 (&v)[20], float *__restrict__ ptile, int f, int k1) {
    dft20(v);
    float pw[20];
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) {
        const float2 X = v[reg_of_bin(k2)];
        pw[k2] = norm_sqr(X);                         // norm_sqr = re*re + im*im (src/spectrogram.rs:1332-1334)
    }
    float *up = ptile + frame_col(f) + kFT * k1;
    float *down = ptile + frame_col(f) + kFT * (200 - k1);
#pragma unroll
    for (int k2 = 0; k2 < 10; ++k2) up[kFT * 20 * k2] = pw[k2];
    if (k1 != 10) down[0] = pw[10];
    if (k1 != 0 && k1 != 10) {
#pragma unroll
        for (int k2 = 11; k2 < 20; ++k2) down[-kFT * 20 * (k2 - 10)] = pw[k2];
    }
}

// Interior-tile prefetch by the four warps of a group, hop block by hop block: block b (160 samples = 80 float2 units)
// goes to word 162 b, so source and destination advance by constants and a block costs three cp.async per lane.
template <int GW>
__device__ __forceinline__ void prefetch_tile_by_group(float *sig, const float *src, int wl, int lane) {
#pragma unroll 1
    for (int b = wl; b < kSigBlocks; b += GW) {
        const float *s = src + b * kHop + 2 * lane;
        float *d = sig + b * kSigBlockStride + 2 * lane;
        if (b < kSigBlocks - 1) {                      // 33 full blocks
            cp_async8(d, s, 8);
            cp_async8(d + 64, s + 64, 8);
            if (lane < 16) cp_async8(d + 128, s + 128, 8);
        } else {                                       // last block: 80 samples
            cp_async8(d, s, 8);
            if (lane < 8) cp_async8(d + 64, s + 64, 8);
        }
    }
}

__device__ __forceinline__ void rows_epilogue(const KParams &p, const float *ptile, const int4 *quads, int q0, int q1, float *ocf, int nf,
                                              int lane) {
    if (p.apply_db) sparse_quads_pipelined<2>(p, ptile, quads, q0, q1, ocf, nf, lane);
    else if (p.amp == SGX_AMP_MAGNITUDE) sparse_quads_pipelined<1>(p, ptile, quads, q0, q1, ocf, nf, lane);
    else sparse_quads_pipelined<0>(p, ptile, quads, q0, q1, ocf, nf, lane);
}

// GW: warps per group (4; 5 / 6 for experiments). SIG: 0 = signal tile staged by cp.async (LDGSTS) in the 162-word layout, 1 = by
// cp.async.bulk in the 164-word layout with 16-byte pass-1 reads (GW = 4, 16-byte aligned input).
template <int GW, int SIG>
__global__ void __launch_bounds__(kGroups * GW * 32, 1) k_r2c_fused_n400_tm(const __grid_constant__ TmParams P) {
    constexpr int kPadW = SIG ? kSigStrideB - kHop : kSigBlockStride - kHop;
    constexpr int kGroupWarps = GW, kTmThreads = kGroups * GW * 32, kGroupThreads = GW * 32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const KParams &p = P.k;
    // p.sched: int n_quads; int qrange[kGroupWarps + 1]; int maxcnt[n_quads]; pad to 16 bytes; int4 {c0, padded cnt, weight offset, row}[4 n]
    const int *blob = reinterpret_cast<const int *>(p.sched);
    const int nq = __ldg(blob);
    const int padded_weights = p.buf_elems;
    const TmSmem S = carve(smem_raw, nq, padded_weights);
    // the warp index through a shuffle: the compiler then treats it -- and the group, tile and address arithmetic that follows
    // from it -- as warp-uniform (uniform registers, single-instruction bulk copies)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

    // ---- one-time setup: tables -> shared memory, zero rows behind the power tiles, TMEM allocation
    {
        const int hdr = (1 + kGroupWarps + 1 + nq + 3) & ~3;
        const int4 *gq = reinterpret_cast<const int4 *>(blob + hdr);
        const float *val = static_cast<const float *>(p.val);
        const unsigned wbase = smem_u32(S.w);
        for (int i = tid; i < kN; i += kTmThreads) S.win[i] = __ldg(static_cast<const float *>(p.window) + i);
        if (tid <= kGroupWarps) S.qrange[tid] = __ldg(blob + 1 + tid);
        for (int i = tid; i < kGroups * kPadRows * kFT; i += kTmThreads)
            S.ptile[(i / (kPadRows * kFT)) * kPWordsTm + kBins * kFT + i % (kPadRows * kFT)] = 0.f;
        for (int i = tid; i < 4 * nq; i += kTmThreads) {
            const int4 e = __ldg(gq + i);
            const int e0 = e.w >= 0 ? __ldg(p.row_ptr + e.w) : 0;
            const int cnt = e.w >= 0 ? __ldg(p.row_ptr + e.w + 1) - e0 : 0;
            // .y: the quad's padded (warp-uniform) column count, and above it the row's own count (tile reads beyond it are skipped)
            const int4 q4 = make_int4(e.x * (kFT * 4), e.y | (cnt << 16), static_cast<int>(wbase + 4u * e.z), e.w);
            S.quads[i] = q4;
            if (i >= 4 * (nq - 1)) S.quads[i + 4] = S.quads[i + 8] = S.quads[i + 12] = q4;      // the rows' look-ahead reads stay on valid quads
#ifdef SGX_TM_DEVW
            for (int k = 0; k < ((e.y + 3) & ~3); ++k) S.w[e.z + k] = k < cnt ? __ldg(val + e0 + k) : 0.f;
#endif
        }
#ifndef SGX_TM_DEVW
        {   // the padded weights, laid out by the host behind the quad table (sgx_api.cu): one coalesced copy
            const float *hw = reinterpret_cast<const float *>(blob + hdr + 16 * nq);
            for (int i = tid; i < padded_weights; i += kTmThreads) S.w[i] = __ldg(hw + i);
            (void)val;
        }
#endif
        if (SIG && tid < kGroups) tc::mbar_init(S.bars + tid, 1);
        if (SIG && tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (warp == 0) tc::alloc(S.tmem_ptr, kTmemCols);
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t tm = *S.tmem_ptr;

    const int tpc = p.tiles_per_clip;
    const int total_tiles = p.n_clips * tpc;                 // the host splits batches so that this fits an int
    // Every group owns a CONTIGUOUS run of tiles [g, g_end): the first / last tiles of the clips (slower: zero-filled cp.async
    // staging, partial stores) then spread evenly over the groups. With a grid-strided walk a group whose stride resonates with
    // the tiles per clip (592 groups, 32 tiles per 10 s clip: the same two tile positions for ever) took the kernel's time.
    const long long n_groups = 4LL * gridDim.x;
    const int q = warp & 3, wl = warp >> 2;                  // group = SM sub-partition = TMEM lane quarter; warp within the group
    const int gt = wl * 32 + lane;                           // thread within the group
    const uint32_t lane_base = tm + (static_cast<uint32_t>(32 * q) << 16);
    float *sig = S.sig + q * kSigWordsTm;
    uint64_t *bar_sig = S.bars + q;
    uint32_t bulk_parity = 0;                                // completed bulk-staged tiles of this group, mod 2
    const bool bulk_ok = SIG && (p.vec_ok & 2);
    float *ptile = S.ptile + q * kPWordsTm;
    const float *xbase = static_cast<const float *>(p.samples);
    const bool vec_ok = p.vec_ok != 0;
    const int q0 = S.qrange[wl], q1 = S.qrange[wl + 1];
    const int bar = 1 + q;

    // pass 1 of a whole tile by this warp: column pairs t = wl, wl + GW, ... (SIG = 0); column quad wl and, on warp 0, quad 4 as
    // well (SIG = 1; the host deals that warp correspondingly fewer filterbank rows)
    auto pass1_all = [&](const float *sg, const float *wn, int f, int w, uint32_t yb) {
        if (SIG) {
            pass1_quad_tm(sg, wn, f, w, yb);
            if (w == 0) pass1_quad_tm(sg, wn, f, 4, yb);
        } else {
#pragma unroll 1
            for (int t = w; t < 10; t += kGroupWarps) pass1_tm(sg, wn, f, t, yb);
        }
    };
    const long long gid = 4LL * blockIdx.x + q;
    int g = static_cast<int>(gid * total_tiles / n_groups);  // this group's global tile index
    const int g_end = static_cast<int>((gid + 1) * total_tiles / n_groups);
    int clip = g / tpc, tile = g % tpc;                      // ... and its (clip, tile); both advance by carry
    if (g < g_end) {                                   // prologue: pass 1 of the group's first tile
        load_tile<kPadW>(sig, xbase + static_cast<long long>(clip) * p.clip_stride, (p.frame_begin + static_cast<long long>(tile) * kFT) * kHop - p.pad, p.n_samples, vec_ok, gt,
                         kGroupThreads, kTileSamples / 2);
        cp_async_commit_wait_all();
        tc::bar_sync(bar, kGroupThreads);
        pass1_all(sig, S.win, lane, wl, lane_base);
        tc::wait_st();
        tc::fence_before_sync();
        tc::bar_sync(bar, kGroupThreads);
        tc::fence_after_sync();
    }
    for (; g < g_end; ++g) {
        // ---- phase B: Y(g) is complete in TMEM, the samples are dead, the power tile is free
        const bool has_next = g + 1 < g_end;
        bool bulk_now = false;
        int cn = clip, tn = tile + 1;
        if (tn >= tpc) { tn = 0; ++cn; }
        if (has_next) {
            const long long sn = (p.frame_begin + static_cast<long long>(tn) * kFT) * kHop - p.pad;
            const float *xn = xbase + static_cast<long long>(cn) * p.clip_stride;
            const bool interior = sn >= 0 && sn + kTileSamples <= p.n_samples;
            if (SIG) {
                if (bulk_ok && interior) {
                    bulk_now = true;
                    if (wl == kGroupWarps - 1) bulk_tile(sig, xn + sn, bar_sig);       // the warp with the fewest pass-2 rows
                } else if (bulk_ok) {
                    // hop blocks [b0, b1) lie wholly inside the clip
                    const int b0 = sn < 0 ? static_cast<int>((-sn + kHop - 1) / kHop) : 0;
                    const long long fb = p.n_samples > sn ? (p.n_samples - sn) / kHop : 0;
                    int b1 = fb < kSigBlocks - 1 ? static_cast<int>(fb) : kSigBlocks - 1;
                    if (b1 == kSigBlocks - 1 && sn + kTileSamples <= p.n_samples) b1 = kSigBlocks;
                    if (b1 < b0) b1 = b0;
                    bulk_now = b1 > b0;
                    if (bulk_now && wl == kGroupWarps - 1) bulk_blocks(sig, xn + sn, bar_sig, b0, b1);
                    edge_blocks(sig, xn, sn, p.n_samples, b0, b1, gt, kGroupThreads);
                } else {
                    load_tile<kPadW>(sig, xn, sn, p.n_samples, vec_ok, gt, kGroupThreads, kTileSamples / 2);
                }
            } else if (vec_ok && interior) {
                prefetch_tile_by_group<GW>(sig, xn + sn, wl, lane);
            } else {
                load_tile<kPadW>(sig, xn, sn, p.n_samples, vec_ok, gt, kGroupThreads, kTileSamples / 2);
            }
        }
#pragma unroll 1
        for (int k1 = wl; k1 <= 10; k1 += kGroupWarps) {
            uint32_t yq[40];
            float2 v[20];
            y_row_request(lane_base, k1, yq);
            y_row_wait(yq);
            pass2_twiddle(yq, P.tw2[k1 ? k1 - 1 : 0], P.tw2r[k1 - 1 < 9u ? k1 - 1 : 0], k1, v);
            pass2_finish_tm(v, ptile, lane, k1);
        }
        cp_async_commit_wait_all();
        if (SIG && bulk_now) {
            tc::mbar_wait(bar_sig, bulk_parity);
            bulk_parity ^= 1u;
        }
        tc::fence_before_sync();
        tc::bar_sync(bar, kGroupThreads);                    // P(g) complete, every Y row read, the next tile's samples have landed
        tc::fence_after_sync();
        // ---- phase A: the filterbank rows of tile g and pass 1 of the next tile, in opposite orders on alternate warps
        const long long f0 = p.frame_begin + static_cast<long long>(tile) * kFT;
        const long long rem = p.frame_begin + p.frames_todo - f0;
        const int nf = rem < kFT ? static_cast<int>(rem) : kFT;
        float *ocf = static_cast<float *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
#ifndef SGX_TM_ORDER
#define SGX_TM_ORDER 4
#endif
        // which warps take their rows before their FFT tasks (measured, same box, alternating: every warp rows first 1.020 ms, odd
        // warps first 1.046, even warps first 1.046, warps 2 / 3 first 1.054, none first 1.073 -- profiles/r2_n400_tm_experiments.md)
        const bool rows_first = SGX_TM_ORDER == 0 ? (wl & 1) != 0 : SGX_TM_ORDER == 1 ? (wl & 1) == 0 : SGX_TM_ORDER == 2 ? wl >= 2 : SGX_TM_ORDER == 3 ? wl < 2 : SGX_TM_ORDER == 4;
        if (rows_first) rows_epilogue(p, ptile, S.quads, q0, q1, ocf, nf, lane);
        if (has_next) pass1_all(sig, S.win, lane, wl, lane_base);
        if (!rows_first) rows_epilogue(p, ptile, S.quads, q0, q1, ocf, nf, lane);
        tc::wait_st();
        tc::fence_before_sync();
        tc::bar_sync(bar, kGroupThreads);                    // Y(next) complete; the power tile and the samples are free again
        tc::fence_after_sync();
        clip = cn;
        tile = tn;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::dealloc(tm, kTmemCols);
}

}  // namespace

// at least half an SM's worth of shared memory so that one CTA (which owns all 512 TMEM columns) is resident per SM
size_t fast400_tm_smem_bytes(int n_quads, int padded_weights) { return std::max<size_t>(tm_smem_bytes(n_quads, padded_weights), 120 * 1024); }
bool fast400_tm_fits(int n_quads, int padded_weights) { return n_quads > 0 && fast400_tm_smem_bytes(n_quads, padded_weights) <= 227 * 1024; }
int fast400_tm_max_group_warps() { return kMaxGroupWarps; }
int fast400_tm_pad_rows() { return kPadRows; }

namespace {
template <int GW, int SIG = 0>
cudaError_t launch_tm(const TmParams &P, long long grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_n400_tm<GW, SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_n400_tm<GW, SIG><<<static_cast<unsigned>(grid), kGroups * GW * 32, smem, stream>>>(P);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_fast400_tm(const KParams &p, const float *window_f32, int n_quads, int padded_weights, int group_warps, int sm_count,
                              cudaStream_t stream) {
    TmParams P;
    P.k = p;
    P.k.FT = f400::kFT;
    P.k.fd_FT = make_fastdiv(static_cast<unsigned>(f400::kFT));
    P.k.tiles_per_clip = static_cast<int>((p.frames_todo + f400::kFT - 1) / f400::kFT);
    P.k.buf_elems = padded_weights;
    {
        static const f400::Consts c = [] { f400::Consts t; std::vector<float> w(f400::kN, 0.f); fast400_fill_consts(t, w.data()); return t; }();
        for (int k1 = 1; k1 <= 10; ++k1)
            for (int n2 = 0; n2 < 20; ++n2) {
                P.tw2[k1 - 1][n2] = c.tw2[k1][n2];
                if (k1 < 10) P.tw2r[k1 - 1][n2] = make_float2(-c.tw2[k1][n2].y, c.tw2[k1][n2].x);
            }
    }
    (void)window_f32;                                       // the kernel reads the plan's device window (p.window)
    const long long total = static_cast<long long>(p.n_clips) * P.k.tiles_per_clip;
    if (total <= 0) return cudaSuccess;
    const long long grid = std::min<long long>((total + 3) / 4, sm_count);        // persistent: one CTA per SM
    const size_t smem = fast400_tm_smem_bytes(n_quads, padded_weights);
    // bulk-copy staging needs 16-byte aligned tiles: base, clip stride and the centre padding (200 or 0 samples; a tile starts at
    // a multiple of 160 samples minus the padding)
    static const bool bulk_off = std::getenv("SGX_N400_TM_SIG") && std::atoi(std::getenv("SGX_N400_TM_SIG")) == 0;
    const bool bulk = !bulk_off && group_warps == 4 && (p.vec_ok & 1) && reinterpret_cast<uintptr_t>(p.samples) % 16 == 0 && p.clip_stride % 4 == 0 &&
                      p.pad % 4 == 0;
    if (bulk) {
        P.k.vec_ok |= 2;
        return launch_tm<4, 1>(P, grid, smem, stream);
    }
    switch (group_warps) {
        case 4: return launch_tm<4>(P, grid, smem, stream);
        case 5: return launch_tm<5>(P, grid, smem, stream);
        case 6: return launch_tm<6>(P, grid, smem, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sgx
