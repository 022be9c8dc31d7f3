// fast400_common.cuh -- pieces shared by the two kernels of the n_fft = 400 / hop = 160 f32 family
// (kernel_fast400.cu: shared-memory exchange + CUDA-core filterbank; kernel_n400_tc.cu: TMEM exchange + tcgen05 filterbank).
#pragma once

#include <cmath>

#include "fft400_core.cuh"
#include "kparams.cuh"

namespace sgx {

struct F400Params {
    KParams k;
    f400::Consts c;
};

// per-plan constants of both kernels: the f32 window and the pass-2 twiddles s(k1) W400^(n2 k1) (extended precision -> f32)
inline void fast400_fill_consts(f400::Consts &c, const float *window_f32) {
    for (int i = 0; i < f400::kN; ++i) c.win[i] = window_f32[i];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int n2 = 0; n2 < 20; ++n2) {
            const long double a = -2.0L * pi * static_cast<long double>((n2 * k1) % 400) / 400.0L;
            const double s = (k1 == 0 || k1 == 10) ? 1.0 : 0.5;
            c.tw2[k1][n2] = make_float2(static_cast<float>(s * static_cast<double>(cosl(a))),
                                        static_cast<float>(s * static_cast<double>(sinl(a))));
        }
}

namespace f400 {

__device__ __forceinline__ void cp_async8(float *dst_smem, const float *src, int src_bytes) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst_smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// stage one tile's samples [s0, s0 + 5360) of a clip into a padded signal buffer
// float2 units j = first, first + step, ... < last of the tile are handled by the calling thread. PADW = pad words per hop
// block (2: the 162-word layout of fft400_core.cuh; 4: the 164-word layout of the bulk-staged r2c_fused_n400_tm)
template <int PADW = 2>
__device__ __forceinline__ void load_tile(float *sig, const float *x, long long s0, long long n, bool vec_ok, int first, int step,
                                          int last) {
    if (vec_ok && s0 >= 0 && s0 + kTileSamples <= n) {
        // interior tile (all but the first / last tile of a clip): no bounds logic at all
        const float *src = x + s0;
#pragma unroll 4
        for (int j = first; j < last; j += step) cp_async8(sig + 2 * j + PADW * (j / (kHop / 2)), src + 2 * j, 8);
    } else if (vec_ok) {
        for (int j = first; j < last; j += step) {
            const long long s = s0 + 2 * j;
            const long long avail = n - s;                // samples available from s on
            const int bytes = (s < 0 || avail <= 0) ? 0 : (avail >= 2 ? 8 : 4);
            cp_async8(sig + 2 * j + PADW * (j / (kHop / 2)), bytes ? x + s : x, bytes);
        }
    } else {
        for (int j = first; j < last; j += step) {
            const long long s = s0 + 2 * j;
            float2 v;
            v.x = (s >= 0 && s < n) ? __ldg(x + s) : 0.f;
            v.y = (s + 1 >= 0 && s + 1 < n) ? __ldg(x + s + 1) : 0.f;
            *reinterpret_cast<float2 *>(sig + 2 * j + PADW * (j / (kHop / 2))) = v;
        }
    }
}

// lg2.approx.ftz: the argument is clamped to eps > 0 first, so the denormal fix-up of __log2f is dead weight
__device__ __forceinline__ float fast_lg2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// AMP: 0 power, 1 magnitude, 2 dB (AmpScale::apply_from_power + apply_db_in_place, src/spectrogram.rs:1986-2037, :2068-2080)
template <int AMP>
__device__ __forceinline__ float finish_value(float acc, float eps) {
    if (AMP == 1) acc = sqrtf(acc);
    if (AMP == 2) acc = 3.01029995663981195f * fast_lg2(fmaxf(acc, eps));
    return acc;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_v4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// Sparse filterbank rows (mel triangles, loghz interpolation pairs: every row's columns are contiguous), "quad"
// schedule: one warp step = 4 rows x 32 frames. Lane (s, j) = (lane >> 3, lane & 7) owns row s of the quad and frames
// j, j+8, j+16, j+24, which the permuted power tile hands over in ONE 16-byte read per column (4 wavefronts per warp
// instruction = the minimum for 512 bytes). Each lane carries four independent accumulators (ILP 4), the row's weights
// and bookkeeping are loaded once for four outputs, and every store instruction writes four 32-byte runs. Rows are sorted by column count
// on the host, so a quad is nearly homogeneous; the per-lane trip count keeps the reference's order of accumulation:
// ascending columns, acc += T(w) * x (SparseMatrix::multiply_vec, src/spectrogram.rs:102-117), here with one fused
// rounding per term (fmaf) -- never further from the exact sum than the reference's two roundings.
// AMP: 0 power, 1 magnitude, 2 dB.
// TO_SMEM: instead of storing to global memory, leave the scaled rows in a shared tile mtile[row][32] (same permuted
// frame order as the power tile) for the fused DCT.
template <int AMP, bool TO_SMEM, bool FULL>
__device__ __forceinline__ void sparse_quads_epilogue_impl(const KParams &p, const float *ptile, const int4 *s_quads, int q0, int q1, int qstep,
                                                           float *out_clip_frame, float *mtile, int nf, int lane, unsigned wrel = 0u) {
    const float eps = static_cast<float>(p.eps);
    const int s = lane >> 3, j = lane & 7;
    const unsigned pbase = smem_u32(ptile) + 16u * j;             // columns 4j..4j+3 = frames j, j+8, j+16, j+24
    const unsigned qbase = smem_u32(s_quads);
    const unsigned ors4 = 4u * static_cast<unsigned>(p.out_row_stride);
    char *ob = reinterpret_cast<char *>(out_clip_frame) + 4 * j;
#pragma unroll 1
    for (int qi = q0; qi < q1; qi += qstep) {
        const float4 rf = lds_v4(qbase + 16u * (4 * qi + s));      // {byte offset of P[c0], cnt, weights address, row}
        const int cnt = __float_as_int(rf.y);
        const unsigned pe = pbase + __float_as_uint(rf.x);
        const unsigned wa = __float_as_uint(rf.z) + wrel;      // wrel: base of the weights when the table holds relative addresses
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        // per-lane trip count: the divergent loop branch masks finished rows, no predicate inside the body. Four columns per
        // step: the row's weights are padded to a multiple of 4, so one 16-byte broadcast load serves four columns; the
        // remainder (< 4 columns) runs column by column so that no tile element beyond the row is ever touched.
        const int cnt4 = cnt >> 2;
#pragma unroll 1
        for (int e4 = 0; e4 < cnt4; ++e4) {
            const float4 w = lds_v4(wa + 16u * e4);
            const float4 x0 = lds_v4(pe + (kFT * 4u) * (4 * e4));
            const float4 x1 = lds_v4(pe + (kFT * 4u) * (4 * e4 + 1));
            const float4 x2 = lds_v4(pe + (kFT * 4u) * (4 * e4 + 2));
            const float4 x3 = lds_v4(pe + (kFT * 4u) * (4 * e4 + 3));
            a0 = fmaf(w.x, x0.x, a0); a1 = fmaf(w.x, x0.y, a1); a2 = fmaf(w.x, x0.z, a2); a3 = fmaf(w.x, x0.w, a3);
            a0 = fmaf(w.y, x1.x, a0); a1 = fmaf(w.y, x1.y, a1); a2 = fmaf(w.y, x1.z, a2); a3 = fmaf(w.y, x1.w, a3);
            a0 = fmaf(w.z, x2.x, a0); a1 = fmaf(w.z, x2.y, a1); a2 = fmaf(w.z, x2.z, a2); a3 = fmaf(w.z, x2.w, a3);
            a0 = fmaf(w.w, x3.x, a0); a1 = fmaf(w.w, x3.y, a1); a2 = fmaf(w.w, x3.z, a2); a3 = fmaf(w.w, x3.w, a3);
        }
#pragma unroll 1
        for (int e = 4 * cnt4; e < cnt; ++e) {
            const float w = lds_f32(wa + 4u * e);
            const float4 x = lds_v4(pe + (kFT * 4u) * e);
            a0 = fmaf(w, x.x, a0);
            a1 = fmaf(w, x.y, a1);
            a2 = fmaf(w, x.z, a2);
            a3 = fmaf(w, x.w, a3);
        }
        const int row = __float_as_int(rf.w);
        if (row >= 0) {
            const float v0 = finish_value<AMP>(a0, eps), v1 = finish_value<AMP>(a1, eps);
            const float v2 = finish_value<AMP>(a2, eps), v3 = finish_value<AMP>(a3, eps);
            if (TO_SMEM) {
                *reinterpret_cast<float4 *>(mtile + row * kFT + 4 * j) = make_float4(v0, v1, v2, v3);
                continue;
            }
            char *orow = ob + static_cast<size_t>(static_cast<unsigned>(row)) * ors4;
            if (FULL || j < nf) *reinterpret_cast<float *>(orow) = v0;
            if (FULL || j + 8 < nf) *reinterpret_cast<float *>(orow + 32) = v1;
            if (FULL || j + 16 < nf) *reinterpret_cast<float *>(orow + 64) = v2;
            if (FULL || j + 24 < nf) *reinterpret_cast<float *>(orow + 96) = v3;
        }
    }
}

// The same rows, software-pipelined for the kernels whose power tile is followed by zero rows (r2c_fused_n400_tm): every
// row of a quad carries the quad's longest count rounded up to 4 columns (zero weights behind the row's own), all column
// steps are 4 wide, and the first step of quad i+1 -- its descriptor, four power-tile reads and four weights -- is loaded
// while quad i is multiplied, scaled and stored, so the shared-memory latency leaves the per-quad dependency chain. The
// loop handles two quads per trip with two register sets (no rotation moves); the multiply-adds are packed (two frames per
// FFMA2, weight as the broadcast operand). Every quad slot holds a real row (the host pads the last quad with copies of one
// of its rows, which store the same values twice), so there is no per-row branch.
struct QuadRegs {
    float4 d;                      // {byte offset of P[c0], cnt (multiple of 4), weights address, row}
    float4 w, x0, x1, x2, x3;      // first 4-column step
    unsigned pe;
};
// 16-byte shared-memory read if col < cnt, else the register keeps its (finite) old contents: it then meets a zero weight. A lane
// group of 8 = one row = one quarter-warp wavefront, so a skipped read is a wavefront the L1 pipe never sees.
template <int OFS>
__device__ __forceinline__ void lds_v4_if(float4 &v, unsigned a, int col, int cnt) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.lt.s32 p, %5, %6;\n"
        "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %7];\n"
        "}\n"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "r"(a), "r"(col), "r"(cnt), "n"(OFS));
}
template <int OFS>
__device__ __forceinline__ float4 lds_v4_ofs(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFS));
    return v;
}
__device__ __forceinline__ void quad_fetch(QuadRegs &r, float4 d, unsigned pbase) {
    constexpr int kRow = kFT * 4;
    r.d = d;
    r.pe = pbase + __float_as_uint(d.x);
    const int cnt = __float_as_int(d.y) >> 16;         // this row's own column count (>= 1 for every row the host schedules)
    r.w = lds_v4(__float_as_uint(d.z));
    r.x0 = lds_v4_ofs<0>(r.pe);
    lds_v4_if<kRow>(r.x1, r.pe, 1, cnt);
    lds_v4_if<2 * kRow>(r.x2, r.pe, 2, cnt);
    lds_v4_if<3 * kRow>(r.x3, r.pe, 3, cnt);
}
template <int AMP, bool FULL>
__device__ __forceinline__ void quad_finish(QuadRegs &r, float eps, char *ob, unsigned ors4, int j, int nf) {
    constexpr unsigned kRow = kFT * 4u;
    float2 lo = make_float2(0.f, 0.f), hi = lo;        // frames (j, j+8) and (j+16, j+24)
    lo = cfma2(bc2(r.w.x), make_float2(r.x0.x, r.x0.y), lo); hi = cfma2(bc2(r.w.x), make_float2(r.x0.z, r.x0.w), hi);
    lo = cfma2(bc2(r.w.y), make_float2(r.x1.x, r.x1.y), lo); hi = cfma2(bc2(r.w.y), make_float2(r.x1.z, r.x1.w), hi);
    lo = cfma2(bc2(r.w.z), make_float2(r.x2.x, r.x2.y), lo); hi = cfma2(bc2(r.w.z), make_float2(r.x2.z, r.x2.w), hi);
    lo = cfma2(bc2(r.w.w), make_float2(r.x3.x, r.x3.y), lo); hi = cfma2(bc2(r.w.w), make_float2(r.x3.z, r.x3.w), hi);
    const int steps = (__float_as_int(r.d.y) & 0xffff) >> 2, cnt = __float_as_int(r.d.y) >> 16;
#pragma unroll 1
    for (int e4 = 1; e4 < steps; ++e4) {               // rows longer than four columns (warp-uniform count)
        const float4 we = lds_v4(__float_as_uint(r.d.z) + 16u * e4);
        // into the (consumed) registers of the first step: whatever finite values a skipped read leaves there meet zero weights
        float4 &z0 = r.x0, &z1 = r.x1, &z2 = r.x2, &z3 = r.x3;
        const unsigned pz = r.pe + kRow * (4 * e4);
        lds_v4_if<0>(z0, pz, 4 * e4, cnt);
        lds_v4_if<kRow>(z1, pz, 4 * e4 + 1, cnt);
        lds_v4_if<2 * kRow>(z2, pz, 4 * e4 + 2, cnt);
        lds_v4_if<3 * kRow>(z3, pz, 4 * e4 + 3, cnt);
        lo = cfma2(bc2(we.x), make_float2(z0.x, z0.y), lo); hi = cfma2(bc2(we.x), make_float2(z0.z, z0.w), hi);
        lo = cfma2(bc2(we.y), make_float2(z1.x, z1.y), lo); hi = cfma2(bc2(we.y), make_float2(z1.z, z1.w), hi);
        lo = cfma2(bc2(we.z), make_float2(z2.x, z2.y), lo); hi = cfma2(bc2(we.z), make_float2(z2.z, z2.w), hi);
        lo = cfma2(bc2(we.w), make_float2(z3.x, z3.y), lo); hi = cfma2(bc2(we.w), make_float2(z3.z, z3.w), hi);
    }
    float v0, v1, v2, v3;
    if (AMP == 2) {
        const float2 l = cmul2(bc2(3.01029995663981195f), make_float2(fast_lg2(fmaxf(lo.x, eps)), fast_lg2(fmaxf(lo.y, eps))));
        const float2 h = cmul2(bc2(3.01029995663981195f), make_float2(fast_lg2(fmaxf(hi.x, eps)), fast_lg2(fmaxf(hi.y, eps))));
        v0 = l.x; v1 = l.y; v2 = h.x; v3 = h.y;
    } else {
        v0 = finish_value<AMP>(lo.x, eps); v1 = finish_value<AMP>(lo.y, eps);
        v2 = finish_value<AMP>(hi.x, eps); v3 = finish_value<AMP>(hi.y, eps);
    }
    char *orow = ob + static_cast<size_t>(__float_as_uint(r.d.w)) * ors4;
    if (FULL || j < nf) *reinterpret_cast<float *>(orow) = v0;
    if (FULL || j + 8 < nf) *reinterpret_cast<float *>(orow + 32) = v1;
    if (FULL || j + 16 < nf) *reinterpret_cast<float *>(orow + 64) = v2;
    if (FULL || j + 24 < nf) *reinterpret_cast<float *>(orow + 96) = v3;
}
template <int AMP, bool FULL>
__device__ __forceinline__ void sparse_quads_pipelined_impl(const KParams &p, const float *ptile, const int4 *s_quads, int q0, int q1,
                                                            float *out_clip_frame, int nf, int lane) {
    if (q0 >= q1) return;
    const float eps = static_cast<float>(p.eps);
    const int s = lane >> 3, j = lane & 7;
    const unsigned pbase = smem_u32(ptile) + 16u * j;             // columns 4j..4j+3 = frames j, j+8, j+16, j+24
    const unsigned qbase = smem_u32(s_quads) + 16u * s;
    const unsigned ors4 = 4u * static_cast<unsigned>(p.out_row_stride);
    char *ob = reinterpret_cast<char *>(out_clip_frame) + 4 * j;
    // The descriptor table is followed by three more valid quads (the next warp's, or the copies the kernel appends behind the
    // last one), so the look-ahead reads below need no clamping: a quad fetched beyond q1 is simply never finished.
    QuadRegs A, B;
    A.x1 = A.x2 = A.x3 = B.x1 = B.x2 = B.x3 = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned qa = qbase + 64u * q0;
    quad_fetch(A, lds_v4(qa), pbase);
    float4 dn = lds_v4(qa + 64u);                                  // descriptor one quad ahead of the fetches
#pragma unroll 1
    for (int qi = q0; qi < q1; qi += 2, qa += 128u) {
        quad_fetch(B, dn, pbase);                                  // quad qi + 1
        dn = lds_v4(qa + 128u);
        quad_finish<AMP, FULL>(A, eps, ob, ors4, j, nf);
        quad_fetch(A, dn, pbase);                                  // quad qi + 2
        dn = lds_v4(qa + 192u);
        if (qi + 1 < q1) quad_finish<AMP, FULL>(B, eps, ob, ors4, j, nf);
    }
}

template <int AMP>
__device__ __forceinline__ void sparse_quads_pipelined(const KParams &p, const float *ptile, const int4 *s_quads, int q0, int q1,
                                                       float *out_clip_frame, int nf, int lane) {
    if (nf == kFT) sparse_quads_pipelined_impl<AMP, true>(p, ptile, s_quads, q0, q1, out_clip_frame, nf, lane);
    else sparse_quads_pipelined_impl<AMP, false>(p, ptile, s_quads, q0, q1, out_clip_frame, nf, lane);
}

template <int AMP, bool TO_SMEM>
__device__ __forceinline__ void sparse_quads_epilogue(const KParams &p, const float *ptile, const int4 *s_quads, int q0, int q1, int qstep,
                                                      float *out_clip_frame, float *mtile, int nf, int lane, unsigned wrel = 0u) {
    if (nf == kFT) sparse_quads_epilogue_impl<AMP, TO_SMEM, true>(p, ptile, s_quads, q0, q1, qstep, out_clip_frame, mtile, nf, lane, wrel);
    else sparse_quads_epilogue_impl<AMP, TO_SMEM, false>(p, ptile, s_quads, q0, q1, qstep, out_clip_frame, mtile, nf, lane, wrel);
}

}  // namespace f400
}  // namespace sgx
