// epilogue.cuh -- the fused back half of every kernel family: power tile in shared memory ->
// frequency mapping -> amplitude scaling -> (DCT-II + lifter) -> bin-major coalesced store.
//
// Reference semantics restated here (bare :N = src/spectrogram.rs:N):
//   FrequencyMapping::apply :1822-1881, SparseMatrix::multiply_vec :102-117 (ascending columns, acc += T(w)*x without
//   FMA), ErbFilterbank::apply_to_power_spectrum src/erb.rs:384-398, AmplitudeScaling::apply_in_place :2068-2080,
//   column write data[[row, frame]] :285-287, mfcc_from_log_mel src/mfcc.rs:224-273.
//
// Thread mapping: the frame index is the fastest-varying index across lanes, so every global store is a run of
// consecutive frames of one output row (the reference layout is (rows, n_frames) with frames contiguous), and every
// shared-memory read of the tile is lane-stride `tile_stride` (odd -> conflict free). Filterbank weights are
// warp-uniform and come through the read-only path.
#pragma once

#include "kparams.cuh"

namespace sgx {

// P: power tile, P[f * p.tile_stride + k], f < nf valid frames (rows beyond nf are not read)
// scratch: second tile region with at least FT * tile_stride elements (used for the log-mel tile when output == MFCC)
template <typename T>
__device__ __forceinline__ void epilogue_from_power(const KParams &p, T *__restrict__ P, T *__restrict__ scratch,
                                                    int clip, long long f0, int nf) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int FT = p.FT;
    const T eps = static_cast<T>(p.eps);
    T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
    const bool to_mfcc = (p.output == SGX_OUT_MFCC);
    const int ts = p.tile_stride;

    if (p.mapping == SGX_MAP_CHROMA) {
        // chromagram() (src/chroma.rs:487-503): magnitude tile -> 12 dense rows in ascending bin order (:384-394) ->
        // per-frame normalisation (:406-453). scratch holds the un-normalised rows [f * ts + c].
        for (int idx = tid; idx < p.out_len * FT; idx += nthr) {
            const int f = fd_div(idx, p.fd_out_len), k = idx - f * p.out_len;
            if (f < nf) P[f * ts + k] = t_sqrt(P[f * ts + k]);
        }
        __syncthreads();
        for (int idx = tid; idx < 12 * FT; idx += nthr) {
            const int row = fd_div(idx, p.fd_FT), f = idx - row * FT;
            if (f >= nf) continue;
            const T *w = static_cast<const T *>(p.dense) + static_cast<long long>(row) * p.out_len;
            const T *pf = P + f * ts;
            T acc = T(0);                                  // columns outside [dense_c0, dense_c1) are exact zeros
            for (int k = p.dense_c0; k < p.dense_c1; ++k) acc = t_add_rn(acc, t_mul_rn(__ldg(w + k), pf[k]));
            scratch[f * ts + row] = acc;
        }
        __syncthreads();
        for (int f = tid; f < nf; f += nthr) {
            T c[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) c[i] = scratch[f * ts + i];
            chroma_normalise<T>(c, p.chroma_norm);
#pragma unroll
            for (int i = 0; i < 12; ++i) out[static_cast<long long>(i) * p.out_row_stride + f] = c[i];
        }
        return;
    }

    // rows x frames, frames fastest
    const int total = p.n_bins * FT;
    for (int idx = tid; idx < total; idx += nthr) {
        const int row = fd_div(idx, p.fd_FT);
        const int f = idx - row * FT;
        if (f >= nf) continue;
        const T *pf = P + f * ts;
        T acc;
        if (p.mapping == SGX_MAP_LINEAR) {
            acc = pf[row];
        } else if (p.mapping == SGX_MAP_ERB) {
            const T *w = static_cast<const T *>(p.dense) + static_cast<long long>(row) * p.out_len;
            acc = T(0);
            for (int k = 0; k < p.out_len; ++k) acc = t_add_rn(acc, t_mul_rn(__ldg(w + k), pf[k]));
        } else {
            const T *val = static_cast<const T *>(p.val);
            acc = T(0);
            if (p.rows_contig) {
                // consecutive columns (every mel / loghz row): one descriptor load per row, unrolled weight / tile loads
                const int4 d = __ldg(p.row_desc + row);          // {first entry, count, first column}
                const T *w = val + d.x;
                const T *px = pf + d.z;
#pragma unroll 4
                for (int i = 0; i < d.y; ++i) acc = t_add_rn(acc, t_mul_rn(__ldg(w + i), px[i]));
            } else {
                const int e0 = __ldg(p.row_ptr + row), e1 = __ldg(p.row_ptr + row + 1);
                for (int e = e0; e < e1; ++e) acc = t_add_rn(acc, t_mul_rn(__ldg(val + e), pf[__ldg(p.col + e)]));
            }
        }
        acc = amp_scale<T>(acc, p.amp, p.apply_db, eps);
        if (to_mfcc) scratch[f * ts + row] = acc;
        else out[static_cast<long long>(row) * p.out_row_stride + f] = acc;
    }
    if (!to_mfcc) return;
    __syncthreads();
    // DCT-II over the mel axis, keep n_mfcc rows, lifter, optional c0 drop
    const T *dct = static_cast<const T *>(p.dct);
    const T *lift = static_cast<const T *>(p.lifter);
    const int rows = p.n_mfcc - p.mfcc_row0;
    for (int idx = tid; idx < rows * FT; idx += nthr) {
        const int r = fd_div(idx, p.fd_FT);
        const int f = idx - r * FT;
        if (f >= nf) continue;
        const int c = r + p.mfcc_row0;
        const T *mf = scratch + f * ts;
        const T *b = dct + static_cast<long long>(c) * p.n_bins;
        T acc = T(0);
        for (int i = 0; i < p.n_bins; ++i) acc = t_fma(mf[i], __ldg(b + i), acc);   // val.mul_add(basis, acc)
        out[static_cast<long long>(r) * p.out_row_stride + f] = acc * __ldg(lift + c);
    }
}

// S: complex spectrum tile, S[f * p.frame_stride + k]; StftPlan::compute column copy (:1440-1442)
template <typename T>
__device__ __forceinline__ void epilogue_complex(const KParams &p, const typename Cplx<T>::type *__restrict__ S, int clip,
                                                 long long f0, int nf) {
    using C = typename Cplx<T>::type;
    const int FT = p.FT;
    C *out = static_cast<C *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
    const int total = p.out_len * FT;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int k = fd_div(idx, p.fd_FT);
        const int f = idx - k * FT;
        if (f >= nf) continue;
        out[static_cast<long long>(k) * p.out_row_stride + f] = S[f * p.frame_stride + k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant for kernel families whose tile is "frames fastest", exactly 32 frames per tile, lane = frame. One warp owns one
// output row at a time, so every filterbank weight / column index / DCT coefficient is warp-uniform and every store is a
// 128-byte run of one output row. The tile is reached through two accessors bound to the calling lane's frame:
// tile(k) = power of bin k (read / write), scratch(r) = row r of a second tile of >= n_bins rows (MFCC only).
template <typename T, typename Tile, typename Scratch>
__device__ __forceinline__ void epilogue_lane_frames_via(const KParams &p, Tile tile, Scratch scratch, int clip, long long f0, int nf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const T eps = static_cast<T>(p.eps);
    T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin) + lane;
    const bool to_mfcc = (p.output == SGX_OUT_MFCC);
    const bool live = lane < nf;

    if (p.mapping == SGX_MAP_CHROMA) {
        // same three steps as in epilogue_from_power; scratch rows hold the un-normalised pitch classes
        for (int k = warp; k < p.out_len; k += nwarps) tile(k) = t_sqrt(tile(k));
        __syncthreads();
        for (int row = warp; row < 12; row += nwarps) {
            const T *w = static_cast<const T *>(p.dense) + static_cast<long long>(row) * p.out_len;
            T acc = T(0);
            for (int k = p.dense_c0; k < p.dense_c1; ++k) acc = t_add_rn(acc, t_mul_rn(__ldg(w + k), tile(k)));
            scratch(row) = acc;
        }
        __syncthreads();
        if (warp == 0) {
            T c[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) c[i] = scratch(i);
            chroma_normalise<T>(c, p.chroma_norm);
            if (live) {
#pragma unroll
                for (int i = 0; i < 12; ++i) out[static_cast<long long>(i) * p.out_row_stride] = c[i];
            }
        }
        return;
    }
    if ((p.mapping == SGX_MAP_MEL || p.mapping == SGX_MAP_LOGHZ) && p.rows_contig && (p.n_bins + nwarps - 1) / nwarps <= 32) {
        // Rows with consecutive columns (every mel / loghz row). The global-memory latency is taken off the per-row chain:
        // lane i fetches the descriptor of the warp's i-th row once, and the weights of the NEXT row (lane i = weight i, one
        // coalesced load) while the current row is summed; both reach the other lanes through shuffles. Accumulation order
        // and rounding are those of the loop below (ascending columns, acc += T(w) * x without FMA).
        const T *val = static_cast<const T *>(p.val);
        const int my_rows = (p.n_bins - warp + nwarps - 1) / nwarps;          // rows warp, warp + nwarps, ...
        int4 mine = make_int4(0, 0, 0, 0);
        if (lane < my_rows) mine = __ldg(p.row_desc + warp + lane * nwarps);  // {first entry, count, first column}
        int e0n = __shfl_sync(0xffffffffu, mine.x, 0), cntn = __shfl_sync(0xffffffffu, mine.y, 0), c0n = __shfl_sync(0xffffffffu, mine.z, 0);
        T wn = (my_rows > 0 && lane < cntn) ? __ldg(val + e0n + lane) : T(0);
        for (int r = 0; r < my_rows; ++r) {
            const int row = warp + r * nwarps;
            const int e0 = e0n, cnt = cntn, c0 = c0n;
            const T wv = wn;
            if (r + 1 < my_rows) {
                e0n = __shfl_sync(0xffffffffu, mine.x, r + 1);
                cntn = __shfl_sync(0xffffffffu, mine.y, r + 1);
                c0n = __shfl_sync(0xffffffffu, mine.z, r + 1);
                wn = lane < cntn ? __ldg(val + e0n + lane) : T(0);
            }
            T acc = T(0);
            const int head = cnt < 32 ? cnt : 32;
#pragma unroll 4
            for (int i = 0; i < head; ++i) acc = t_add_rn(acc, t_mul_rn(__shfl_sync(0xffffffffu, wv, i), tile(c0 + i)));
            for (int i = 32; i < cnt; ++i) acc = t_add_rn(acc, t_mul_rn(__ldg(val + e0 + i), tile(c0 + i)));
            acc = amp_scale<T>(acc, p.amp, p.apply_db, eps);
            if (to_mfcc) scratch(row) = acc;
            else if (live) out[static_cast<long long>(row) * p.out_row_stride] = acc;
        }
    } else
    for (int row = warp; row < p.n_bins; row += nwarps) {
        T acc;
        if (p.mapping == SGX_MAP_LINEAR) {
            acc = tile(row);
        } else if (p.mapping == SGX_MAP_ERB) {
            const T *w = static_cast<const T *>(p.dense) + static_cast<long long>(row) * p.out_len;
            acc = T(0);
            for (int k = 0; k < p.out_len; ++k) acc = t_add_rn(acc, t_mul_rn(__ldg(w + k), tile(k)));
        } else {
            const T *val = static_cast<const T *>(p.val);
            const int e0 = __ldg(p.row_ptr + row), e1 = __ldg(p.row_ptr + row + 1);
            acc = T(0);
            for (int e = e0; e < e1; ++e) acc = t_add_rn(acc, t_mul_rn(__ldg(val + e), tile(__ldg(p.col + e))));
        }
        acc = amp_scale<T>(acc, p.amp, p.apply_db, eps);
        if (to_mfcc) scratch(row) = acc;
        else if (live) out[static_cast<long long>(row) * p.out_row_stride] = acc;
    }
    if (!to_mfcc) return;
    __syncthreads();
    const T *dct = static_cast<const T *>(p.dct);
    const T *lift = static_cast<const T *>(p.lifter);
    const int rows = p.n_mfcc - p.mfcc_row0;
    for (int r = warp; r < rows; r += nwarps) {
        const int c = r + p.mfcc_row0;
        const T *b = dct + static_cast<long long>(c) * p.n_bins;
        T acc = T(0);
        for (int i = 0; i < p.n_bins; ++i) acc = t_fma(scratch(i), __ldg(b + i), acc);
        if (live) out[static_cast<long long>(r) * p.out_row_stride] = acc * __ldg(lift + c);
    }
}

// the plain layout: P[k * 32 + f], scratch [row * 32 + f] (>= n_bins * 32 elements, MFCC only).
// lane_col: the column of this lane's frame inside a tile row (identity unless the family permutes frames in a row).
template <typename T> struct LinearTileAccess {
    T *base;
    __device__ __forceinline__ T &operator()(int k) const { return base[k * 32]; }
};
template <typename T>
__device__ __forceinline__ void epilogue_lane_frames(const KParams &p, T *__restrict__ P, T *__restrict__ scratch,
                                                     int clip, long long f0, int nf, int lane_col) {
    epilogue_lane_frames_via<T>(p, LinearTileAccess<T>{P + lane_col}, LinearTileAccess<T>{scratch + (threadIdx.x & 31)}, clip, f0, nf);
}

}  // namespace sgx
