// kernel_mixed.cu -- "r2c_fused_mixed": even n_fft = 2 L with L = R1 R2 R3 a product of small primes (2, 3, 5) that is
// neither a power of two >= 256 nor the 400 / 160 f32 shape: 64 ... 1600 (f32) / 64 ... 800 (f64), e.g. 400 in f64 or at
// any hop, 480, 800, 960, 1000, 1200 (the sizes the reference benches beside the powers of two,
// benches/fft1d_benchmarks.rs:163-171). The per-frame work is SpectrogramPlan::compute (src/spectrogram.rs:240-294):
// gather + centre zero padding + window (:1301-1320), real FFT (src/fft_backend.rs:423-431, here a packed L-point complex
// FFT plus the split post pass), |X|^2 (:1332-1334), mapping, scaling, column store.
//
// Layout: one CTA = one tile of 32 consecutive frames of one clip, LANE = FRAME. The tile's packed spectra live in shared
// memory as z[element][frame] with a row stride of 33 complex values: a warp reading or writing one element of all 32
// frames touches 32 consecutive words (conflict free), and the load phase, where a warp walks along the elements of ONE
// frame (coalesced global reads), is conflict free as well because of the odd stride. Everything that depends on the
// position inside a frame -- butterfly index, twiddle, filterbank row, output row -- is warp uniform: no per-thread index
// arithmetic, twiddles come through uniform loads, and every store is a 128-byte run of one output row.
//
//   load     frame by frame: x[2n], x[2n+1] (coalesced 8/16-byte reads, zero outside the clip) times the window -> z[n][f]
//   stages   decimation-in-frequency, IN PLACE, radix R1, R2 (, R3) register butterflies (mixed_dft.cuh: compile-time
//            Cooley-Tukey on 2^a 3^b 5^c sizes), one butterfly per warp step, twiddles W_Ls^(m k) from the plan's W_L table.
//            In place means no ping-pong buffer (the tile is the whole shared memory of an SM for L = 800) and no values held
//            across a barrier; the price is a digit-reversed spectrum: bin k sits at pos(k) = (k % R1) L/R1 + ...
//   post     bins k and L - k from Z[pos(k)], Z[pos(L-k)] and W_N^k; |X|^2 (or X) is held in registers across one barrier
//            and then written in NATURAL order over the (now dead) spectrum tile: P[bin][frame]
//   epilogue the lane = frame epilogue (epilogue.cuh): any mapping, scaling, fused DCT; 128-byte row stores
#include "epilogue.cuh"
#include "fast400_common.cuh"
#include "launch.hpp"
#include "mixed_dft.cuh"

namespace sgx {
namespace {

using mx::Cx;

constexpr int kFrames = 32;          // frames per tile = lanes
constexpr int kRow = 33;             // complex elements between consecutive positions of the tile

template <typename T> __device__ __forceinline__ Cx<T> ldg_cx(const Cx<T> *p);
template <> __device__ __forceinline__ Cx<float> ldg_cx<float>(const Cx<float> *p) {
    const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    return {v.x, v.y};
}
template <> __device__ __forceinline__ Cx<double> ldg_cx<double>(const Cx<double> *p) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return {v.x, v.y};
}

// position of bin k (k < L) after the in-place decimation-in-frequency stages
template <int R1, int R2, int R3> __host__ __device__ constexpr int pos_of(int k) {
    constexpr int L = R1 * R2 * R3;
    return (k % R1) * (L / R1) + ((k / R1) % R2) * R3 + k / (R1 * R2);
}

// One in-place stage on blocks of length LS: butterfly b = (blk, m), elements blk LS + m + j LS/R, outputs times W_LS^(m k)
template <typename T, int L, int LS, int R, int W>
__device__ __forceinline__ void stage(Cx<T> *zf, const Cx<T> *__restrict__ tw, int warp) {
    constexpr int M = LS / R;
#pragma unroll 1
    for (int b = warp; b < L / R; b += W) {
        const int blk = b / M, m = b - blk * M;
        Cx<T> *zb = zf + (blk * LS + m) * kRow;
        Cx<T> v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = zb[j * M * kRow];
        mx::Dft<T, R>::run(v);
        if (M > 1 && m > 0) {
            // W_LS^(m k) = W_L^(m k L / LS); m k < LS, so the table index needs no reduction
            const Cx<T> *twm = tw + m * (L / LS);
#pragma unroll
            for (int k = 1; k < R; ++k) v[k] = v[k] * ldg_cx<T>(twm + (k - 1) * m * (L / LS));
        }
#pragma unroll
        for (int k = 0; k < R; ++k) zb[k * M * kRow] = v[k];
    }
}

template <typename T, int R1, int R2, int R3, int W>
__global__ void __launch_bounds__(32 * W, (32 * W <= 320 && sizeof(T) == 4) ? 2 : 1) k_r2c_fused_mixed(const __grid_constant__ KParams p) {
    constexpr int L = R1 * R2 * R3, N = 2 * L;
    using C = Cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *z = reinterpret_cast<C *>(smem_raw);                     // [(L + 1) * kRow]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const long long f0 = p.frame_begin + static_cast<long long>(tile) * kFrames;
    const long long rem = p.frame_begin + p.frames_todo - f0;
    const int nf = rem < kFrames ? static_cast<int>(rem) : kFrames;
    const T *x = static_cast<const T *>(p.samples) + static_cast<long long>(clip) * p.clip_stride;
    const T *win = static_cast<const T *>(p.window);
    // Quad epilogue (f32 mel / loghz spectrogram plans; the sparse-row schedule of the n400 family, fast400_common.cuh): its table --
    // int4 {byte offset of P[c0], cnt, weight byte offset, row}[4 n_quads] + zero-padded weights, built by the host -- starts its
    // way into the shared memory behind the tile now (cp.async, nothing held) and is complete long before the epilogue.
    const bool quads = sizeof(T) == 4 && p.lane_w_smem != 0;
    if (quads) {
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw + p.lane_w_smem));
        const char *src = static_cast<const char *>(p.sched);
        for (int o = 16 * static_cast<int>(threadIdx.x); o < p.lane_w_bytes; o += 16 * 32 * W)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- load + window: warp = frame (strided), lanes along the packed elements n (samples 2n, 2n+1). The window pairs of a
    //      lane's elements are loaded once; interior frames are taken two at a time up to n_fft 512 (one at a time above: two
    //      would spill), every load issued before the first use. (Staging the raw samples with cp.async and applying the window in the first stage was measured
    //      slower: 2.48 against 2.15 ms on 800 / 200.)
    {
        constexpr int NI = (L + 31) / 32;
        auto has = [&](int i) { return 32 * i + 31 < L || lane + 32 * i < L; };
        const bool vec = (p.vec_ok & 1) != 0;
        C wv[NI];
        if (vec) {
            const C *wf = reinterpret_cast<const C *>(win) + lane;
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (has(i)) wv[i] = ldg_cx<T>(wf + 32 * i);
        }
        auto interior = [&](int f) {
            const long long base = (f0 + f) * p.hop - p.pad;
            return vec && f < nf && base >= 0 && base + N <= p.n_samples;
        };
        auto issue = [&](int f, C (&sv)[NI]) {
            const C *xf = reinterpret_cast<const C *>(x + ((f0 + f) * p.hop - p.pad)) + lane;
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (has(i)) sv[i] = ldg_cx<T>(xf + 32 * i);
        };
        auto finish = [&](int f, const C (&sv)[NI]) {
            C *zf = z + f + lane * kRow;
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (has(i)) zf[32 * i * kRow] = C{sv[i].x * wv[i].x, sv[i].y * wv[i].y};      // sample * window[i] (src/spectrogram.rs:1319)
        };
        auto slow = [&](int f) {
            const long long base = (f0 + f) * p.hop - p.pad;
            C *zf = z + f;
            if (f >= nf) {                                      // frames beyond the clip's last: zeros (never stored)
                for (int n = lane; n < L; n += 32) zf[n * kRow] = C{T(0), T(0)};
                return;
            }
            for (int n = lane; n < L; n += 32) {
                const long long s0 = base + 2 * n;
                const T a = (s0 >= 0 && s0 < p.n_samples) ? __ldg(x + s0) : T(0);
                const T b = (s0 + 1 >= 0 && s0 + 1 < p.n_samples) ? __ldg(x + s0 + 1) : T(0);
                zf[n * kRow] = C{a * __ldg(win + 2 * n), b * __ldg(win + 2 * n + 1)};
            }
        };
        if constexpr (NI <= 8) {
#pragma unroll 1
            for (int f = warp; f < kFrames; f += 2 * W) {
                const int g = f + W;
                const bool fa = interior(f), fb = g < kFrames && interior(g);
                C sa[NI], sb[NI];
                if (fa) issue(f, sa);
                if (fb) issue(g, sb);
                if (fa) finish(f, sa); else slow(f);
                if (fb) finish(g, sb); else if (g < kFrames) slow(g);
            }
        } else {                                                // larger frames: one at a time (two would spill)
#pragma unroll 1
            for (int f = warp; f < kFrames; f += W) {
                if (interior(f)) {
                    C sa[NI];
                    issue(f, sa);
                    finish(f, sa);
                } else {
                    slow(f);
                }
            }
        }
    }
    __syncthreads();

    // ---- stages (lane = frame from here on)
    C *zl = z + lane;
    const C *tw = static_cast<const C *>(p.tw);
    stage<T, L, L, R1, W>(zl, tw, warp);
    __syncthreads();
    if constexpr (R2 > 1) {
        stage<T, L, L / R1, R2, W>(zl, tw, warp);
        __syncthreads();
    }
    if constexpr (R3 > 1) {
        stage<T, L, L / (R1 * R2), R3, W>(zl, tw, warp);
        __syncthreads();
    }

    // ---- post pass: X[k] = E + W_N^k O, X[L-k] = conj(E - W_N^k O) from Z[pos(k)], Z[pos(L-k)]. Warp w owns the pairs
    //      k = w, w + W, ... <= L/2; the results wait in registers for the barrier after which the spectrum tile is dead.
    const C *post = static_cast<const C *>(p.post);
    const bool want_complex = p.output == SGX_OUT_COMPLEX_STFT;
    constexpr int KPW = (L / 2 + 1 + W - 1) / W;                 // pairs per warp
    auto pair_of = [&](int k, C &xa, C &xb) {
        if (k == 0) {
            const C z0 = zl[0];
            xa = C{z0.x + z0.y, T(0)};                          // bin 0
            xb = C{z0.x - z0.y, T(0)};                          // bin L
        } else {
            const C a = zl[pos_of<R1, R2, R3>(k) * kRow], b = zl[pos_of<R1, R2, R3>(L - k) * kRow];
            const C ev = {T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y)};
            const C od = {T(0.5) * (a.y + b.y), T(0.5) * (b.x - a.x)};
            const C wo = od * ldg_cx<T>(post + k);
            xa = ev + wo;                                       // bin k
            const C d = ev - wo;
            xb = C{d.x, -d.y};                                  // bin L - k
        }
    };
    if (want_complex) {
        // StftPlan::compute column copy (src/spectrogram.rs:1440-1442): the spectrum goes straight from registers to one run
        // of frames per bin row
        using OC = typename Cplx<T>::type;
        OC *out = static_cast<OC *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin) + lane;
#pragma unroll 1
        for (int k = warp; k <= L / 2; k += W) {
            C xa, xb;
            pair_of(k, xa, xb);
            if (lane < nf) {
                out[static_cast<long long>(k) * p.out_row_stride] = mk<T>(xa.x, xa.y);
                out[static_cast<long long>(L - k) * p.out_row_stride] = mk<T>(xb.x, xb.y);
            }
        }
        return;
    }
    T pa[KPW], pb[KPW];
#pragma unroll
    for (int i = 0; i < KPW; ++i) {
        const int k = warp + W * i;
        if (k <= L / 2) {
            C xa, xb;
            pair_of(k, xa, xb);
            pa[i] = xa.x * xa.x + xa.y * xa.y;                  // norm_sqr (src/spectrogram.rs:1332-1334)
            pb[i] = xb.x * xb.x + xb.y * xb.y;
        }
    }
    __syncthreads();
    T *P = reinterpret_cast<T *>(smem_raw);                     // P[bin][32 frames], then the scratch rows of the fused MFCC
    // quad epilogue: frames permuted inside a row so that one 16-byte read returns frames j, j+8, j+16, j+24 (f400::frame_col)
    const int pcol = quads ? f400::frame_col(lane) : lane;
#pragma unroll
    for (int i = 0; i < KPW; ++i) {
        const int k = warp + W * i;
        if (k <= L / 2) {
            P[k * 32 + pcol] = pa[i];
            P[(L - k) * 32 + pcol] = pb[i];
        }
    }
    if (quads) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if constexpr (sizeof(T) == 4) {
        if (quads) {
            // 4 rows x 32 frames per warp step, rows dealt to the warps quad by quad (longest first); ascending columns, one fused
            // rounding per term (never further from the exact sum than SparseMatrix::multiply_vec, src/spectrogram.rs:102-117)
            const int4 *tab = reinterpret_cast<const int4 *>(smem_raw + p.lane_w_smem);
            const int nq = tab[0].x;
            const unsigned wrel = static_cast<unsigned>(__cvta_generic_to_shared(tab + 1 + 4 * nq));
            float *ocf = reinterpret_cast<float *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
            const float *Pf = reinterpret_cast<const float *>(P);
            if (p.apply_db) f400::sparse_quads_epilogue<2, false>(p, Pf, tab + 1, warp, nq, W, ocf, nullptr, nf, lane, wrel);
            else if (p.amp == SGX_AMP_MAGNITUDE) f400::sparse_quads_epilogue<1, false>(p, Pf, tab + 1, warp, nq, W, ocf, nullptr, nf, lane, wrel);
            else f400::sparse_quads_epilogue<0, false>(p, Pf, tab + 1, warp, nq, W, ocf, nullptr, nf, lane, wrel);
            return;
        }
    }
    epilogue_lane_frames<T>(p, P, P + (L + 1) * 32, clip, f0, nf, lane);
}

template <typename T, int R1, int R2, int R3, int W>
cudaError_t launch_one(const KParams &p, cudaStream_t stream) {
    constexpr int L = R1 * R2 * R3;
    size_t smem = sizeof(Cx<T>) * static_cast<size_t>(L + 1) * kRow;
    if (p.lane_w_smem) smem = static_cast<size_t>(p.lane_w_smem) + static_cast<size_t>(p.lane_w_bytes);     // + the quad epilogue's table
    const long long grid = static_cast<long long>(p.n_clips) * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_mixed<T, R1, R2, R3, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_mixed<T, R1, R2, R3, W><<<static_cast<unsigned>(grid), 32 * W, smem, stream>>>(p);
    return cudaGetLastError();
}

// the sizes with a compiled instance: n_fft -> (R1, R2, R3, warps). The warp count divides (or nearly divides) the butterfly
// counts L / R of the stages, and stays at <= 10 warps up to n_fft 800 so that two CTAs share an SM.
#define SGX_MIXED_SIZES(X) \
    X(64, 8, 4, 1, 4) X(128, 8, 8, 1, 8) X(160, 10, 8, 1, 8) X(200, 10, 10, 1, 10) X(240, 12, 10, 1, 10) X(320, 16, 10, 1, 10) \
    X(400, 20, 10, 1, 10) X(480, 16, 15, 1, 8) X(500, 25, 10, 1, 10) X(600, 20, 15, 1, 10) X(640, 20, 16, 1, 10) X(800, 20, 20, 1, 10) \
    X(960, 24, 20, 1, 12) X(1000, 25, 20, 1, 10) X(1200, 25, 24, 1, 12) X(1600, 10, 10, 8, 20)

}  // namespace

bool mixed_supported(size_t n_fft, bool f64) {
#define X(NFFT, A, B, C, WARPS) if (n_fft == NFFT) return !f64 || NFFT <= 800;
    SGX_MIXED_SIZES(X)
#undef X
    return false;
}
size_t mixed_smem_bytes(size_t n_fft, bool f64) { return (f64 ? 16 : 8) * (n_fft / 2 + 1) * static_cast<size_t>(kRow); }
int mixed_tile_frames() { return kFrames; }
// rows of the log-mel tile the fused MFCC can park behind the power tile (both inside the dead spectrum tile)
int mixed_max_scratch_rows(size_t n_fft) { return static_cast<int>((n_fft / 2 + 1) * (2 * kRow - 32) / 32); }

cudaError_t launch_mixed(const KParams &p, bool f64, cudaStream_t stream) {
#define X(NFFT, A, B, C, WARPS)                                                              \
    if (p.n_fft == NFFT) {                                                                   \
        if (!f64) return launch_one<float, A, B, C, WARPS>(p, stream);                       \
        if constexpr (NFFT <= 800) return launch_one<double, A, B, C, WARPS>(p, stream);     \
    }
    SGX_MIXED_SIZES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sgx
