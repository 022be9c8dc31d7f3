// kernel_pow2.cu -- "r2c_fused_pow2": power-of-two n_fft (256 .. 8192), f32 and f64 (BASELINE configs 0, 2 and 4).
//
// One CTA = FT consecutive frames of one clip; a frame's packed M = n_fft/2 complex FFT is shared by TPF = M/16
// threads, each holding 16 complex values in registers:
//
//   load      samples straight from global memory (coalesced 8/16-byte pairs, zero outside the clip = centre padding,
//             src/spectrogram.rs:1309-1320) times the window, packed as z[n] = x[2n] + i x[2n+1]
//   passes    Stockham autosort with register radix-16 butterflies (one per thread and pass) and a last pass of radix
//             M / 16^p in {1, 2, 4, 8} (16/r butterflies per thread). The exchange between passes goes through ONE
//             shared buffer per frame (read everything - barrier - write everything), padded by one element per 16 so
//             that both the strided writes of the early passes and the contiguous reads are conflict free.
//   post      split of the packed spectrum (realfft-style): one twiddle multiply serves bins k and M-k
//   epilogue  |X|^2 (or X) tile -> mapping -> scaling -> (DCT) -> store            (epilogue.cuh)
//
// Compared with the generic family this cuts the shared-memory round trips from log4(M) to 2-3 and removes every
// runtime division from the index math (all radices, strides and shifts are template constants).
#include "binaural.cuh"
#include "epilogue.cuh"
#include <type_traits>

#include "launch.hpp"
#include "tcgen05.cuh"

namespace sgx {
namespace {

template <typename T> struct Cx { T x, y; };
// f32: Blackwell packed FP32x2 (FADD2 / FMUL2 / FFMA2) -- a complex add is one issue slot and the operand swizzle of the
// packed forms absorbs the (y, -x) rotations; f64 stays scalar.
__device__ __forceinline__ unsigned long long pk2(Cx<float> v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ Cx<float> upk2(unsigned long long v) { Cx<float> r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ Cx<float> operator+(Cx<float> a, Cx<float> b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> operator-(Cx<float> a, Cx<float> b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> mul2(Cx<float> a, Cx<float> b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ Cx<float> fma2(Cx<float> a, Cx<float> b, Cx<float> c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c))); return upk2(r); }
// a * b = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
__device__ __forceinline__ Cx<float> operator*(Cx<float> a, Cx<float> b) { return fma2(Cx<float>{a.y, a.y}, Cx<float>{-b.y, b.x}, mul2(Cx<float>{a.x, a.x}, b)); }
__device__ __forceinline__ Cx<double> mul2(Cx<double> a, Cx<double> b) { return {a.x * b.x, a.y * b.y}; }
__device__ __forceinline__ Cx<double> operator+(Cx<double> a, Cx<double> b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ Cx<double> operator-(Cx<double> a, Cx<double> b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ Cx<double> operator*(Cx<double> a, Cx<double> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T> __device__ __forceinline__ Cx<T> mul_mi(Cx<T> a) { return {a.y, -a.x}; }   // * (-i)

template <typename T> __device__ __forceinline__ void dft2(Cx<T> &a, Cx<T> &b) { const Cx<T> t = a - b; a = a + b; b = t; }
template <typename T> __device__ __forceinline__ void dft4(Cx<T> &a, Cx<T> &b, Cx<T> &c, Cx<T> &d) {
    const Cx<T> t0 = a + c, t1 = a - c, t2 = b + d, t3 = mul_mi(b - d);
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

// forward DFT of v[0..R-1] in natural order, in place
template <typename T, int R> struct Dft;
template <typename T> struct Dft<T, 1> { static __device__ __forceinline__ void run(Cx<T> *) {} };
template <typename T> struct Dft<T, 2> { static __device__ __forceinline__ void run(Cx<T> *v) { dft2(v[0], v[1]); } };
template <typename T> struct Dft<T, 4> { static __device__ __forceinline__ void run(Cx<T> *v) { dft4(v[0], v[1], v[2], v[3]); } };
template <typename T> struct Dft<T, 8> {
    static __device__ __forceinline__ void run(Cx<T> *v) {
        // n = j + 2 n1: A_j = DFT4 over n1; X[k1] = A0[k1] + W8^k1 A1[k1], X[k1+4] = A0[k1] - W8^k1 A1[k1]
        Cx<T> a0 = v[0], a1 = v[2], a2 = v[4], a3 = v[6], b0 = v[1], b1 = v[3], b2 = v[5], b3 = v[7];
        dft4(a0, a1, a2, a3);
        dft4(b0, b1, b2, b3);
        const T h = T(0.70710678118654752440084436210485);
        b1 = {h * (b1.x + b1.y), h * (b1.y - b1.x)};     // * W8^1
        b2 = mul_mi(b2);                                  // * W8^2
        b3 = {h * (b3.y - b3.x), -h * (b3.x + b3.y)};    // * W8^3
        v[0] = a0 + b0; v[4] = a0 - b0; v[1] = a1 + b1; v[5] = a1 - b1;
        v[2] = a2 + b2; v[6] = a2 - b2; v[3] = a3 + b3; v[7] = a3 - b3;
    }
};
template <typename T> struct Dft<T, 16> {
    static __device__ __forceinline__ void run(Cx<T> *v) {
        // n = j + 4 n1: A_j[k1] = DFT4 over n1 of v[j + 4 n1]; X[k1 + 4 k2] = DFT4 over j of W16^(j k1) A_j[k1]
        const T c = T(0.92387953251128675612818318939679), s = T(0.38268343236508977172845998403040);
        const T h = T(0.70710678118654752440084436210485);
        Cx<T> a[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j][0] = v[j]; a[j][1] = v[j + 4]; a[j][2] = v[j + 8]; a[j][3] = v[j + 12];
            dft4(a[j][0], a[j][1], a[j][2], a[j][3]);
        }
        const Cx<T> w1 = {c, -s}, w2 = {h, -h}, w3 = {s, -c}, w6 = {-h, -h}, w9 = {-c, s};
        a[1][1] = a[1][1] * w1; a[1][2] = a[1][2] * w2; a[1][3] = a[1][3] * w3;
        a[2][1] = a[2][1] * w2; a[2][2] = mul_mi(a[2][2]); a[2][3] = a[2][3] * w6;
        a[3][1] = a[3][1] * w3; a[3][2] = a[3][2] * w6; a[3][3] = a[3][3] * w9;
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            dft4(a[0][k1], a[1][k1], a[2][k1], a[3][k1]);
            v[k1] = a[0][k1]; v[k1 + 4] = a[1][k1]; v[k1 + 8] = a[2][k1]; v[k1 + 12] = a[3][k1];
        }
    }
};

template <typename T> struct Dft<T, 32> {
    static __device__ __forceinline__ void run(Cx<T> *v) {
        // n = j + 4 n1: A_j[k1] = DFT8 over n1 of v[j + 4 n1]; X[k1 + 8 k2] = DFT4 over j of W32^(j k1) A_j[k1]
        Cx<T> a[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) a[j][n1] = v[j + 4 * n1];
            Dft<T, 8>::run(a[j]);
        }
        // W32^m = (cos(2 pi m / 32), -sin(2 pi m / 32)), m = j k1 <= 21
        constexpr double kC[8] = {1.0, 0.98078528040323044912618223613424, 0.92387953251128675612818318939679, 0.83146961230254523707878837761791,
                                  0.70710678118654752440084436210485, 0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                                  0.19509032201612826784828486847702};
#pragma unroll
        for (int j = 1; j < 4; ++j)
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) {
                const int m = j * k1;                        // angle m * pi / 16, quadrant m / 8
                const int r = m & 7, qd = m >> 3;
                const T c = static_cast<T>(kC[r]), sn = static_cast<T>(r == 0 ? 0.0 : kC[8 - r]);
                // (cos, -sin) of the first-quadrant angle, rotated by -i per quadrant
                Cx<T> w = {c, -sn};
                if (qd == 1) w = {-sn, -c};
                else if (qd == 2) w = {-c, sn};
                if (r == 0 && qd == 1) a[j][k1] = mul_mi(a[j][k1]);
                else a[j][k1] = a[j][k1] * w;
            }
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            dft4(a[0][k1], a[1][k1], a[2][k1], a[3][k1]);
            v[k1] = a[0][k1]; v[k1 + 8] = a[1][k1]; v[k1 + 16] = a[2][k1]; v[k1 + 24] = a[3][k1];
        }
    }
};

template <typename T> __device__ __forceinline__ Cx<T> ldg_cx(const Cx<T> *p);
template <> __device__ __forceinline__ Cx<float> ldg_cx<float>(const Cx<float> *p) {
    const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    return {v.x, v.y};
}
template <> __device__ __forceinline__ Cx<double> ldg_cx<double>(const Cx<double> *p) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return {v.x, v.y};
}

__host__ __device__ constexpr int pad16(int i) { return i + (i >> 4); }

// One Stockham pass of radix R with accumulated length CUR on a frame's M-point buffer; the thread owns butterflies
// b = t + TPF*u (u < 16/R). v[16] is the thread's register file for the whole pass: load - (CTA barrier) - store.
template <typename T, int M, int R, int CUR, int RPT = 16>
__device__ __forceinline__ void pass_load(const Cx<T> *z, const Cx<T> *__restrict__ tw, int t, Cx<T> *v) {
    constexpr int TPF = M / RPT, B = M / R, MM = M / (CUR * R);
#pragma unroll
    for (int u = 0; u < RPT / R; ++u) {
        const int b = t + TPF * u;
        const int q = b & (CUR - 1);
        // twiddles W_{CUR*R}^{j q} = w^j with w = W_M^{q MM}: two table loads (w, w^4); the other powers are products of
        // one of {w, w^2, w^3} and one of {w^4, w^8, w^12}, formed on the fly (at most three chained multiplications) --
        // the memory-instruction queue, not the FP pipe, is what these passes run out of
        Cx<T> wl[4], wh[4], w16 = {T(1), T(0)};
        if (CUR > 1) {
            wl[1] = ldg_cx<T>(tw + q * MM);
            if (R > 2) { wl[2] = wl[1] * wl[1]; wl[3] = wl[2] * wl[1]; }
            if (R > 4) wh[1] = ldg_cx<T>(tw + 4 * q * MM);
            if (R > 8) { wh[2] = wh[1] * wh[1]; wh[3] = wh[2] * wh[1]; }
            if (R > 16) w16 = ldg_cx<T>(tw + 16 * q * MM);        // radix 32: j = 16 + j' costs one more multiplication
        }
        // B is a multiple of 16, so pad16(j*B + b) = j*pad16(B) + pad16(b): one base address, constant strides
        const Cx<T> *zb = z + pad16(b);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            Cx<T> x = zb[j * pad16(B)];
            if (CUR > 1 && j > 0) {
                const int hi = (j >> 2) & 3, lo = j & 3;
                if ((j & 15) != 0) {
                    const Cx<T> wj = hi == 0 ? wl[lo] : (lo == 0 ? wh[hi] : wh[hi] * wl[lo]);
                    x = x * wj;
                }
                if (j >= 16) x = x * w16;
            }
            v[u * R + j] = x;
        }
    }
}
template <typename T, int M, int R, int CUR, int RPT = 16>
__device__ __forceinline__ void pass_store(Cx<T> *z, int t, Cx<T> *v) {
    constexpr int TPF = M / RPT;
#pragma unroll
    for (int u = 0; u < RPT / R; ++u) {
        const int b = t + TPF * u;
        const int q = b & (CUR - 1), i = b / CUR;
        Dft<T, R>::run(v + u * R);
        if (CUR >= 16) {
            // k*CUR is a multiple of 16: pad16((i*R + k)*CUR + q) = pad16(i*R*CUR + q) + k*pad16(CUR)
            Cx<T> *zo = z + pad16(i * R * CUR + q);
#pragma unroll
            for (int k = 0; k < R; ++k) zo[k * pad16(CUR)] = v[u * R + k];
        } else {
            // first pass (CUR = 1, R = 16 or 32): outputs R b + k; R b is a multiple of 16, so the pad of k adds on
            Cx<T> *zo = z + pad16(i * R * CUR + q);
#pragma unroll
            for (int k = 0; k < R; ++k) zo[pad16(k * CUR)] = v[u * R + k];
        }
    }
}

// Thread -> (frame, thread-in-frame) map. Small frames (TPF <= 16) keep their threads together inside a warp and need only
// warp-level syncs between passes. Frames of 32 threads and more (n_fft >= 1024, FT <= 8) are INTERLEAVED: every warp holds
// LPF = 32 / FT consecutive threads of each of the tile's FT frames, so that the transposed tile writes P[bin][FT] of a
// warp cover 32 consecutive words (they were FT-way bank conflicts with one frame per warp), the window load of a warp
// is one line shared by its FT frames, and the frame-strided FFT accesses stay conflict free because the frame stride
// is LPF mod 16 complex elements (zs_of). The price is a CTA-wide barrier between passes instead of a per-frame one.
template <int TPF, int FT, int RPT = 16> struct ThreadMap {
    // the two-pass form (RPT = 32, M = 1024) has exactly one warp per frame: keeping the warp together makes every exchange
    // of the FFT a warp-level sync, and its epilogue works on a frame-major tile, so it is not interleaved
    static constexpr bool kInterleaved = RPT == 16 && TPF >= 32 && FT > 1;
    static constexpr int LPF = kInterleaved ? 32 / FT : TPF;        // lanes of one frame inside a warp
    static __device__ __forceinline__ void get(int tid, int &fl, int &t) {
        if (kInterleaved) {
            const int lane = tid & 31, w = tid >> 5;
            fl = lane / LPF;
            t = (lane % LPF) + LPF * w;
        } else {
            fl = tid / TPF;
            t = tid - fl * TPF;
        }
    }
};
template <int TPF, int FT, int RPT = 16> __device__ __forceinline__ void frame_sync(int) {
    if (ThreadMap<TPF, FT, RPT>::kInterleaved || TPF > 32) __syncthreads();
    else __syncwarp();
}

// complex elements per frame buffer: >= M + 1 bins; for the interleaved map the stride is LPF mod 16 (M >= 256 makes
// pad16(M) a multiple of 16), which spreads the FT frame segments a warp touches over all banks
constexpr int zs_of(int M, int FT, int RPT = 16) {
    const int tpf = M / RPT;
    const bool inter = tpf >= 32 && FT > 1;        // (RPT = 32: the post pass and the epilogue use the interleaved map)
    return pad16(M) + (inter ? 32 / FT : 8);
}

template <int M> struct Radices {          // M = 16^P16 * LAST, LAST in {1, 2, 4, 8}
    static constexpr int P16 = M >= 4096 ? 3 : (M >= 256 ? 2 : 1);
    static constexpr int LAST = M / (P16 == 3 ? 4096 : (P16 == 2 ? 256 : 16));
};
// Complex values a thread holds. 32 (f32, M = 512 / 1024: n_fft 1024 / 2048) makes the FFT two passes -- radix 32 then radix
// M / 32 -- with ONE exchange instead of two: the L1 / shared-memory data pipe is what these sizes run out of (94 % busy on
// BASELINE configs[2] with three passes). f64 and the other sizes keep 16 (64 f64 registers of data per thread are too many).
template <typename T> __device__ __forceinline__ T lds_t(unsigned a);
template <> __device__ __forceinline__ float lds_t<float>(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <> __device__ __forceinline__ double lds_t<double>(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
template <typename T> constexpr int rpt_of(int M) { return (sizeof(T) == 4 && (M == 512 || M == 1024)) ? 32 : 16; }

template <typename T, int M, int FT, bool PAIR>
__global__ void __launch_bounds__(FT *(M / rpt_of<T>(M)), rpt_of<T>(M) == 32 ? 512 / (FT * (M / 32)) : (sizeof(T) == 4 ? 4 : 2) * 256 / (FT * (M / 16) > 256 ? FT * (M / 16) : 256))
k_r2c_fused_pow2(const __grid_constant__ KParams p) {
    constexpr int RPT = rpt_of<T>(M);                    // complex values per thread: 16, or 32 for the two-pass sizes
    constexpr int TPF = M / RPT, N = 2 * M;
    constexpr int ZS = zs_of(M, FT, RPT);                // complex elements per frame buffer (M+1 spectrum bins fit too)
    using C = Cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *zbuf = reinterpret_cast<C *>(smem_raw);

    const int tid = threadIdx.x;
    int fl, t;                                       // frame within the tile, thread within the frame
    ThreadMap<TPF, FT, RPT>::get(tid, fl, t);
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    // Ordinary launch: the tile is FT consecutive frames of one clip. Stereo-pair launch (samples_b != null, FT >= 2): FT/2
    // frames of the left channel in slots 0 .. FT/2-1 and the SAME frames of the right channel in the upper half, so that
    // the cue epilogue finds both spectra of a frame in the tile.
    constexpr int HF = FT >= 2 ? FT / 2 : 1;
    constexpr bool pair = PAIR && FT >= 2;           // a separate instantiation: the ordinary kernels carry none of this
    const int tf = pair ? HF : FT;                   // time frames per tile
    const int ch = pair ? fl / HF : 0;
    const int fi = fl - ch * HF * (pair ? 1 : 0);    // time-frame slot of this thread's frame
    const long long f0 = p.frame_begin + static_cast<long long>(tile) * tf;
    const long long rem = p.frame_begin + p.frames_todo - f0;
    const int nf = rem < tf ? static_cast<int>(rem) : tf;
    C *z = zbuf + fl * ZS;

    const T *x = static_cast<const T *>(ch ? p.samples_b : p.samples) + static_cast<long long>(clip) * p.clip_stride;
    const T *win = static_cast<const T *>(p.window);
    const C *tw = static_cast<const C *>(p.tw);
    C v[RPT];

    // The CTA that will take this one's place on the SM (about blockIdx.x + resident CTAs, in launch order) finds its samples in L2:
    // one 128-byte line per thread, no register or scoreboard held. The first global load of a CTA is otherwise an exposed HBM
    // round trip with 16 warps per SM.
    if constexpr (!PAIR) {
        if (p.l2_ahead) {
            const long long b2 = static_cast<long long>(blockIdx.x) + p.l2_ahead;
            const long long clip2 = b2 / p.tiles_per_clip;
            if (clip2 < p.n_clips) {
                constexpr int EPL = 128 / static_cast<int>(sizeof(T));       // elements per 128-byte line
                const long long s0 = (p.frame_begin + (b2 - clip2 * p.tiles_per_clip) * FT) * p.hop - p.pad;
                const T *x2 = static_cast<const T *>(p.samples) + clip2 * p.clip_stride;
                for (int o = EPL * tid; o < (FT - 1) * p.hop + N + EPL; o += EPL * FT * TPF) {
                    const long long s2 = s0 + o;
                    if (s2 >= 0 && s2 < p.n_samples) asm volatile("prefetch.global.L2 [%0];" ::"l"(x2 + s2));
                }
            }
        }
    }
    // Rows-per-thread epilogue (small tiles, sparse mapping): its lane-major weights start their way into shared memory now
    // (cp.async, no registers held) and are read from there after the FFT -- from global memory every batch of weight loads
    // was an exposed L2 round trip with 16 warps per SM (10 % of the music shard's stall samples).
    if constexpr (!PAIR) {
        if (p.lane_w_smem) {
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw + p.lane_w_smem));
            const char *src = static_cast<const char *>(p.lane_w);
            for (int o = 16 * tid; o < p.lane_w_bytes; o += 16 * FT * TPF)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }

    // ---- load + window + first pass (CUR = 1: no twiddles). Element n of the packed frame = samples 2n, 2n+1.
    {
        constexpr int R = RPT, B = M / R;
        const long long base = (f0 + (pair ? fi : fl)) * p.hop - p.pad;
        // Optional bulk-copy staging (p.vec_ok & 2; f32 two-pass form only): the tile's (FT - 1) hop + N contiguous samples are
        // brought into shared memory by ONE cp.async.bulk (the TMA unit's 1-D form, SASS UBLKCP) tracked by an mbarrier, so each
        // sample crosses L2 -> SM once per CTA instead of once per overlapping frame. Edge tiles keep the register path.
        bool staged = false;
        if constexpr (RPT == 32 && sizeof(T) == 4 && !PAIR) {
            const long long base0 = f0 * p.hop - p.pad;
            const long long span = static_cast<long long>(FT - 1) * p.hop + N;
            T *sx = reinterpret_cast<T *>(zbuf + FT * ZS);
            uint64_t *bar = reinterpret_cast<uint64_t *>(sx + ((FT - 1) * 1024 + N));      // sized for hop <= 1024 by the host
            if ((p.vec_ok & 2) && nf == FT && base0 >= 0 && base0 + span <= p.n_samples &&
                ((reinterpret_cast<uintptr_t>(x + base0)) & 15) == 0) {                  // block-uniform condition
                staged = true;
                if (tid == 0) {
                    tc::mbar_init(bar, 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
                __syncthreads();
                if (tid == 0) {
                    const unsigned bytes = static_cast<unsigned>(span * sizeof(T));
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_addr(bar)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_addr(sx)),
                                 "l"(x + base0), "r"(bytes), "r"(tc::smem_addr(bar))
                                 : "memory");
                }
                tc::mbar_wait(bar, 0);
                const C *xf = reinterpret_cast<const C *>(sx + fl * p.hop) + t;
                const C *wf = reinterpret_cast<const C *>(win) + t;
#pragma unroll
                for (int j = 0; j < R; ++j) v[j] = xf[j * B];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const C w = ldg_cx<T>(wf + j * B);
                    v[j] = mul2(v[j], w);                  // sample * window[i] (src/spectrogram.rs:1319), both halves in one instruction
                }
            }
        }
        if (staged) {
        } else if ((p.vec_ok & 1) && base >= 0 && base + N <= p.n_samples) {
            // interior frame (all but the first / last few of a clip): no bounds logic
            const C *xf = reinterpret_cast<const C *>(x + base) + t;
            const C *wf = reinterpret_cast<const C *>(win) + t;
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = ldg_cx<T>(xf + j * B);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const C w = ldg_cx<T>(wf + j * B);
                v[j] = mul2(v[j], w);                  // sample * window[i] (src/spectrogram.rs:1319), both halves in one instruction
            }
        } else {
#pragma unroll                      // (static indices only: a rolled loop would push v[] into local memory)
            for (int j = 0; j < R; ++j) {
                const int n = j * B + t;
                const long long s0 = base + 2 * n;
                const T a = (s0 >= 0 && s0 < p.n_samples) ? __ldg(x + s0) : T(0);
                const T b = (s0 + 1 >= 0 && s0 + 1 < p.n_samples) ? __ldg(x + s0 + 1) : T(0);
                v[j] = {a * __ldg(win + 2 * n), b * __ldg(win + 2 * n + 1)};
            }
        }
        pass_store<T, M, RPT, 1, RPT>(z, t, v);
    }
    frame_sync<TPF, FT, RPT>(fl);
    if constexpr (RPT == 32) {
        // two-pass form: radix 32 done, one pass of radix M / 32 (32 or 16) left -- a single exchange
        constexpr int R2 = M / 32;
        pass_load<T, M, R2, 32, RPT>(z, tw, t, v);
        frame_sync<TPF, FT, RPT>(fl);
        pass_store<T, M, R2, 32, RPT>(z, t, v);
        frame_sync<TPF, FT, RPT>(fl);
    } else {
    if constexpr (Radices<M>::P16 >= 2) {
        pass_load<T, M, 16, 16>(z, tw, t, v);
        frame_sync<TPF, FT, RPT>(fl);
        pass_store<T, M, 16, 16>(z, t, v);
        frame_sync<TPF, FT, RPT>(fl);
    }
    if constexpr (Radices<M>::P16 >= 3) {
        pass_load<T, M, 16, 256>(z, tw, t, v);
        frame_sync<TPF, FT, RPT>(fl);
        pass_store<T, M, 16, 256>(z, t, v);
        frame_sync<TPF, FT, RPT>(fl);
    }
    if constexpr (Radices<M>::LAST > 1) {
        constexpr int CUR = M / Radices<M>::LAST;
        pass_load<T, M, Radices<M>::LAST, CUR>(z, tw, t, v);
        frame_sync<TPF, FT, RPT>(fl);
        pass_store<T, M, Radices<M>::LAST, CUR>(z, t, v);
        frame_sync<TPF, FT, RPT>(fl);
    }
    }

    if constexpr (RPT == 32 && TPF >= 32 && FT > 1) {
        // The FFT passes ran with one warp per frame (warp-level syncs only). The post pass only READS the frame buffers, so it
        // may use any thread -> (frame, bin) map: switch to the interleaved one (every warp holds 32 / FT consecutive threads
        // of each frame), which makes the transposed tile writes P[bin][FT] of the epilogue conflict free.
        __syncthreads();
        constexpr int LPF = 32 / FT;
        const int lane = tid & 31, w = tid >> 5;
        fl = lane / LPF;
        t = (lane % LPF) + LPF * w;
        z = zbuf + fl * ZS;
    }

    // ---- post pass: X[k] = E + W_N^k O, X[M-k] = conj(E - W_N^k O); thread owns k = t + TPF*u (u < RPT/2), plus k = M/2
    //      and k = 0 / M on thread 0. Values are held in registers across the barrier because the tile may alias z.
    const C *post = static_cast<const C *>(p.post);
    // k = t + TPF u and M - k walk the padded buffer with constant strides when TPF is a multiple of 16 (pad16 is linear on such
    // steps): two base addresses per thread, immediate offsets per u. The halvings of the split are taken once on the sums --
    // 0.5 (E' + W O') instead of 0.5 E' + W (0.5 O'): scaling by a power of two commutes with every rounding, the bits are the same.
    constexpr int PSTEP = TPF + TPF / 16;
    const int pa0 = pad16(t), pb0 = pad16(M - t);
    auto post_pair = [&](int u, C &xa, C &xb) {
        const int k = t + TPF * u;                 // 0 .. M/2 - 1
        if (k == 0) {
            const C z0 = z[0];
            xa = {z0.x + z0.y, T(0)};              // bin 0
            xb = {z0.x - z0.y, T(0)};              // bin M
        } else {
            const C a = TPF % 16 == 0 ? z[pa0 + PSTEP * u] : z[pad16(k)], b = TPF % 16 == 0 ? z[pb0 - PSTEP * u] : z[pad16(M - k)];
            const C ev = a + C{b.x, -b.y};         // 2 E
            const C od = C{a.y, -a.x} + C{b.y, b.x};   // 2 O
            const C wo = od * ldg_cx<T>(post + k);
            const C s = ev + wo, d = ev - wo;
            xa = mul2(C{T(0.5), T(0.5)}, s);       // bin k
            xb = mul2(C{T(0.5), T(-0.5)}, d);      // bin M - k
        }
    };
    if (p.output == SGX_OUT_COMPLEX_STFT) {
        C xa[RPT / 2], xb[RPT / 2];
#pragma unroll
        for (int u = 0; u < RPT / 2; ++u) post_pair(u, xa[u], xb[u]);
        C xm = {T(0), T(0)};
        if (t == 0) {                              // bin M/2: E and O are both Z[M/2]-derived, W_N^(M/2) = -i
            const C a = z[pad16(M / 2)];
            xm = {a.x, -a.y};
        }
        __syncthreads();
        typename Cplx<T>::type *S = reinterpret_cast<typename Cplx<T>::type *>(zbuf) + fl * p.frame_stride;
#pragma unroll
        for (int u = 0; u < RPT / 2; ++u) {
            const int k = t + TPF * u;
            S[k] = mk<T>(xa[u].x, xa[u].y);
            S[M - k] = mk<T>(xb[u].x, xb[u].y);
        }
        if (t == 0) S[M / 2] = mk<T>(xm.x, xm.y);
        __syncthreads();
        if (pair) {
            // compute_{itd,ipd,ild,ilr}_spectrogram (src/binaural.rs:472, :830, :1187, :1530) on the tile: (bin, frame) with
            // frames fastest, both channels' spectra read from shared memory, one run of HF frames stored per band row
            const typename Cplx<T>::type *S0 = reinterpret_cast<const typename Cplx<T>::type *>(zbuf);
            T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
            const T bw = static_cast<T>(p.cue_bin_width);
            for (int idx = tid; idx < p.cue_band * HF; idx += FT * TPF) {
                const int b = idx / HF, f = idx - b * HF;
                if (f >= nf) continue;
                const int k = p.cue_start_bin + b;
                const typename Cplx<T>::type l = S0[f * p.frame_stride + k], r = S0[(HF + f) * p.frame_stride + k];
                out[static_cast<long long>(b) * p.out_row_stride + f] = binaural_cue<T>(p.cue, l.x, l.y, r.x, r.y, k, bw, p.cue_power, p.cue_wrapped);
            }
            return;
        }
        epilogue_complex<T>(p, reinterpret_cast<typename Cplx<T>::type *>(zbuf), clip, f0, nf);
        return;
    }
    // every other output starts from the power spectrum (norm_sqr, src/spectrogram.rs:1332-1334): square before the
    // barrier, so only one value per bin stays live across it
    T pa[RPT / 2], pb[RPT / 2], pm = T(0);
#pragma unroll
    for (int u = 0; u < RPT / 2; ++u) {
        C xa, xb;
        post_pair(u, xa, xb);
        pa[u] = xa.x * xa.x + xa.y * xa.y;
        pb[u] = xb.x * xb.x + xb.y * xb.y;
    }
    if (t == 0) {
        const C a = z[pad16(M / 2)];
        pm = a.x * a.x + a.y * a.y;
    }
    __syncthreads();

    T *P = reinterpret_cast<T *>(zbuf);
    if (FT <= 8 && p.mapping == SGX_MAP_CHROMA && p.dense_t != nullptr) {
        // chromagram() (src/chroma.rs:487-503) on a small tile. Magnitude tile transposed to P[bin][FT]; thread (r, c) sums
        // pitch class r over bin chunk c for all FT frames (one 16-byte tile read per bin, shared by the 12 rows of a
        // chunk; the filterbank is passed transposed, [bin][12], so a chunk's 12 weights are one 48-byte run); the chunk
        // partials are then added in ascending chunk order. Only the bins inside [f_min, f_max] carry weight (:308-310);
        // the reference adds exact zeros for the others. Chunked summation is not the reference's strictly sequential
        // order -- it stays inside the f32 1e-5 / f64 1e-12 budgets (tests/test_chroma.py).
        constexpr int NT = FT * TPF, CH = NT / 12;
#pragma unroll
        for (int u = 0; u < RPT / 2; ++u) {
            const int k = t + TPF * u;
            P[k * FT + fl] = t_sqrt(pa[u]);
            P[(M - k) * FT + fl] = t_sqrt(pb[u]);
        }
        if (t == 0) P[(M / 2) * FT + fl] = t_sqrt(pm);
        __syncthreads();
        T *part = P + (M + 1) * FT;                          // [CH][12][FT] behind the tile (second tile region)
        const int r = tid % 12, c = tid / 12;
        if (c < CH) {
            const int span = p.dense_c1 - p.dense_c0, per = (span + CH - 1) / CH;
            const int k0 = p.dense_c0 + c * per, k1 = min(k0 + per, p.dense_c1);
            const T *w = static_cast<const T *>(p.dense_t) + r;
            T acc[FT];
#pragma unroll
            for (int f = 0; f < FT; ++f) acc[f] = T(0);
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                const T wk = __ldg(w + 12 * k);
                T x[FT];
#pragma unroll
                for (int f = 0; f < FT; ++f) x[f] = P[k * FT + f];
#pragma unroll
                for (int f = 0; f < FT; ++f) acc[f] = t_add_rn(acc[f], t_mul_rn(wk, x[f]));
            }
#pragma unroll
            for (int f = 0; f < FT; ++f) part[(c * 12 + r) * FT + f] = acc[f];
        }
        __syncthreads();
        if (tid < 12 * FT) {                                  // (row, frame): chunk partials in ascending order
            T sum = T(0);
            for (int cc = 0; cc < CH; ++cc) sum = t_add_rn(sum, part[cc * 12 * FT + tid]);
            P[tid] = sum;                                     // the tile is dead: chroma[r][f] at P[r * FT + f]
        }
        __syncthreads();
        if (tid < nf) {
            T cv[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) cv[i] = P[i * FT + tid];
            chroma_normalise<T>(cv, p.chroma_norm);
            T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin) + tid;
#pragma unroll
            for (int i = 0; i < 12; ++i) out[static_cast<long long>(i) * p.out_row_stride] = cv[i];
        }
        return;
    }
    const bool rows_per_thread = FT <= 8 && p.output == SGX_OUT_SPECTROGRAM && p.mapping != SGX_MAP_LINEAR && p.n_lane_slots > 0;
    if (rows_per_thread) {
        // Small tiles, sparse mapping: power tile transposed to P[bin][FT] so that one filterbank row = one thread reads all
        // FT frames of a column with a single vector load, and every weight / column index is loaded once per FT outputs.
        // Same ascending-column, un-fused arithmetic as SparseMatrix::multiply_vec (src/spectrogram.rs:102-117).
#pragma unroll
        for (int u = 0; u < RPT / 2; ++u) {
            const int k = t + TPF * u;
            P[k * FT + fl] = pa[u];
            P[(M - k) * FT + fl] = pb[u];
        }
        if (t == 0) P[(M / 2) * FT + fl] = pm;
        if (p.lane_w_smem) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const T eps = static_cast<T>(p.eps);
        T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin);
        // lane slots from the host-built schedule (sgx_api.cu, build_lane_rows): rows of similar length share a warp, the
        // weight of entry i of the warp's 32 rows is one coalesced line, the tile read is a 16-byte vector per column.
        // Two copies of the loop so that the weight reads are plain shared-memory loads in one and read-only global loads in the other.
        auto rows = [&](auto staged) {
            constexpr bool STAGED = decltype(staged)::value;
            const unsigned ws = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw)) + static_cast<unsigned>(p.lane_w_smem);
            const T *lw = static_cast<const T *>(p.lane_w);
            for (int slot = tid; slot < p.n_lane_slots; slot += FT * TPF) {
                const int4 d = __ldg(p.lane_rows + slot);            // {row, first column, count, weight block}
                const T *wl = lw + d.w + (slot & 31);
                const unsigned wsl = ws + static_cast<unsigned>(sizeof(T)) * static_cast<unsigned>(d.w + (slot & 31));
                const T *pc = P + d.y * FT;
                T acc[FT];
#pragma unroll
                for (int f = 0; f < FT; ++f) acc[f] = T(0);
#pragma unroll 4
                for (int i = 0; i < d.z; ++i) {
                    T w;
                    if constexpr (STAGED) w = lds_t<T>(wsl + static_cast<unsigned>(sizeof(T)) * 32u * static_cast<unsigned>(i));
                    else w = __ldg(wl + i * 32);
                    T x[FT];
#pragma unroll
                    for (int f = 0; f < FT; ++f) x[f] = pc[i * FT + f];
#pragma unroll
                    for (int f = 0; f < FT; ++f) acc[f] = t_add_rn(acc[f], t_mul_rn(w, x[f]));
                }
                if (d.x < 0) continue;
                T *orow = out + static_cast<long long>(d.x) * p.out_row_stride;
#pragma unroll
                for (int f = 0; f < FT; ++f)
                    if (f < nf) orow[f] = amp_scale<T>(acc[f], p.amp, p.apply_db, eps);
            }
        };
        if (!PAIR && p.lane_w_smem != 0) rows(std::true_type{});
        else rows(std::false_type{});
        return;
    }
    if (p.output == SGX_OUT_SPECTROGRAM && p.mapping == SGX_MAP_LINEAR) {
        // Identity mapping (FrequencyMapping::apply, Identity arm :1829-1844): park the power tile frame-major
        // (stride PS chosen so that both sides are conflict free), then store with lanes = (bin, frame), frame fastest:
        // every bin row receives its FT consecutive frames from adjacent lanes, i.e. one contiguous run per row.
        constexpr int PS = ((M + 1 + 31) & ~31) + ((32 / FT) / (int(sizeof(T)) / 4) > 0 ? (32 / FT) / (int(sizeof(T)) / 4) : 1);
        static_assert(PS <= 2 * ZS, "frame-major tile must fit the FFT buffer");
        const T eps = static_cast<T>(p.eps);
        T *pfm = P + fl * PS;
#pragma unroll
        for (int u = 0; u < RPT / 2; ++u) {
            const int k = t + TPF * u;
            pfm[k] = pa[u];
            pfm[M - k] = pb[u];
        }
        if (t == 0) pfm[M / 2] = pm;
        __syncthreads();
        // the scaling sits in the (rolled) store loop: one inlined sqrt / log instead of 17 copies -- the f64 square
        // root is long enough that the unrolled form ran out of instruction cache
        // A thread keeps its frame f = tid % FT and walks bins k = tid / FT, += NT / FT: pointer increments instead of a
        // 64-bit multiply per element, and the amplitude mode is resolved outside the loop.
        constexpr int NT = FT * TPF, KSTEP = NT / FT;
        const int f = tid % FT, k0 = tid / FT;
        if (f < nf) {
            T *o = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin) +
                   static_cast<long long>(k0) * p.out_row_stride + f;
            const long long ostep = static_cast<long long>(KSTEP) * p.out_row_stride;
            const T *src = P + f * PS + k0;
            if (p.amp == SGX_AMP_MAGNITUDE && !p.apply_db) {
#pragma unroll 2
                for (int k = k0; k <= M; k += KSTEP, o += ostep, src += KSTEP) *o = t_sqrt(*src);
            } else if (p.amp != SGX_AMP_MAGNITUDE && !p.apply_db) {
#pragma unroll 4
                for (int k = k0; k <= M; k += KSTEP, o += ostep, src += KSTEP) *o = *src;
            } else {
#pragma unroll 2
                for (int k = k0; k <= M; k += KSTEP, o += ostep, src += KSTEP) *o = amp_scale<T>(*src, p.amp, p.apply_db, eps);
            }
        }
        return;
    }
    T *pf = P + fl * p.tile_stride;
#pragma unroll
    for (int u = 0; u < RPT / 2; ++u) {
        const int k = t + TPF * u;
        pf[k] = pa[u];
        pf[M - k] = pb[u];
    }
    if (t == 0) pf[M / 2] = pm;
    if (!PAIR && p.lane_w_smem) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (FT >= 16 && p.output == SGX_OUT_SPECTROGRAM && p.rows_contig && (p.mapping == SGX_MAP_MEL || p.mapping == SGX_MAP_LOGHZ)) {
        // Wide tiles (n_fft 256 / 512), sparse rows: lane = frame, FT consecutive lanes share a row, so the row's extent
        // and every weight are one broadcast load and the store is a run of FT frames; all index math is compile time.
        // Ascending-column, un-fused accumulation as in SparseMatrix::multiply_vec (src/spectrogram.rs:102-117).
        // STAGED: the row descriptors and the weights were copied into shared memory by cp.async at kernel start (the block
        // p.lane_w: int4 {first entry, count, first column}[n_bins], then the CSR values): every broadcast read is then a
        // shared-memory read instead of an exposed global / L1 round trip (73 % of the 512 / 160 / 80 kernel's stall samples sat here).
        constexpr int NT = FT * TPF, GROUPS = NT / FT;
        const int f = tid % FT, g = tid / FT;
        const T eps = static_cast<T>(p.eps);
        const T *px = P + f * p.tile_stride;
        T *out = static_cast<T *>(p.out) + static_cast<long long>(clip) * p.out_clip_stride + (f0 - p.out_frame_origin) + f;
        auto rows = [&](auto staged) {
            constexpr bool STAGED = decltype(staged)::value;
            const int4 *sdesc = reinterpret_cast<const int4 *>(smem_raw + p.lane_w_smem);
            const T *val = STAGED ? reinterpret_cast<const T *>(sdesc + p.n_bins) : static_cast<const T *>(p.val);
            for (int row = g; row < p.n_bins; row += GROUPS) {
                const int4 d = STAGED ? sdesc[row] : __ldg(p.row_desc + row);              // {first entry, count, first column}
                const int cnt = d.y;
                const T *w = val + d.x;
                const T *pc = px + d.z;
                T acc = T(0);
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) acc = t_add_rn(acc, t_mul_rn(STAGED ? w[i] : __ldg(w + i), pc[i]));
                if (f < nf) out[static_cast<long long>(row) * p.out_row_stride] = amp_scale<T>(acc, p.amp, p.apply_db, eps);
            }
        };
        if (!PAIR && p.lane_w_smem != 0) rows(std::true_type{});
        else rows(std::false_type{});
        return;
    }
    // scratch for the fused-MFCC log-mel tile sits behind the power tile (host sizes the buffer for it)
    epilogue_from_power<T>(p, P, P + FT * p.tile_stride, clip, f0, nf);
    (void)N;
}

// Inverse of the above for istft / irfft (src/spectrogram.rs:4789-4911): one CTA = FT frames of one clip. The Hermitian
// half spectrum X (bins, frames) is staged frame-major through shared memory (global reads run along frames), every
// thread builds its 16 first-pass inputs conj(Z[n]), Z = E + i O with E = (X[n] + conj X[M-n]) / 2 and
// O = (X[n] - conj X[M-n]) / 2 * conj(W_N^n), the same register radix-16 Stockham passes run FORWARD on conj(Z), and
// x[2m] = Re Y[m] / M, x[2m+1] = -Im Y[m] / M go out times the synthesis window as time frames [clip][frame][N]
// (ola_gather in kernel_inverse.cu adds them up). DC / Nyquist imaginary parts are ignored, as realfft does.
template <typename T, int M, int FT>
__global__ void __launch_bounds__(FT *(M / 16), (sizeof(T) == 4 ? 4 : 2) * 256 / (FT * (M / 16) > 256 ? FT * (M / 16) : 256))
k_c2r_pow2(const __grid_constant__ KParams p, const typename Cplx<T>::type *__restrict__ stft, T *__restrict__ frames_out, long long n_frames,
           int apply_window) {
    constexpr int TPF = M / 16, N = 2 * M, B = M / 16, NT = FT * TPF;
    constexpr int ZS = zs_of(M, FT);
    using C = Cx<T>;
    using C2 = typename Cplx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *zbuf = reinterpret_cast<C *>(smem_raw);
    const int tid = threadIdx.x;
    int fl, t;
    ThreadMap<TPF, FT>::get(tid, fl, t);
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const long long f0 = static_cast<long long>(tile) * FT;
    const long long rem = n_frames - f0;
    const int nf = rem < FT ? static_cast<int>(rem) : FT;
    C *z = zbuf + fl * ZS;
    const C *tw = static_cast<const C *>(p.tw);
    const C *post = static_cast<const C *>(p.post);

    // ---- stage X[k][f0 .. f0 + nf) as S[f][k]: thread (kb, f) = (tid / FT, tid % FT) walks the bins kb, kb + TPF, ... with all
    //      17 loads issued before the first store (one DRAM round trip per tile)
    {
        const int kb = tid / FT, f = tid - kb * FT;
        const bool live = f < nf;
        const C2 *X = stft + static_cast<long long>(clip) * (M + 1) * n_frames + f0 + (live ? f : 0);
        C2 *S = reinterpret_cast<C2 *>(zbuf) + f * ZS;
        constexpr int KI = (M + 1 + TPF - 1) / TPF;
        C2 xv[KI];
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int k = kb + i * TPF;
            xv[i] = (live && k <= M) ? __ldg(X + static_cast<long long>(k) * n_frames) : mk<T>(T(0), T(0));
        }
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int k = kb + i * TPF;
            if (k <= M) S[k] = xv[i];
        }
    }
    __syncthreads();
    C v[16];
    {
        const C *S = z;                                // this frame's spectrum, natural order
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = t + j * B;
            C a = S[n], b = S[M - n];
            if (n == 0) { a.y = T(0); b.y = T(0); }
            const C e = {T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y)};
            const C d = {T(0.5) * (a.x - b.x), T(0.5) * (a.y + b.y)};
            const C w = ldg_cx<T>(post + n);
            const C o = {d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y};
            v[j] = {e.x - o.y, -(e.y + o.x)};
        }
    }
    __syncthreads();                                   // every spectrum value is in registers: the buffers may be overwritten
    pass_store<T, M, 16, 1>(z, t, v);
    frame_sync<TPF, FT>(fl);
    if constexpr (Radices<M>::P16 >= 2) {
        pass_load<T, M, 16, 16>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, 16, 16>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    if constexpr (Radices<M>::P16 >= 3) {
        pass_load<T, M, 16, 256>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, 16, 256>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    if constexpr (Radices<M>::LAST > 1) {
        constexpr int CUR = M / Radices<M>::LAST;
        pass_load<T, M, Radices<M>::LAST, CUR>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, Radices<M>::LAST, CUR>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    if (fl >= nf) return;
    // ---- conj, scale, window, store: thread t owns packed samples m = t + TPF * u
    const T scale = T(1) / static_cast<T>(M);
    const C *win = reinterpret_cast<const C *>(static_cast<const T *>(p.window));
    C *dst = reinterpret_cast<C *>(frames_out + (static_cast<long long>(clip) * n_frames + f0 + fl) * N);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const int m = t + TPF * u;
        const C y = z[pad16(m)];
        C x = {y.x * scale, -y.y * scale};
        if (apply_window) {
            const C w = ldg_cx<T>(win + m);
            x = {x.x * w.x, x.y * w.y};
        }
        dst[m] = x;
    }
}

// istft in ONE kernel (src/spectrogram.rs:4813-4911): inverse FFT, synthesis window, overlap-add and window-energy
// normalisation on a halo tile, so the windowed time frames never travel through HBM (they were n_fft / hop times the signal
// size written and read back). The overlap-add buffer is cut into slots of `hop` samples; a CTA owns FT - H consecutive slots
// (H = ceil(n_fft / hop) - 1) and inverts the FT frames that reach into them: frames s0 - H .. s0 + FT - H - 1. The frames stay
// in shared memory (conj, 1 / M scale and window applied in place by the thread that owns each pair); one thread per output
// sample then adds the frames that cover it in ascending frame order -- the reference's accumulation order (:4874-4881) --
// and divides by the accumulated squared window where it exceeds 1e-10 (:4885-4890). Redundancy: FT / (FT - H) inverse FFTs.
template <typename T, int M, int FT>
__global__ void __launch_bounds__(FT *(M / 16), (sizeof(T) == 4 ? 4 : 2) * 256 / (FT * (M / 16) > 256 ? FT * (M / 16) : 256))
k_istft_pow2(const __grid_constant__ KParams p, const typename Cplx<T>::type *__restrict__ stft, T *__restrict__ out, long long n_frames,
             int halo, long long out_len, long long trim) {
    constexpr int TPF = M / 16, N = 2 * M, B = M / 16, NT = FT * TPF;
    constexpr int ZS = zs_of(M, FT);
    using C = Cx<T>;
    using C2 = typename Cplx<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *zbuf = reinterpret_cast<C *>(smem_raw);
    const int tid = threadIdx.x;
    int fl, t;
    ThreadMap<TPF, FT>::get(tid, fl, t);
    const int clip = blockIdx.x / p.tiles_per_clip;
    const int tile = blockIdx.x - clip * p.tiles_per_clip;
    const int own = FT - halo;                                        // slots this CTA writes
    const long long s0 = static_cast<long long>(tile) * own;           // first slot
    const long long f_first = s0 - halo;                               // frame held in tile slot 0 (may be negative)
    C *z = zbuf + fl * ZS;
    const C *tw = static_cast<const C *>(p.tw);
    const C *post = static_cast<const C *>(p.post);

    // ---- stage X[k][f_first .. f_first + FT) bin-major as S[k][f] with rows of FT + 1 elements: the global reads run along
    //      frames, the shared-memory writes are consecutive, and the per-frame reads below (lanes along n, stride FT + 1) are at
    //      most 2-way conflicted. Frames outside the clip are zero.
    constexpr int SR = FT + 1;
    {
        // thread (kb, f) = (tid / FT, tid % FT) walks the bins kb, kb + TPF, ...: 17 independent loads, all issued before the
        // first store (one DRAM round trip per tile instead of 17)
        const int kb = tid / FT, f = tid - kb * FT;
        const long long fg = f_first + f;
        const bool live = fg >= 0 && fg < n_frames;
        const C2 *X = stft + static_cast<long long>(clip) * (M + 1) * n_frames + (live ? fg : 0);
        C2 *S = reinterpret_cast<C2 *>(zbuf) + f;
        constexpr int KI = (M + 1 + TPF - 1) / TPF;
        C2 xv[KI];
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int k = kb + i * TPF;
            xv[i] = (live && k <= M) ? __ldg(X + static_cast<long long>(k) * n_frames) : mk<T>(T(0), T(0));
        }
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int k = kb + i * TPF;
            if (k <= M) S[k * SR] = xv[i];
        }
    }
    __syncthreads();
    C v[16];
    {
        const C *S = zbuf + fl;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = t + j * B;
            C a = S[n * SR], b = S[(M - n) * SR];
            if (n == 0) { a.y = T(0); b.y = T(0); }
            const C e = {T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y)};
            const C d = {T(0.5) * (a.x - b.x), T(0.5) * (a.y + b.y)};
            const C w = ldg_cx<T>(post + n);
            const C o = {d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y};
            v[j] = {e.x - o.y, -(e.y + o.x)};
        }
    }
    __syncthreads();
    pass_store<T, M, 16, 1>(z, t, v);
    frame_sync<TPF, FT>(fl);
    if constexpr (Radices<M>::P16 >= 2) {
        pass_load<T, M, 16, 16>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, 16, 16>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    if constexpr (Radices<M>::P16 >= 3) {
        pass_load<T, M, 16, 256>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, 16, 256>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    if constexpr (Radices<M>::LAST > 1) {
        constexpr int CUR = M / Radices<M>::LAST;
        pass_load<T, M, Radices<M>::LAST, CUR>(z, tw, t, v);
        frame_sync<TPF, FT>(fl);
        pass_store<T, M, Radices<M>::LAST, CUR>(z, t, v);
        frame_sync<TPF, FT>(fl);
    }
    // ---- conj, scale, window in place: thread t owns packed samples m = t + TPF * u of its frame
    {
        const T scale = T(1) / static_cast<T>(M);
        const C *win = reinterpret_cast<const C *>(static_cast<const T *>(p.window));
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int m = t + TPF * u;
            const C y = z[pad16(m)];
            const C w = ldg_cx<T>(win + m);
            z[pad16(m)] = C{y.x * scale * w.x, -y.y * scale * w.y};        // time_frame[i] *= window[i] (:4868-4870)
        }
    }
    __syncthreads();
    // ---- overlap-add + normalisation of the slots this CTA owns. Slot s (tile-local), offset i0 < hop: position
    //      (s0 + s) hop + i0 is covered by the frames s0 + s - j at sample i0 + j hop (< n_fft), j = 0 .. halo; ascending frame
    //      order = descending j. No divisions: everything is tile-local 32-bit arithmetic.
    const int hop = p.hop;
    const long long total = (n_frames - 1) * static_cast<long long>(hop) + N;      // untrimmed length (:4837)
    const T *winT = static_cast<const T *>(p.window);
    T *oc = out + static_cast<long long>(clip) * out_len;
    const FastDiv fd_hop = p.fd_out_len;                               // the host stores make_fastdiv(hop) here for this kernel
    for (int idx = tid; idx < own * hop; idx += NT) {
        const int sl = fd_div(idx, fd_hop), i0 = idx - sl * hop;
        const long long slot = s0 + sl;                                 // this slot = frame index of its newest covering frame
        const long long pos0 = slot * hop;
        const long long o = pos0 + i0 - trim;
        if (pos0 + i0 >= total || o < 0 || o >= out_len) continue;
        // frames that exist: 0 <= slot - j <= n_frames - 1
        const int j_min = slot > n_frames - 1 ? static_cast<int>(slot - (n_frames - 1)) : 0;
        const int j_cap = slot < halo ? static_cast<int>(slot) : halo;
        T acc = T(0), norm = T(0);
        for (int j = j_cap; j >= j_min; --j) {                         // ascending frames (:4874-4881)
            const int i = i0 + j * hop;
            if (i >= N) continue;
            const C y = zbuf[(sl + halo - j) * ZS + pad16(i >> 1)];
            const T w = __ldg(winT + i);
            acc = t_add_rn(acc, (i & 1) ? y.y : y.x);
            norm = t_add_rn(norm, t_mul_rn(w, w));
        }
        if (norm > static_cast<T>(1e-10)) acc = acc / norm;            // :4885-4890
        oc[o] = acc;
    }
}

template <typename T, int M, int FT>
cudaError_t launch_istft_one(const KParams &p, size_t smem, const void *stft, void *out, long long n_clips, long long n_frames, int halo,
                             long long out_len, long long trim, cudaStream_t stream) {
    const long long grid = n_clips * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_istft_pow2<T, M, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_istft_pow2<T, M, FT><<<static_cast<unsigned>(grid), FT *(M / 16), smem, stream>>>(p, static_cast<const typename Cplx<T>::type *>(stft),
                                                                                        static_cast<T *>(out), n_frames, halo, out_len, trim);
    return cudaGetLastError();
}

template <typename T, int M, int FT>
cudaError_t launch_c2r_one(const KParams &p, size_t smem, const void *stft, void *frames_out, long long n_clips, long long n_frames,
                           int apply_window, cudaStream_t stream) {
    const long long grid = n_clips * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_c2r_pow2<T, M, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_c2r_pow2<T, M, FT><<<static_cast<unsigned>(grid), FT *(M / 16), smem, stream>>>(p, static_cast<const typename Cplx<T>::type *>(stft),
                                                                                      static_cast<T *>(frames_out), n_frames, apply_window);
    return cudaGetLastError();
}

template <typename T, int M, int FT, bool PAIR>
cudaError_t launch_one(const KParams &p, size_t smem, cudaStream_t stream) {
    const long long grid = static_cast<long long>(p.n_clips) * p.tiles_per_clip;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k_r2c_fused_pow2<T, M, FT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_r2c_fused_pow2<T, M, FT, PAIR><<<static_cast<unsigned>(grid), FT *(M / rpt_of<T>(M)), smem, stream>>>(p);
    return cudaGetLastError();
}

// frames per tile: 256 threads per CTA. (512-thread f64 tiles for n_fft >= 4096 store wider rows but leave one CTA per
// SM with nothing to overlap its load and drain phases with: 3.52 ms against 3.20 ms on BASELINE configs[4].)
constexpr int ft16_of(int M) {          // the inverse kernel keeps 16 values per thread for every size
    const int ft = 256 / (M / 16);
    return ft < 1 ? 1 : (ft > 32 ? 32 : ft);
}
constexpr int ft_of(int M, bool f64) {
    const int tpf = M / (f64 ? rpt_of<double>(M) : rpt_of<float>(M));
    // the one-warp-per-frame two-pass form (f32, M = 1024) runs 128-thread CTAs: 128 registers per thread allow 512 threads per
    // SM, and four small CTAs overlap their load / FFT / epilogue phases better than two large ones
    const int want = (!f64 && rpt_of<float>(M) == 32) ? 128 : 256;
    const int ft = want / tpf;
    return ft < 1 ? 1 : (ft > 32 ? 32 : ft);
}

}  // namespace

bool pow2_supported(size_t n_fft) {
    return n_fft >= 256 && n_fft <= 8192 && (n_fft & (n_fft - 1)) == 0;
}
int pow2_frames_per_tile(size_t n_fft, bool f64) { return ft_of(static_cast<int>(n_fft / 2), f64); }
// extra dynamic shared memory of the bulk-copy staged variant (0: the size does not have one)
size_t pow2_bulk_stage_bytes(size_t n_fft, size_t hop, bool f64) {
    const int M = static_cast<int>(n_fft / 2);
    if (f64 || rpt_of<float>(M) != 32 || hop > 1024 || hop % 4 != 0) return 0;
    return sizeof(float) * (static_cast<size_t>(ft_of(M, false) - 1) * 1024 + n_fft) + 16;
}
// CTAs per SM the forward kernel's launch bounds ask for (the shared-memory budget of a CTA follows from it)
int pow2_min_blocks(size_t n_fft, bool f64) {
    const int M = static_cast<int>(n_fft / 2), ft = ft_of(M, f64);
    if (!f64 && rpt_of<float>(M) == 32) return 512 / (ft * (M / 32));
    const int threads = ft * (M / 16);
    return (f64 ? 2 : 4) * 256 / (threads > 256 ? threads : 256);
}
int pow2_frame_elems(size_t n_fft, bool f64) {
    const int M = static_cast<int>(n_fft / 2);
    return zs_of(M, ft_of(M, f64), f64 ? rpt_of<double>(M) : rpt_of<float>(M));
}

cudaError_t launch_pow2(const KParams &p, bool f64, size_t smem, cudaStream_t stream) {
#define SGX_POW2_CASE(MM)                                                                         \
    case MM:                                                                                      \
        if (p.samples_b != nullptr)                                                                \
            return f64 ? launch_one<double, MM, ft_of(MM, true), true>(p, smem, stream) : launch_one<float, MM, ft_of(MM, false), true>(p, smem, stream); \
        return f64 ? launch_one<double, MM, ft_of(MM, true), false>(p, smem, stream) : launch_one<float, MM, ft_of(MM, false), false>(p, smem, stream);
    switch (p.n_fft / 2) {
        SGX_POW2_CASE(128)
        SGX_POW2_CASE(256)
        SGX_POW2_CASE(512)
        SGX_POW2_CASE(1024)
        SGX_POW2_CASE(2048)
        SGX_POW2_CASE(4096)
        default: return cudaErrorInvalidValue;
    }
#undef SGX_POW2_CASE
}

int pow2_c2r_frames_per_tile(size_t n_fft) { return ft16_of(static_cast<int>(n_fft / 2)); }

cudaError_t launch_c2r_pow2(const KParams &p, bool f64, const void *stft, void *frames_out, long long n_clips,
                            long long n_frames, int apply_window, cudaStream_t stream) {
#define SGX_POW2_C2R(MM)                                                                                                         \
    case MM:                                                                                                                     \
        return f64 ? launch_c2r_one<double, MM, ft16_of(MM)>(p, sizeof(double) * 2 * ft16_of(MM) * zs_of(MM, ft16_of(MM)), stft, frames_out, n_clips, n_frames, apply_window, stream) \
                   : launch_c2r_one<float, MM, ft16_of(MM)>(p, sizeof(float) * 2 * ft16_of(MM) * zs_of(MM, ft16_of(MM)), stft, frames_out, n_clips, n_frames, apply_window, stream);
    switch (p.n_fft / 2) {
        SGX_POW2_C2R(128)
        SGX_POW2_C2R(256)
        SGX_POW2_C2R(512)
        SGX_POW2_C2R(1024)
        SGX_POW2_C2R(2048)
        SGX_POW2_C2R(4096)
        default: return cudaErrorInvalidValue;
    }
#undef SGX_POW2_C2R
}

// complex elements of the fused istft's shared memory: the pass buffers, or the bin-major staging tile if that is larger
static constexpr size_t istft_elems(int M) {
    const size_t a = static_cast<size_t>(ft16_of(M)) * zs_of(M, ft16_of(M)), b = static_cast<size_t>(M + 1) * (ft16_of(M) + 1);
    return a > b ? a : b;
}
// fused istft on a halo tile (k_istft_pow2): p.tiles_per_clip = ceil(ceil(total / hop) / (FT - halo)) with FT =
// pow2_c2r_frames_per_tile(n_fft); requires halo < FT
cudaError_t launch_istft_pow2(const KParams &p, bool f64, const void *stft, void *out, long long n_clips, long long n_frames, int halo,
                              long long out_len, long long trim, cudaStream_t stream) {
#define SGX_POW2_ISTFT(MM)                                                                                                       \
    case MM:                                                                                                                     \
        return f64 ? launch_istft_one<double, MM, ft16_of(MM)>(p, sizeof(double) * 2 * istft_elems(MM), stft, out, n_clips, n_frames, halo, out_len, trim, stream) \
                   : launch_istft_one<float, MM, ft16_of(MM)>(p, sizeof(float) * 2 * istft_elems(MM), stft, out, n_clips, n_frames, halo, out_len, trim, stream);
    switch (p.n_fft / 2) {
        SGX_POW2_ISTFT(128)
        SGX_POW2_ISTFT(256)
        SGX_POW2_ISTFT(512)
        SGX_POW2_ISTFT(1024)
        SGX_POW2_ISTFT(2048)
        SGX_POW2_ISTFT(4096)
        default: return cudaErrorInvalidValue;
    }
#undef SGX_POW2_ISTFT
}

}  // namespace sgx
