// sgx_api.cu -- the C ABI (include/sgx_b200.h): plan objects, validation, table upload, dispatch, host staging.
// No torch types, no CPU compute path: without a CUDA device every compute call fails with SGX_BACKEND_ERROR.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/sgx_b200.h"
#include "kparams.cuh"
#include "launch.hpp"
#include "tables.hpp"

using namespace sgx;

// ------------------------------------------------------------------------------------------------ errors
namespace {

thread_local std::string g_err_msg = "";
thread_local size_t g_exp = 0, g_got = 0;

sgx_status set_error(const Error &e) {
    switch (e.code) {
        case SGX_INVALID_INPUT: g_err_msg = "Invalid input: " + e.msg; break;
        case SGX_DIMENSION_MISMATCH:
            g_exp = e.expected; g_got = e.got;
            g_err_msg = "Dimension mismatch: expected " + std::to_string(e.expected) + ", got " + std::to_string(e.got);
            break;
        case SGX_BACKEND_ERROR: g_err_msg = "cuda -- FFT backend error: " + e.msg; break;
        default: g_err_msg = "Internal error: " + e.msg; break;
    }
    return e.code;
}

[[noreturn]] void invalid(const std::string &m) { throw Error{SGX_INVALID_INPUT, m}; }
[[noreturn]] void backend(const std::string &m) { throw Error{SGX_BACKEND_ERROR, m}; }
[[noreturn]] void mismatch(size_t expected, size_t got) { throw Error{SGX_DIMENSION_MISMATCH, "", expected, got}; }

void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        cudaGetLastError();
        backend(std::string(what) + ": " + cudaGetErrorString(e));
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        ck(cudaGetDevice(&prev), "cudaGetDevice");
        if (dev != prev) ck(cudaSetDevice(dev), "cudaSetDevice");
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename F> sgx_status guarded(F &&f) {
    try {
        f();
        return SGX_OK;
    } catch (const Error &e) {
        return set_error(e);
    } catch (const std::bad_alloc &) {
        return set_error(Error{SGX_INTERNAL_ERROR, "out of host memory"});
    } catch (const std::exception &e) {
        return set_error(Error{SGX_INTERNAL_ERROR, e.what()});
    }
}

enum class PtrKind { Host, Device };
PtrKind ptr_kind(const void *p) {
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return PtrKind::Host; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? PtrKind::Device : PtrKind::Host;
}

template <typename T> void *upload_as(const std::vector<double> &src) {
    if (src.empty()) return nullptr;
    std::vector<T> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = static_cast<T>(src[i]);   // T::from_f64
    void *d = nullptr;
    ck(cudaMalloc(&d, tmp.size() * sizeof(T)), "cudaMalloc(table)");
    ck(cudaMemcpy(d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(table)");
    return d;
}
void *upload(const std::vector<double> &src, bool f64) { return f64 ? upload_as<double>(src) : upload_as<float>(src); }

// B operand of dct2_lifter_tc (launch.hpp): basis[n_mfcc][n_mels] (f64) cast to f32 and split into a TF32-exact hi part and the
// f32 remainder, per 8-mel K step as K-major core matrices; lifter weights behind
std::vector<float> build_dct_tc_blob(size_t n_mels, size_t n_mfcc, const std::vector<double> &basis, const std::vector<double> &lift) {
    const size_t kp = static_cast<size_t>(mfcc_tc_padded_mels(static_cast<int>(n_mels))), N = static_cast<size_t>(mfcc_tc_padded_coeffs(static_cast<int>(n_mfcc)));
    std::vector<float> blob(mfcc_tc_blob_floats(static_cast<int>(n_mels), static_cast<int>(n_mfcc)), 0.0f);
    const size_t half = N * 8;
    for (size_t step = 0; step < kp / 8; ++step)
        for (size_t n = 0; n < N; ++n)
            for (size_t k = 0; k < 8; ++k) {
                const size_t mel = 8 * step + k;
                const float w = (n < n_mfcc && mel < n_mels) ? static_cast<float>(basis[n * n_mels + mel]) : 0.0f;
                uint32_t bits;
                std::memcpy(&bits, &w, 4);
                bits &= 0xffffe000u;
                float hi;
                std::memcpy(&hi, &bits, 4);
                const size_t idx = (n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4);
                blob[step * 2 * half + idx] = hi;
                blob[step * 2 * half + half + idx] = w - hi;
            }
    for (size_t c = 0; c < n_mfcc; ++c) blob[(kp / 8) * 2 * half + c] = static_cast<float>(lift[c]);
    return blob;
}
float *upload_floats(const std::vector<float> &v) {
    float *d = nullptr;
    if (v.empty()) return nullptr;
    ck(cudaMalloc(&d, v.size() * sizeof(float)), "cudaMalloc(table)");
    ck(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(table)");
    return d;
}
bool mfcc_tc_enabled() {
    static const bool off = std::getenv("SGX_MFCC_TC") && std::atoi(std::getenv("SGX_MFCC_TC")) == 0;
    return !off;
}
int *upload_int(const std::vector<int> &src) {
    if (src.empty()) return nullptr;
    int *d = nullptr;
    ck(cudaMalloc(&d, src.size() * sizeof(int)), "cudaMalloc(index)");
    ck(cudaMemcpy(d, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(index)");
    return d;
}

constexpr long double kPiL = 3.14159265358979323846264338327950288L;
// interleaved (cos, sin) of -2*pi*k/n, k = 0..count-1, evaluated in extended precision
std::vector<double> twiddle_table(size_t n, size_t count) {
    std::vector<double> t(2 * count);
    for (size_t k = 0; k < count; ++k) {
        const long double a = -2.0L * kPiL * static_cast<long double>(k % n) / static_cast<long double>(n);
        t[2 * k] = static_cast<double>(cosl(a));
        t[2 * k + 1] = static_cast<double>(sinl(a));
    }
    return t;
}

constexpr int kStagingSlots = 3;

}  // namespace

// ------------------------------------------------------------------------------------------------ plan
struct sgx_plan {
    sgx_plan_desc desc{};
    std::vector<double> custom;
    HostTables tab;
    int device = 0;
    bool f64 = false;
    size_t esize = 4;
    size_t rows = 0;                 // output rows
    // FFT factorisation
    int L = 0, even = 0;
    std::vector<int> radix;
    // device tables
    void *d_window = nullptr, *d_tw = nullptr, *d_post = nullptr, *d_val = nullptr, *d_dense = nullptr;
    void *d_dct = nullptr, *d_lifter = nullptr, *d_dct_folded = nullptr;
    std::vector<double> dct_folded;  // [tasks][n_mels/2][4]
    int dct_tasks = 0;
    int *d_row_ptr = nullptr, *d_col = nullptr, *d_wofs = nullptr, *d_wofs_tm = nullptr;
    std::vector<int> wofs;           // quad schedule blob of the sparse mapping (r2c_fused_n400)
    std::vector<int> wofs_tm;        // the same quads in one descending-cost list (r2c_fused_n400_tm)
    // generic-family geometry
    int FT = 1, buf_elems = 0, frame_stride = 0, tile_stride = 0;
    size_t smem_bytes = 0;
    // bookkeeping
    bool force_generic = false;
    bool on_device = false;
    size_t last_launches = 0;
    std::string kernel_name = "r2c_fused_generic";
    bool fast400 = false;            // eligible for r2c_fused_n400
    bool pow2 = false;               // eligible for r2c_fused_pow2
    bool mixed = false;              // eligible for r2c_fused_mixed (even 2^a 3^b 5^c sizes with a compiled instance)
    int pow2_ft = 1, pow2_frame_stride = 0, pow2_tile_stride = 0;
    size_t pow2_smem = 0;
    bool fast400_sparse = false;     // ... with the shared-memory sparse table
    bool fast400_tc = false;         // ... on the TMEM / tcgen05 kernel (r2c_fused_n400_tc)
    bool fast400_tm = false;         // ... on the TMEM-exchange kernel (r2c_fused_n400_tm): sparse mel / loghz spectrogram outputs
    int tm_mode = -1;                // sgx_plan_set_tmem_exchange: -1 auto (used wherever it applies), 0 never, 1 whenever available
    int tm_warps = 4;                // warps per 32-frame group of r2c_fused_n400_tm (4, 5 or 6; SGX_N400_TM_WARPS)
    int tc_mode = -1;                // sgx_plan_set_tensor_cores: -1 auto (dense mappings only), 0 never, 1 whenever available
    std::vector<int> tc_blob;        // step blob of r2c_fused_n400_tc (launch.hpp)
    int tc_steps = 0, tc_rounds = 0;
    size_t tc_b_floats = 0;
    int *d_tc_blob = nullptr;
    int sparse_quads = 0, sparse_weights = 0, tm_weights = 0;
    bool rows_contig = false;        // CSR rows have consecutive columns
    std::vector<int> row_desc;       // contiguous CSR rows: int4 {e0, cnt, c0, 0} per row
    int *d_row_desc = nullptr;
    std::vector<int> lane_rows;      // r2c_fused_pow2 rows epilogue: int4 per lane slot
    std::vector<double> lane_w;      // ... and its lane-major weights
    int *d_lane_rows = nullptr;
    void *d_frames = nullptr;               // istft: windowed time frames of a chunk of clips
    bool mfcc_split = false;                // n400 f32 mfcc(): log-mel by r2c_fused_n400_tm, DCT-II by dct2_lifter_tc (tcgen05)
    float *d_dct_tc = nullptr;              // ... basis blob of dct2_lifter_tc
    // dense (ERB) f32 spectrogram plans outside the n400_tc kernel: linear power spectrogram of a chunk of clips into plan scratch by
    // the plan's FFT family, then the filterbank as row blocks of the tcgen05 GEMM (dense_rows_tc = the dct2_lifter_tc kernel)
    std::vector<int> wide_rows;             // r2c_fused_pow2 wide tiles (FT >= 16): int4 row descriptors + CSR values (bit patterns of T), staged into shared memory
    void *d_wide_rows = nullptr;
    std::vector<int> mixed_q;               // r2c_fused_mixed quad epilogue: {n_quads,0,0,0} + int4 entries + zero-padded f32 weights (bit patterns)
    int *d_mixed_q = nullptr;
    bool eps_denormal = false;              // dB floor below the f32 normal range: kernels that take lg2.approx.ftz are not used
    bool dense_split = false;
    int dense_block_rows = 0;               // rows per GEMM pass
    std::vector<float *> d_dense_tc;        // one B-operand blob per row block
    void *d_logmel = nullptr;               // ... log-mel scratch of a chunk of clips
    size_t logmel_cap = 0;
    size_t frames_cap = 0;
    void *d_pair[2] = {nullptr, nullptr};   // binaural: complex STFTs of the two channels of a chunk of pairs
    // Device-pointer istft / binaural calls return asynchronously but stage through the plan-owned scratch above: the last
    // use is recorded here and the next call (possibly on another stream) waits for it before touching the scratch again.
    cudaEvent_t scratch_done = nullptr;
    void scratch_acquire(cudaStream_t st) {
        if (!scratch_done) { if (cudaEventCreateWithFlags(&scratch_done, cudaEventDisableTiming) != cudaSuccess) { scratch_done = nullptr; return; } }
        else cudaStreamWaitEvent(st, scratch_done, 0);
    }
    void scratch_release(cudaStream_t st) { if (scratch_done) cudaEventRecord(scratch_done, st); }
    size_t pair_cap = 0;
    std::vector<double> dense_t;     // chroma: dense matrix transposed to [out_len][n_bins]
    void *d_dense_t = nullptr;
    int dense_c0 = 0, dense_c1 = 0;  // nonzero column range of the dense matrix
    void *d_lane_w = nullptr;
    int sm_count = 148;
    std::vector<float> window_f32;
    // staging for host-pointer calls
    struct Slot { void *d_in = nullptr; void *d_out = nullptr; size_t in_cap = 0, out_cap = 0; cudaStream_t s = nullptr; } slot[kStagingSlots];

    ~sgx_plan() {
        if (!on_device) return;
        int prev = -1;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) cudaSetDevice(device); else prev = -1;
        for (void *p : {d_window, d_tw, d_post, d_val, d_dense, d_dct, d_lifter, d_dct_folded}) if (p) cudaFree(p);
        if (d_row_ptr) cudaFree(d_row_ptr);
        if (d_col) cudaFree(d_col);
        if (d_wofs) cudaFree(d_wofs);
        if (d_wofs_tm) cudaFree(d_wofs_tm);
        if (d_tc_blob) cudaFree(d_tc_blob);
        if (d_lane_rows) cudaFree(d_lane_rows);
        if (d_row_desc) cudaFree(d_row_desc);
        if (d_lane_w) cudaFree(d_lane_w);
        if (d_dense_t) cudaFree(d_dense_t);
        for (void *q : d_pair) if (q) cudaFree(q);
        if (d_frames) cudaFree(d_frames);
        if (d_dct_tc) cudaFree(d_dct_tc);
        for (float *q : d_dense_tc) if (q) cudaFree(q);
        if (d_mixed_q) cudaFree(d_mixed_q);
        if (d_wide_rows) cudaFree(d_wide_rows);
        if (d_logmel) cudaFree(d_logmel);
        if (scratch_done) cudaEventDestroy(scratch_done);
        for (auto &s : slot) {
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.s) cudaStreamDestroy(s.s);
        }
        if (prev >= 0) cudaSetDevice(prev);
    }
};

namespace {

void factorise(sgx_plan &pl) {
    const size_t n = pl.desc.n_fft;
    pl.even = (n % 2 == 0) ? 1 : 0;
    size_t L = pl.even ? n / 2 : n;
    if (L > (1u << 24)) backend("n_fft too large for the CUDA plan");
    pl.L = static_cast<int>(L);
    size_t rem = L;
    pl.radix.clear();
    while (rem % 4 == 0) { pl.radix.push_back(4); rem /= 4; }
    while (rem % 2 == 0) { pl.radix.push_back(2); rem /= 2; }
    while (rem % 3 == 0) { pl.radix.push_back(3); rem /= 3; }
    while (rem % 5 == 0) { pl.radix.push_back(5); rem /= 5; }
    if (rem > 1) pl.radix.insert(pl.radix.begin(), static_cast<int>(rem));   // cofactor stage (primes >= 7), evaluated directly
    if (static_cast<int>(pl.radix.size()) > kMaxStages) backend("too many FFT stages");
}


// Lane-major schedule of the filterbank rows for the rows-per-thread epilogue of r2c_fused_pow2 (one lane = one row, all
// frames of the tile): rows sorted by length so that a warp's 32 rows finish together, quarter-warps arranged so that
// their first columns differ modulo 8 (the 16-byte tile reads of 8 lanes then hit 8 different bank groups), and the
// weights transposed per warp block to [entry][lane] so that one warp-wide weight load is one 128-byte line instead of
// 32 lines. Dense ERB rows are the same thing with first column 0 and length out_len.
void build_lane_rows(sgx_plan &pl) {
    pl.lane_rows.clear();
    pl.lane_w.clear();
    const sgx_plan_desc &d = pl.desc;
    const bool dense = d.mapping == SGX_MAP_ERB;
    if (!(dense || pl.rows_contig) || d.output != SGX_OUT_SPECTROGRAM) return;
    const size_t nb = pl.tab.n_bins, ol = pl.tab.out_len;
    std::vector<int> c0(nb), cnt(nb);
    for (size_t r = 0; r < nb; ++r) {
        if (dense) { c0[r] = 0; cnt[r] = static_cast<int>(ol); continue; }
        const int e0 = pl.tab.row_ptr[r], e1 = pl.tab.row_ptr[r + 1];
        cnt[r] = e1 - e0;
        c0[r] = cnt[r] ? pl.tab.col[e0] : 0;
    }
    std::vector<int> order(nb);
    for (size_t r = 0; r < nb; ++r) order[r] = static_cast<int>(r);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
    while (order.size() % 32) order.push_back(-1);
    for (size_t w0 = 0; w0 < order.size(); w0 += 32) {
        // greedy: fill each group of 8 lanes with rows whose (first column mod 8) is not in the group yet
        std::vector<int> pool(order.begin() + w0, order.begin() + w0 + 32), arranged;
        while (!pool.empty()) {
            bool used[8] = {false, false, false, false, false, false, false, false};
            for (int k = 0; k < 8 && !pool.empty(); ++k) {
                size_t pick = 0;
                for (size_t i = 0; i < pool.size(); ++i)
                    if (pool[i] >= 0 && !used[c0[pool[i]] & 7]) { pick = i; break; }
                if (pool[pick] >= 0) used[c0[pool[pick]] & 7] = true;
                arranged.push_back(pool[pick]);
                pool.erase(pool.begin() + pick);
            }
        }
        int maxc = 0;
        for (int r : arranged) if (r >= 0) maxc = std::max(maxc, cnt[r]);
        // A warp runs as long as its longest row, so a shorter row may start up to (maxc - cnt) columns early on zero weights at no
        // cost (acc = 0 + 0 * x stays 0: results unchanged bit for bit). The slack is used to give the 8 lanes of every quarter warp
        // distinct (first column mod 8): their 16-byte tile reads then hit 8 different bank groups at every step of the walk.
        std::vector<int> shift(32, 0);
        if (!dense)
            for (int g0 = 0; g0 < 32; g0 += 8) {
                bool used[8] = {false, false, false, false, false, false, false, false};
                std::vector<int> lanes;
                for (int l = g0; l < g0 + 8; ++l) if (arranged[l] >= 0 && cnt[arranged[l]] > 0) lanes.push_back(l);
                std::stable_sort(lanes.begin(), lanes.end(), [&](int a, int b) {       // least slack first
                    return std::min(maxc - cnt[arranged[a]], c0[arranged[a]]) < std::min(maxc - cnt[arranged[b]], c0[arranged[b]]);
                });
                for (int l : lanes) {
                    const int r = arranged[l], slack = std::min(maxc - cnt[r], c0[r]);
                    int pick = 0;
                    for (int dlt = 0; dlt <= std::min(slack, 7); ++dlt)
                        if (!used[(c0[r] - dlt) & 7]) { pick = dlt; break; }
                    shift[l] = pick;
                    used[(c0[r] - pick) & 7] = true;
                }
            }
        const size_t wofs = pl.lane_w.size();
        pl.lane_w.resize(wofs + static_cast<size_t>(maxc) * 32, 0.0);
        for (int lane = 0; lane < 32; ++lane) {
            const int r = arranged[lane], sh = shift[lane];
            pl.lane_rows.push_back(r);
            pl.lane_rows.push_back(r >= 0 ? c0[r] - sh : 0);
            pl.lane_rows.push_back(r >= 0 ? cnt[r] + sh : 0);
            pl.lane_rows.push_back(static_cast<int>(wofs));
            if (r < 0) continue;
            for (int i = 0; i < cnt[r]; ++i)
                pl.lane_w[wofs + static_cast<size_t>(i + sh) * 32 + lane] =
                    dense ? pl.tab.dense[static_cast<size_t>(r) * ol + i] : pl.tab.val[pl.tab.row_ptr[r] + i];
        }
    }
}

// r2c_fused_n400_tc is measured slower than the CUDA-core kernel for the banded mel / loghz filterbanks (2.1 ms against
// 1.5 ms per configs[1] step: ~80 MMAs of ~125 cycles each per 128 frames) and faster for the dense ERB projection, whose
// cost on CUDA cores grows with n_filters x 201 while the MMA count does not. Auto therefore picks it for ERB only.
// r2c_fused_n400_tm: TMEM as the exchange medium of the two FFT passes (kernel_n400_tm.cu)
bool use_tm(const sgx_plan &pl) {
    return pl.fast400 && pl.fast400_tm && !pl.force_generic && pl.tm_mode != 0 && pl.tc_mode != 1;
}

bool use_tc(const sgx_plan &pl) {
    if (!pl.fast400 || !pl.fast400_tc || pl.force_generic || pl.tc_mode == 0) return false;
    return pl.tc_mode == 1 || pl.desc.mapping == SGX_MAP_ERB;
}

// Step blob of r2c_fused_n400_tc: the filterbank as the B operand of tcgen05.mma. Rows are taken 16 at a time (one MMA
// block, N = 16) and 64 at a time per round (the D columns); for every block only the 8-bin K steps between its first and
// last non-zero column are emitted (mel / loghz rows are banded; a dense ERB block spans all 26 steps). Each step carries
// its weights as two K-major SWIZZLE_NONE tiles of 16 rows x 8 bins -- T::from_f64(w) split into a TF32-exact hi part and
// the f32 remainder -- laid out as 8 x 16-byte core matrices: float index (n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + k % 4.
void build_tc_blob(sgx_plan &pl) {
    pl.tc_blob.clear();
    pl.tc_steps = pl.tc_rounds = 0;
    const sgx_plan_desc &d = pl.desc;
    const bool csr = d.mapping == SGX_MAP_MEL || d.mapping == SGX_MAP_LOGHZ;
    const bool dense = d.mapping == SGX_MAP_ERB;
    if (!pl.fast400 || !(csr || dense) || d.output != SGX_OUT_SPECTROGRAM) return;
    const size_t nb = pl.tab.n_bins, ol = pl.tab.out_len;
    auto weight = [&](size_t r, size_t k) -> double {
        if (r >= nb || k >= ol) return 0.0;
        if (dense) return pl.tab.dense[r * ol + k];
        for (int e = pl.tab.row_ptr[r]; e < pl.tab.row_ptr[r + 1]; ++e)
            if (static_cast<size_t>(pl.tab.col[e]) == k) return pl.tab.val[e];
        return 0.0;
    };
    const int n_rounds = static_cast<int>((nb + 63) / 64);
    std::vector<int> round_start(1, 0), steps;
    std::vector<float> bw;
    // first / last 8-bin K step holding a non-zero weight (as T) of rows [r0, r1); an all-zero range still clears its D columns
    auto k_range = [&](size_t r0, size_t r1, long &j0, long &j1) {
        long cmin = -1, cmax = -1;
        for (size_t r = r0; r < std::min(nb, r1); ++r)
            for (size_t k = 0; k < ol; ++k)
                if (static_cast<float>(weight(r, k)) != 0.0f) {
                    if (cmin < 0 || static_cast<long>(k) < cmin) cmin = static_cast<long>(k);
                    cmax = std::max(cmax, static_cast<long>(k));
                }
        if (cmin < 0) cmin = cmax = 0;
        j0 = cmin / 8;
        j1 = cmax / 8;
    };
    auto emit = [&](size_t r0, int N, int dcol, long j0, long j1) {
        for (long j = j0; j <= j1; ++j) {
            steps.insert(steps.end(), {static_cast<int>(8 * j), dcol, static_cast<int>(bw.size() * sizeof(float)), (j != j0 ? 1 : 0) | (N << 8)});
            const size_t base = bw.size(), half = static_cast<size_t>(N) * 8;
            bw.resize(base + 2 * half, 0.0f);
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < 8; ++k) {
                    const float w = static_cast<float>(weight(r0 + static_cast<size_t>(n), static_cast<size_t>(8 * j + k)));
                    uint32_t bits;
                    std::memcpy(&bits, &w, 4);
                    bits &= 0xffffe000u;
                    float hi;
                    std::memcpy(&hi, &bits, 4);
                    const size_t idx = static_cast<size_t>((n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4));
                    bw[base + idx] = hi;
                    bw[base + half + idx] = w - hi;
                }
        }
    };
    for (int round = 0; round < n_rounds; ++round) {
        // an MMA costs ~125 cycles whatever its N (tools/ubench/mma_rate_probe.cu), so a round of 64 rows is issued either as
        // four banded N = 16 blocks or as one N = 64 block over the union of their K ranges -- whichever needs fewer MMAs
        const size_t r0 = static_cast<size_t>(round) * 64;
        long u0, u1, banded = 0;
        k_range(r0, r0 + 64, u0, u1);
        for (int b = 0; b < 4 && r0 + 16 * static_cast<size_t>(b) < nb; ++b) {
            long j0, j1;
            k_range(r0 + 16 * static_cast<size_t>(b), r0 + 16 * static_cast<size_t>(b) + 16, j0, j1);
            banded += j1 - j0 + 1;
        }
        if (u1 - u0 + 1 <= banded) {
            emit(r0, 64, 0, u0, u1);
        } else {
            for (int b = 0; b < 4 && r0 + 16 * static_cast<size_t>(b) < nb; ++b) {
                long j0, j1;
                k_range(r0 + 16 * static_cast<size_t>(b), r0 + 16 * static_cast<size_t>(b) + 16, j0, j1);
                emit(r0 + 16 * static_cast<size_t>(b), 16, 16 * b, j0, j1);
            }
        }
        round_start.push_back(static_cast<int>(steps.size() / 4));
    }
    const size_t b_floats = bw.size();
    const int n_steps = static_cast<int>(steps.size() / 4);
    if (!fast400_tc_fits(n_steps, n_rounds, b_floats)) return;
    std::vector<int> blob;
    blob.push_back(n_steps);
    blob.push_back(n_rounds);
    blob.insert(blob.end(), round_start.begin(), round_start.end());
    while (blob.size() % 4) blob.push_back(0);
    blob.insert(blob.end(), steps.begin(), steps.end());
    const size_t fofs = blob.size();
    blob.resize(fofs + bw.size());
    std::memcpy(blob.data() + fofs, bw.data(), bw.size() * sizeof(float));
    pl.tc_blob = std::move(blob);
    pl.tc_steps = n_steps;
    pl.tc_rounds = n_rounds;
    pl.tc_b_floats = b_floats;
}

void select_family(sgx_plan &pl) {
    const sgx_plan_desc &d = pl.desc;
    // the n400 kernels take lg2.approx.ftz of max(v, eps): an eps that is an f32 denormal (floor_db below about -379 dB) would be
    // flushed to zero and silent bins would come out as -inf instead of the floor, so such plans stay on the other families
    const bool eps_denormal = d.amp == SGX_AMP_DECIBELS && d.has_floor_db &&
                              static_cast<float>(std::pow(10.0, d.floor_db / 10.0)) < 1.17549435e-38f;
    pl.eps_denormal = eps_denormal;
    pl.fast400 = !pl.f64 && !eps_denormal && d.n_fft == 400 && d.hop_size == 160 && d.output != SGX_OUT_COMPLEX_STFT &&
                 (d.output != SGX_OUT_MFCC || static_cast<int>(pl.tab.n_bins) <= fast400_max_scratch_rows());
    if (pl.fast400) {
        pl.window_f32.resize(d.n_fft);
        for (size_t i = 0; i < d.n_fft; ++i) pl.window_f32[i] = static_cast<float>(pl.tab.window[i]);
    }
    // power-of-two family
    pl.pow2 = false;
    if (!pl.fast400 && pow2_supported(d.n_fft)) {
        const size_t es = pl.esize;
        const int ft = pow2_frames_per_tile(d.n_fft, pl.f64);
        const size_t zs = static_cast<size_t>(pow2_frame_elems(d.n_fft, pl.f64));
        size_t ts = std::max(pl.tab.out_len, pl.tab.n_bins);
        if (ts % 2 == 0) ts += 1;
        const size_t smem = std::max(ft * zs * 2 * es, 2 * ft * ts * es);
        if (smem <= 200 * 1024) {
            pl.pow2 = true;
            pl.pow2_ft = ft;
            pl.pow2_frame_stride = static_cast<int>(zs);
            pl.pow2_tile_stride = static_cast<int>(ts);
            pl.pow2_smem = smem;
        }
    }
    // mixed-radix family: the other even 2^a 3^b 5^c sizes (n_fft 400 in f64 or at another hop included)
    pl.mixed = !pl.fast400 && !pl.pow2 && mixed_supported(d.n_fft, pl.f64) &&
               (d.output != SGX_OUT_MFCC || static_cast<int>(pl.tab.n_bins) <= mixed_max_scratch_rows(d.n_fft));
    const bool csr = d.mapping == SGX_MAP_MEL || d.mapping == SGX_MAP_LOGHZ;
    // The shared-memory sparse schedule needs rows with contiguous columns (mel triangles, loghz pairs). Rows are sorted
    // by column count, grouped four at a time ("quads", padded with row = -1), and the quads are dealt to the kernel's
    // warps longest-first onto the least loaded warp. Blob: int n_quads; int qrange[W + 1]; int maxcnt[n_quads];
    // pad to 16 bytes; int4 {c0, cnt, padded weight offset, row}[4 * n_quads] in warp order.
    bool contiguous = csr;
    int padded = 0;
    pl.mfcc_split = false;
    {
        const char *e = std::getenv("SGX_N400_TM_WARPS");      // experiments only: 4 (default), 5 or 6 warps per group
        const int w = e ? std::atoi(e) : 4;
        pl.tm_warps = w >= 4 && w <= 6 ? w : 4;
    }
    pl.wofs.clear();
    pl.wofs_tm.clear();
    if (csr) {
        const int W = fast400_warps();
        const size_t nb = pl.tab.n_bins;
        std::vector<int> cnt(nb), wo(nb);
        for (size_t r = 0; r < nb; ++r) {
            const int e0 = pl.tab.row_ptr[r], e1 = pl.tab.row_ptr[r + 1];
            for (int e = e0 + 1; e < e1; ++e) contiguous = contiguous && pl.tab.col[e] == pl.tab.col[e - 1] + 1;
            cnt[r] = e1 - e0;
            wo[r] = padded;
            padded += (e1 - e0 + 3) & ~3;
        }
        std::vector<int> order(nb);
        for (size_t r = 0; r < nb; ++r) order[r] = static_cast<int>(r);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
        while (order.size() % 4) order.push_back(-1);
        const int nq = static_cast<int>(order.size() / 4);
        std::vector<int> qmax(nq, 0);
        for (int q = 0; q < nq; ++q)
            for (int k = 0; k < 4; ++k)
                if (order[4 * q + k] >= 0) qmax[q] = std::max(qmax[q], cnt[order[4 * q + k]]);
        // r2c_fused_mixed, f32: the same rows as one flat table for the quad epilogue (quads longest first, dealt to the warps round
        // robin): {n_quads, 0, 0, 0}, int4 {byte offset of P[c0], cnt, byte offset of the row's weights, row}[4 n_quads], then the
        // weights, every row zero-padded to a multiple of four
        pl.mixed_q.clear();
        if (contiguous && !pl.f64 && nq > 0) {
            std::vector<int> tq(4 + 16 * static_cast<size_t>(nq), 0);
            std::vector<float> wq;
            tq[0] = nq;
            for (int q = 0; q < nq; ++q)
                for (int k = 0; k < 4; ++k) {
                    const int r = order[4 * q + k];
                    int *e = &tq[4 + 4 * (4 * static_cast<size_t>(q) + k)];
                    e[0] = (r >= 0 && cnt[r]) ? pl.tab.col[pl.tab.row_ptr[r]] * 128 : 0;
                    e[1] = r >= 0 ? cnt[r] : 0;
                    e[2] = static_cast<int>(wq.size() * sizeof(float));
                    e[3] = r;
                    if (r < 0) continue;
                    for (int i = 0; i < cnt[r]; ++i) wq.push_back(static_cast<float>(pl.tab.val[pl.tab.row_ptr[r] + i]));
                    while (wq.size() % 4) wq.push_back(0.0f);
                }
            if (wq.empty()) wq.assign(4, 0.0f);
            const size_t at = tq.size();
            tq.resize(at + wq.size());
            std::memcpy(&tq[at], wq.data(), wq.size() * sizeof(float));
            pl.mixed_q = tq;
        }
        auto make_blob = [&](int W) {
            std::vector<std::vector<int>> per_warp(W);
            std::vector<long> load(W, 0);
            for (int q = 0; q < nq; ++q) {          // quads are already in descending cost order
                int best = 0;
                for (int w = 1; w < W; ++w) if (load[w] < load[best]) best = w;
                per_warp[best].push_back(q);
                load[best] += 26 + 16 * qmax[q];
            }
            const int hdr = (1 + W + 1 + nq + 3) & ~3;
            std::vector<int> blob(static_cast<size_t>(hdr) + 16 * static_cast<size_t>(nq), 0);
            blob[0] = nq;
            int idx = 0;
            for (int w = 0; w < W; ++w) {
                blob[1 + w] = idx;
                for (int q : per_warp[w]) {
                    blob[1 + W + 1 + idx] = qmax[q];
                    for (int k = 0; k < 4; ++k) {
                        const int r = order[4 * q + k];
                        int *e = &blob[static_cast<size_t>(hdr) + 4 * (4 * static_cast<size_t>(idx) + k)];
                        e[0] = (r >= 0 && cnt[r]) ? pl.tab.col[pl.tab.row_ptr[r]] : 0;
                        e[1] = r >= 0 ? cnt[r] : 0;
                        e[2] = r >= 0 ? wo[r] : 0;
                        e[3] = r;
                    }
                    ++idx;
                }
            }
            blob[1 + W] = idx;
            return blob;
        };
        std::vector<int> blob = make_blob(W);
        // r2c_fused_n400_tm: the quads dealt to the 4 warps of a group; every row of a quad is padded (zero weights) to the
        // quad's longest row (weight slots rounded up to 4 floats), so the kernel's trip count is warp-uniform. Warps 0 and 1 of a group
        // also run one more pass-1 task than warps 2 and 3 in the same phase: they start with that much load.
        {
            const int Wt = pl.tm_warps;
            static const long bias = std::getenv("SGX_N400_TM_BIAS") ? std::atol(std::getenv("SGX_N400_TM_BIAS")) : 200;
            std::vector<int> cntU(nq), woq(4 * static_cast<size_t>(nq));
            int padded_tm = 0;
            bool ok = nq > 0;
            for (int q = 0; q < nq; ++q) {
                cntU[q] = std::max(4, (qmax[q] + 3) & ~3);
                for (int k = 0; k < 4; ++k) {
                    const int r = order[4 * q + k] >= 0 ? order[4 * q + k] : order[4 * q];
                    woq[4 * static_cast<size_t>(q) + k] = padded_tm;
                    padded_tm += (cntU[q] + 3) & ~3;
                    const int c0 = (r >= 0 && cnt[r]) ? pl.tab.col[pl.tab.row_ptr[r]] : 0;
                    if (c0 + cntU[q] > static_cast<int>(pl.tab.out_len) + fast400_tm_pad_rows()) ok = false;
                }
            }
            std::vector<std::vector<int>> per_warp(Wt);
            std::vector<long> load(Wt, 0);
            // pass-1 column pairs of warp w: w, w + Wt, ... < 10 -- with four warps the bulk-staged kernel runs them as column quads
            // w and, on warp 0, quad 4: four, two, two and two pairs
            for (int w = 0; w < Wt; ++w) load[w] = bias * (Wt == 4 ? (w == 0 ? 4 : 2) : (10 - w + Wt - 1) / Wt);
            for (int q = 0; q < nq; ++q) {
                int best = 0;
                for (int w = 1; w < Wt; ++w) if (load[w] < load[best]) best = w;
                per_warp[best].push_back(q);
                load[best] += 26 + 4 * cntU[q];
            }
            const int hdr = (1 + Wt + 1 + nq + 3) & ~3;
            std::vector<int> tb(static_cast<size_t>(hdr) + 16 * static_cast<size_t>(nq), 0);
            tb[0] = nq;
            int idx = 0;
            for (int w = 0; w < Wt; ++w) {
                tb[1 + w] = idx;
                for (int q : per_warp[w]) {
                    tb[1 + Wt + 1 + idx] = cntU[q];
                    for (int k = 0; k < 4; ++k) {
                        const int r = order[4 * q + k] >= 0 ? order[4 * q + k] : order[4 * q];      // pad slots repeat the quad's first row
                        int *e = &tb[static_cast<size_t>(hdr) + 4 * (4 * static_cast<size_t>(idx) + k)];
                        e[0] = (r >= 0 && cnt[r]) ? pl.tab.col[pl.tab.row_ptr[r]] : 0;
                        e[1] = cntU[q];
                        e[2] = woq[4 * static_cast<size_t>(q) + k];
                        e[3] = r;
                    }
                    ++idx;
                }
            }
            tb[1 + Wt] = idx;
            // behind the table: the rows' weights as f32 bit patterns in the kernel's padded layout (row k of quad q at woq[4 q + k], zeros
            // up to the quad's padded count), so that a CTA stages them with one coalesced copy instead of a serial gather per row
            {
                const size_t wat = tb.size();
                tb.resize(wat + static_cast<size_t>(std::max(padded_tm, 4)), 0);
                for (int q = 0; q < nq && ok; ++q)
                    for (int k = 0; k < 4; ++k) {
                        const int r = order[4 * q + k] >= 0 ? order[4 * q + k] : order[4 * q];
                        if (r < 0) continue;
                        const int e0 = pl.tab.row_ptr[r], n = pl.tab.row_ptr[r + 1] - e0;
                        for (int j = 0; j < n && j < cntU[q]; ++j) {
                            const float v = static_cast<float>(pl.tab.val[static_cast<size_t>(e0 + j)]);
                            std::memcpy(&tb[wat + static_cast<size_t>(woq[4 * static_cast<size_t>(q) + k] + j)], &v, 4);
                        }
                    }
            }
            if (ok) {
                pl.wofs_tm = tb;
                pl.tm_weights = std::max(padded_tm, 4);
            }
        }
        pl.wofs = blob;
        padded = std::max(padded, 4);
        pl.sparse_quads = nq;
        pl.sparse_weights = padded;
    }
    pl.rows_contig = csr && contiguous;
    pl.row_desc.clear();
    if (pl.rows_contig)
        for (size_t r = 0; r < pl.tab.n_bins; ++r) {
            const int e0 = pl.tab.row_ptr[r], cnt = pl.tab.row_ptr[r + 1] - e0;
            pl.row_desc.insert(pl.row_desc.end(), {e0, cnt, cnt ? pl.tab.col[e0] : 0, 0});
        }
    // dense mappings: the column range outside which every row is exactly zero (chroma: bins outside [f_min, f_max]),
    // and, for chroma, the matrix transposed to [bin][12] for the chunked row sums of r2c_fused_pow2
    pl.dense_c0 = 0;
    pl.dense_c1 = static_cast<int>(pl.tab.out_len);
    pl.dense_t.clear();
    if (d.mapping == SGX_MAP_CHROMA) {
        const size_t nb = pl.tab.n_bins, ol = pl.tab.out_len;
        auto col_zero = [&](size_t k) { for (size_t r = 0; r < nb; ++r) if (pl.tab.dense[r * ol + k] != 0.0) return false; return true; };
        size_t lo = 0, hi = ol;
        while (lo < hi && col_zero(lo)) ++lo;
        while (hi > lo && col_zero(hi - 1)) --hi;
        pl.dense_c0 = static_cast<int>(lo);
        pl.dense_c1 = static_cast<int>(hi);
        pl.dense_t.resize(ol * nb);
        for (size_t r = 0; r < nb; ++r)
            for (size_t k = 0; k < ol; ++k) pl.dense_t[k * nb + r] = pl.tab.dense[r * ol + k];
    }
    build_lane_rows(pl);
    pl.fast400_sparse = pl.fast400 && csr && contiguous && fast400_sparse_fits(pl.sparse_quads, pl.sparse_weights);
    build_tc_blob(pl);
    pl.fast400_tc = pl.tc_steps > 0;
    {
        static const bool dense_tc_off = std::getenv("SGX_DENSE_TC") && std::atoi(std::getenv("SGX_DENSE_TC")) == 0;
        const int rows = dense_tc_max_rows(static_cast<int>(pl.tab.out_len));
        pl.dense_block_rows = rows;
        pl.dense_split = !dense_tc_off && !pl.f64 && d.mapping == SGX_MAP_ERB && d.output == SGX_OUT_SPECTROGRAM && rows >= 16 &&
                         !pl.fast400_tc && !pl.tab.dense.empty();
    }
    pl.fast400_tm = pl.fast400 && csr && contiguous && d.output == SGX_OUT_SPECTROGRAM && !pl.wofs_tm.empty() &&
                    fast400_tm_fits(pl.sparse_quads, pl.tm_weights);
    // fused mfcc() on the n400 family: the log-mel spectrogram by the TMEM-exchange kernel, the DCT-II on the tensor cores
    pl.mfcc_split = pl.fast400 && csr && contiguous && d.output == SGX_OUT_MFCC && !pl.wofs_tm.empty() && mfcc_tc_enabled() &&
                    fast400_tm_fits(pl.sparse_quads, pl.tm_weights) && mfcc_tc_supported(static_cast<int>(pl.tab.n_bins), static_cast<int>(d.n_mfcc));
    pl.kernel_name = pl.fast400 ? "r2c_fused_n400" : pl.pow2 ? "r2c_fused_pow2" : pl.mixed ? "r2c_fused_mixed" : "r2c_fused_generic";
    // folded DCT basis for the fused MFCC epilogue: B[c][n-1-i] = (-1)^c B[c][i] -> half basis, tasks of 4 coefficients of
    // one parity: [task][i < n/2][4], even-coefficient tasks first
    pl.dct_folded.clear();
    pl.dct_tasks = 0;
    const size_t dct_tasks_wanted = ((d.n_mfcc + 1) / 2 + 3) / 4 + (d.n_mfcc / 2 + 3) / 4;
    if (pl.fast400_sparse && d.output == SGX_OUT_MFCC && pl.tab.n_bins % 2 == 0 &&
        pl.tab.n_bins * 32 + dct_tasks_wanted * (pl.tab.n_bins / 2) * 4 <= static_cast<size_t>(fast400_max_scratch_rows()) * 32) {
        const size_t n = pl.tab.n_bins, half = n / 2, nm = d.n_mfcc;
        const size_t ge = ((nm + 1) / 2 + 3) / 4, go = (nm / 2 + 3) / 4;
        pl.dct_tasks = static_cast<int>(ge + go);
        pl.dct_folded.assign((ge + go) * half * 4, 0.0);
        for (size_t task = 0; task < ge + go; ++task) {
            const bool odd = task >= ge;
            const size_t g = odd ? task - ge : task;
            for (size_t k = 0; k < 4; ++k) {
                const size_t c = (odd ? 1 : 0) + 2 * (4 * g + k);
                if (c >= nm) continue;
                for (size_t i = 0; i < half; ++i) pl.dct_folded[(task * half + i) * 4 + k] = pl.tab.dct[c * n + i];
            }
        }
    }
}

void choose_generic_geometry(sgx_plan &pl) {
    const size_t es = pl.esize;
    const size_t out_len = pl.tab.out_len;
    size_t fs = static_cast<size_t>(pl.L) + 1;
    if (fs % 2 == 0) fs += 1;                                  // odd complex stride -> conflict-free frame-major reads
    size_t ts = std::max(out_len, pl.tab.n_bins);
    if (ts % 2 == 0) ts += 1;
    const size_t per_frame_cplx = std::max(fs, (ts + 1) / 2);  // complex elements per frame per buffer
    const size_t per_frame_bytes = 2 * per_frame_cplx * 2 * es;   // two ping-pong buffers
    size_t ft = (110 * 1024) / per_frame_bytes;
    if (ft < 2) ft = (200 * 1024) / per_frame_bytes;
    if (ft < 1) backend("n_fft too large for the CUDA plan (one frame exceeds shared memory)");
    if (ft > 32) ft = 32;
    pl.FT = static_cast<int>(ft);
    pl.tile_stride = static_cast<int>(ts);
    pl.buf_elems = static_cast<int>(per_frame_cplx * ft);
    pl.frame_stride = static_cast<int>(fs);
    pl.smem_bytes = per_frame_bytes * ft;
}

// Resolve the device and upload the per-plan tables once. Host-only queries (shape, axes, window, filterbank) work
// without a GPU; anything that computes does not.
void ensure_device(sgx_plan &pl) {
    if (pl.on_device) return;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        backend("no CUDA device available (this library has no CPU fallback)");
    }
    int dev = pl.device;
    if (dev < 0) ck(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev >= ndev) invalid("device ordinal out of range");
    pl.device = dev;
    DeviceGuard g(dev);
    pl.d_window = upload(pl.tab.window, pl.f64);
    pl.d_tw = upload(twiddle_table(static_cast<size_t>(pl.L), static_cast<size_t>(pl.L)), pl.f64);
    if (pl.even) pl.d_post = upload(twiddle_table(pl.desc.n_fft, static_cast<size_t>(pl.L) + 1), pl.f64);
    pl.d_row_ptr = upload_int(pl.tab.row_ptr);
    pl.d_col = upload_int(pl.tab.col);
    pl.d_wofs = upload_int(pl.wofs);
    pl.d_wofs_tm = upload_int(pl.wofs_tm);
    pl.wide_rows.clear();
    if (pl.pow2 && pl.pow2_ft >= 16 && pl.rows_contig && !pl.row_desc.empty() && pl.desc.output == SGX_OUT_SPECTROGRAM &&
        (pl.desc.mapping == SGX_MAP_MEL || pl.desc.mapping == SGX_MAP_LOGHZ)) {
        pl.wide_rows = pl.row_desc;                             // int4 {first entry, count, first column, 0} per row
        const size_t at = pl.wide_rows.size(), nnz = pl.tab.val.size();
        const size_t words = (nnz * pl.esize + 15) / 16 * 4;   // CSR values as T, padded to 16 bytes
        pl.wide_rows.resize(at + words, 0);
        if (pl.f64) for (size_t i = 0; i < nnz; ++i) std::memcpy(reinterpret_cast<char *>(&pl.wide_rows[at]) + 8 * i, &pl.tab.val[i], 8);
        else for (size_t i = 0; i < nnz; ++i) { const float v = static_cast<float>(pl.tab.val[i]); std::memcpy(&pl.wide_rows[at + i], &v, 4); }
        pl.d_wide_rows = upload_int(pl.wide_rows);
    }
    if (pl.mixed && !pl.f64) pl.d_mixed_q = upload_int(pl.mixed_q);
    pl.d_tc_blob = upload_int(pl.tc_blob);
    pl.d_lane_rows = upload_int(pl.lane_rows);
    pl.d_row_desc = upload_int(pl.row_desc);
    pl.d_lane_w = upload(pl.lane_w, pl.f64);
    pl.d_dense_t = upload(pl.dense_t, pl.f64);
    pl.d_val = upload(pl.tab.val, pl.f64);
    pl.d_dense = upload(pl.tab.dense, pl.f64);
    pl.d_dct = upload(pl.tab.dct, pl.f64);
    pl.d_lifter = upload(pl.tab.lifter, pl.f64);
    pl.d_dct_folded = upload(pl.dct_folded, pl.f64);
    if (pl.mfcc_split) pl.d_dct_tc = upload_floats(build_dct_tc_blob(pl.tab.n_bins, pl.desc.n_mfcc, pl.tab.dct, pl.tab.lifter));
    if (pl.dense_split) {
        const size_t nb = pl.tab.n_bins, ol = pl.tab.out_len;
        for (size_t r0 = 0; r0 < nb; r0 += static_cast<size_t>(pl.dense_block_rows)) {
            const size_t nr = std::min<size_t>(static_cast<size_t>(pl.dense_block_rows), nb - r0);
            const std::vector<double> block(pl.tab.dense.begin() + static_cast<std::ptrdiff_t>(r0 * ol), pl.tab.dense.begin() + static_cast<std::ptrdiff_t>((r0 + nr) * ol));
            pl.d_dense_tc.push_back(upload_floats(build_dct_tc_blob(ol, nr, block, std::vector<double>(nr, 1.0))));
        }
    }
    cudaDeviceProp prop;
    ck(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    pl.sm_count = prop.multiProcessorCount;
    pl.on_device = true;
}

void fill_params(const sgx_plan &pl, KParams &p) {
    std::memset(&p, 0, sizeof p);
    const sgx_plan_desc &d = pl.desc;
    p.n_fft = static_cast<int>(d.n_fft);
    p.hop = static_cast<int>(d.hop_size);
    p.pad = d.centre ? static_cast<int>(d.n_fft / 2) : 0;
    p.out_len = static_cast<int>(pl.tab.out_len);
    p.L = pl.L;
    p.even = pl.even;
    p.n_stages = static_cast<int>(pl.radix.size());
    for (size_t i = 0; i < pl.radix.size(); ++i) p.radix[i] = pl.radix[i];
    {
        unsigned cur = 1;
        for (size_t i = 0; i < pl.radix.size() && i < static_cast<size_t>(kFdStages); ++i) {
            p.fd_stage_B[i] = make_fastdiv(static_cast<unsigned>(pl.L / pl.radix[i]));
            p.fd_stage_cur[i] = make_fastdiv(cur);
            cur *= static_cast<unsigned>(pl.radix[i]);
        }
        p.fd_L = make_fastdiv(static_cast<unsigned>(pl.L));
        p.fd_out_len = make_fastdiv(static_cast<unsigned>(pl.tab.out_len));
        p.fd_r0 = make_fastdiv(pl.radix.empty() ? 1u : static_cast<unsigned>(pl.radix[0]));
    }
    p.window = pl.d_window; p.tw = pl.d_tw; p.post = pl.d_post;
    p.mapping = d.mapping;
    p.n_bins = static_cast<int>(pl.tab.n_bins);
    p.rows_contig = pl.rows_contig ? 1 : 0;
    p.row_desc = reinterpret_cast<const int4 *>(pl.d_row_desc);
    p.lane_rows = reinterpret_cast<const int4 *>(pl.d_lane_rows);
    p.lane_w = pl.d_lane_w;
    p.n_lane_slots = static_cast<int>(pl.lane_rows.size() / 4);
    p.row_ptr = pl.d_row_ptr; p.col = pl.d_col; p.val = pl.d_val; p.dense = pl.d_dense;
    p.chroma_norm = d.chroma_norm;
    p.dense_t = pl.d_dense_t;
    p.dense_c0 = pl.dense_c0;
    p.dense_c1 = pl.dense_c1;
    p.amp = d.amp;
    p.apply_db = (d.amp == SGX_AMP_DECIBELS && d.has_floor_db) ? 1 : 0;     // quirk F7: Decibels + None = raw power
    p.eps = d.has_floor_db ? std::pow(10.0, d.floor_db / 10.0) : 0.0;
    p.output = d.output;
    p.n_mfcc = static_cast<int>(d.n_mfcc);
    p.mfcc_row0 = (d.output == SGX_OUT_MFCC && !d.include_c0 && d.n_mfcc > 1) ? 1 : 0;
    p.dct = pl.d_dct; p.lifter = pl.d_lifter;
    p.dct_folded = pl.d_dct_folded; p.dct_tasks = pl.dct_tasks;
    p.FT = pl.FT; p.buf_elems = pl.buf_elems; p.frame_stride = pl.frame_stride; p.tile_stride = pl.tile_stride;
    p.fd_FT = make_fastdiv(static_cast<unsigned>(pl.FT));
}

// run frames [frame_begin, frame_begin+frames_todo) of n_clips device-resident clips
void run_device(sgx_plan &pl, const void *d_samples, size_t n_clips, size_t n_samples, size_t clip_stride, void *d_out,
                long long out_row_stride, long long out_clip_stride, long long frame_begin, long long frames_todo,
                cudaStream_t stream, long long pad_override = -1, bool as_linear_power = false) {
    KParams p;
    fill_params(pl, p);
    if (pad_override >= 0) p.pad = static_cast<int>(pad_override);
    if (as_linear_power) {
        // the plan's FFT family with the identity mapping and no scaling: |X|^2 [out_len][frames] (first half of a dense split)
        p.mapping = SGX_MAP_LINEAR;
        p.n_bins = static_cast<int>(pl.tab.out_len);
        p.amp = SGX_AMP_POWER;
        p.apply_db = 0;
        p.output = SGX_OUT_SPECTROGRAM;
        p.n_lane_slots = 0;
        p.rows_contig = 0;
    }
    p.samples = d_samples;
    p.n_samples = static_cast<long long>(n_samples);
    p.clip_stride = static_cast<long long>(clip_stride);
    p.frame_begin = frame_begin;
    p.frames_todo = frames_todo;
    p.out = d_out;
    p.out_row_stride = out_row_stride;
    p.out_clip_stride = out_clip_stride;
    p.out_frame_origin = frame_begin;
    const int tile_frames = pl.force_generic ? p.FT : pl.fast400 ? 32 : pl.pow2 ? pl.pow2_ft : pl.mixed ? mixed_tile_frames() : p.FT;
    p.tiles_per_clip = static_cast<int>((frames_todo + tile_frames - 1) / tile_frames);
    // mfcc() on the n400 family, dense outputs: log-mel spectrogram of a chunk of clips into plan scratch by r2c_fused_n400_tm,
    // then the DCT-II + lifter as a tcgen05 GEMM (dct2_lifter_tc). Two launches; only the log-mel tile (a sixth of the input's
    // bytes) makes a round trip through HBM / L2.
    const long long mfcc_rows = static_cast<long long>(pl.desc.n_mfcc) - p.mfcc_row0;
    if (pl.dense_split && !as_linear_power && !pl.force_generic && pl.tc_mode != 0 && out_row_stride == frames_todo &&
        out_clip_stride == static_cast<long long>(pl.tab.n_bins) * frames_todo) {
        const size_t ol = pl.tab.out_len, nb = pl.tab.n_bins;
        const size_t per_clip = ol * static_cast<size_t>(frames_todo) * sizeof(float);
        size_t chunk = std::max<size_t>(1, (size_t(1) << 30) / per_clip);
        chunk = std::min(chunk, n_clips);
        if (pl.logmel_cap < chunk * per_clip) {
            if (pl.d_logmel) { ck(cudaDeviceSynchronize(), "sync"); cudaFree(pl.d_logmel); pl.d_logmel = nullptr; pl.logmel_cap = 0; }
            ck(cudaMalloc(&pl.d_logmel, chunk * per_clip), "cudaMalloc(power scratch)");
            pl.logmel_cap = chunk * per_clip;
        }
        pl.scratch_acquire(stream);
        for (size_t c0 = 0; c0 < n_clips; c0 += chunk) {
            const size_t nc = std::min(chunk, n_clips - c0);
            run_device(pl, static_cast<const char *>(d_samples) + c0 * clip_stride * pl.esize, nc, n_samples, clip_stride, pl.d_logmel, frames_todo,
                       static_cast<long long>(ol) * frames_todo, frame_begin, frames_todo, stream, pad_override, true);
            size_t blk = 0;
            for (size_t r0 = 0; r0 < nb; r0 += static_cast<size_t>(pl.dense_block_rows), ++blk) {
                const size_t nr = std::min<size_t>(static_cast<size_t>(pl.dense_block_rows), nb - r0);
                ck(launch_dense_tc(static_cast<const float *>(pl.d_logmel), frames_todo, static_cast<float *>(d_out) + c0 * static_cast<size_t>(out_clip_stride),
                                   out_clip_stride, static_cast<long long>(nc), static_cast<int>(ol), frames_todo, static_cast<int>(r0), static_cast<int>(nr),
                                   pl.d_dense_tc[blk], p.amp, p.apply_db, static_cast<float>(p.eps), pl.sm_count, stream),
                   "kernel launch (dense_rows_tc)");
                pl.last_launches += 1;
            }
        }
        pl.scratch_release(stream);
        return;
    }
    if (pl.mfcc_split && !as_linear_power && !pl.force_generic && pl.tm_mode != 0 && out_row_stride == frames_todo && out_clip_stride == mfcc_rows * frames_todo) {
        const size_t nb = pl.tab.n_bins;
        const size_t per_clip = nb * static_cast<size_t>(frames_todo) * sizeof(float);
        size_t chunk = std::max<size_t>(1, (size_t(1) << 30) / per_clip);
        chunk = std::min(chunk, n_clips);
        if (pl.logmel_cap < chunk * per_clip) {
            if (pl.d_logmel) { ck(cudaDeviceSynchronize(), "sync"); cudaFree(pl.d_logmel); pl.d_logmel = nullptr; pl.logmel_cap = 0; }
            ck(cudaMalloc(&pl.d_logmel, chunk * per_clip), "cudaMalloc(log-mel scratch)");
            pl.logmel_cap = chunk * per_clip;
        }
        pl.scratch_acquire(stream);
        for (size_t c0 = 0; c0 < n_clips; c0 += chunk) {
            const size_t nc = std::min(chunk, n_clips - c0);
            KParams q = p;
            q.n_clips = static_cast<int>(nc);
            q.samples = static_cast<const char *>(d_samples) + c0 * clip_stride * pl.esize;
            q.output = SGX_OUT_SPECTROGRAM;
            q.out = pl.d_logmel;
            q.out_row_stride = frames_todo;
            q.out_clip_stride = static_cast<long long>(nb) * frames_todo;
            q.vec_ok = (reinterpret_cast<uintptr_t>(q.samples) % 8 == 0 && clip_stride % 2 == 0) ? 1 : 0;
            q.sched = pl.d_wofs_tm;
            ck(launch_fast400_tm(q, pl.window_f32.data(), pl.sparse_quads, pl.tm_weights, pl.tm_warps, pl.sm_count, stream), "kernel launch (r2c_fused_n400_tm)");
            ck(launch_mfcc_tc(static_cast<const float *>(pl.d_logmel), frames_todo, static_cast<float *>(d_out) + c0 * static_cast<size_t>(out_clip_stride),
                              static_cast<long long>(nc), static_cast<int>(nb), frames_todo, static_cast<int>(pl.desc.n_mfcc), p.mfcc_row0,
                              pl.d_dct_tc, pl.sm_count, stream), "kernel launch (dct2_lifter_tc)");
            pl.last_launches += 2;
        }
        pl.scratch_release(stream);
        return;
    }
    // the grid is limited to 2^31-1 CTAs: split very large batches
    const long long max_clips = std::max<long long>(1, 2000000000LL / std::max(1, p.tiles_per_clip));
    for (size_t c0 = 0; c0 < n_clips; c0 += static_cast<size_t>(max_clips)) {
        const size_t nc = std::min<size_t>(static_cast<size_t>(max_clips), n_clips - c0);
        KParams q = p;
        q.n_clips = static_cast<int>(nc);
        q.samples = static_cast<const char *>(d_samples) + c0 * clip_stride * pl.esize;
        q.out = static_cast<char *>(d_out) + c0 * static_cast<size_t>(out_clip_stride) * pl.esize *
                                                 (pl.desc.output == SGX_OUT_COMPLEX_STFT ? 2 : 1);
        if (pl.fast400 && !pl.force_generic) {
            // 8-byte vector loads need an 8-byte aligned base and an even clip stride
            q.vec_ok = (reinterpret_cast<uintptr_t>(q.samples) % 8 == 0 && clip_stride % 2 == 0) ? 1 : 0;
            if (use_tm(pl) && !as_linear_power) {
                q.sched = pl.d_wofs_tm;
                ck(launch_fast400_tm(q, pl.window_f32.data(), pl.sparse_quads, pl.tm_weights, pl.tm_warps, pl.sm_count, stream), "kernel launch (r2c_fused_n400_tm)");
                pl.last_launches += 1;
                continue;
            }
            if (use_tc(pl) && !as_linear_power) {
                q.sched = pl.d_tc_blob;
                ck(launch_fast400_tc(q, pl.window_f32.data(), pl.tc_steps, pl.tc_rounds, pl.tc_b_floats, pl.sm_count, stream), "kernel launch (r2c_fused_n400_tc)");
                pl.last_launches += 1;
                continue;
            }
            const bool sparse = pl.fast400_sparse && !as_linear_power;
            q.sched = sparse ? pl.d_wofs : nullptr;
            ck(launch_fast400(q, pl.window_f32.data(), sparse, pl.sparse_quads, pl.sparse_weights, pl.sm_count, stream), "kernel launch (r2c_fused_n400)");
        } else if (pl.pow2 && !pl.force_generic) {
            q.FT = pl.pow2_ft;
            q.fd_FT = make_fastdiv(static_cast<unsigned>(q.FT));
            q.frame_stride = pl.pow2_frame_stride;
            q.tile_stride = pl.pow2_tile_stride;
            // vector loads of (x[2n], x[2n+1]) pairs need pair-aligned addresses: aligned base, even stride, even hop and pad
            q.vec_ok = (reinterpret_cast<uintptr_t>(q.samples) % (2 * pl.esize) == 0 && clip_stride % 2 == 0 &&
                           pl.desc.hop_size % 2 == 0 && q.pad % 2 == 0) ? 1 : 0;
            // cp.async.bulk staging of the signal tile (measured, profiles/r2_pow2_experiments.md): opt-in through SGX_POW2_BULK=1
            static const bool bulk_env = std::getenv("SGX_POW2_BULK") && std::atoi(std::getenv("SGX_POW2_BULK")) != 0;
            size_t smem = pl.pow2_smem;
            const size_t extra = bulk_env && pl.desc.output != SGX_OUT_COMPLEX_STFT ? pow2_bulk_stage_bytes(pl.desc.n_fft, pl.desc.hop_size, pl.f64) : 0;
            if (extra && q.vec_ok && reinterpret_cast<uintptr_t>(q.samples) % 16 == 0 && clip_stride % 4 == 0 && q.pad % 4 == 0 &&
                smem + extra <= 200 * 1024) {
                q.vec_ok |= 2;
                smem += extra;
            }
            // rows-per-thread epilogue (FT <= 8, sparse mapping): stage the lane-major weights in shared memory when that does not
            // cost a resident CTA (kernel_pow2.cu); SGX_POW2_STAGE_W=0 keeps them in global memory
            static const bool stage_w_off = std::getenv("SGX_POW2_STAGE_W") && std::atoi(std::getenv("SGX_POW2_STAGE_W")) == 0;
            q.lane_w_smem = 0;
            q.lane_w_bytes = 0;
            // (measured: n_fft 2048 -9 .. -11 %, n_fft 1024 +1.7 % -- its rows are half as long -- so from 2048 points on)
            if (!stage_w_off && !(q.vec_ok & 2) && !as_linear_power && pl.d_wide_rows) {
                // wide tiles (n_fft 256 / 512): row descriptors + CSR values of the lane = frame sparse rows
                const size_t wbytes = pl.wide_rows.size() * sizeof(int), base = (smem + 15) & ~size_t(15);
                const size_t budget = size_t(227) * 1024 / static_cast<size_t>(std::max(1, pow2_min_blocks(pl.desc.n_fft, pl.f64))) - 1024;
                if (base + wbytes <= budget) {
                    q.lane_w = pl.d_wide_rows;
                    q.lane_w_smem = static_cast<int>(base);
                    q.lane_w_bytes = static_cast<int>(wbytes);
                    smem = base + wbytes;
                }
            } else
            if (!stage_w_off && !(q.vec_ok & 2) && q.FT <= 8 && pl.desc.n_fft >= 2048 && pl.desc.output == SGX_OUT_SPECTROGRAM && pl.desc.mapping != SGX_MAP_LINEAR &&
                q.n_lane_slots > 0 && !pl.lane_w.empty()) {
                const size_t wbytes = pl.lane_w.size() * pl.esize, base = (smem + 15) & ~size_t(15);
                const size_t budget = size_t(227) * 1024 / static_cast<size_t>(std::max(1, pow2_min_blocks(pl.desc.n_fft, pl.f64))) - 1024;
                if (wbytes % 16 == 0 && base + wbytes <= budget) {
                    q.lane_w_smem = static_cast<int>(base);
                    q.lane_w_bytes = static_cast<int>(wbytes);
                    smem = base + wbytes;
                }
            }
            {
                // L2 prefetch of the successor CTA's samples: measured -1.5 % on the f64 multichannel batch (two CTAs per SM, 40 KB
                // of samples each), +0.5 .. 2 % on every f32 size (four CTAs per SM already cover the first load) -> f64 only
                static const int l2pf = std::getenv("SGX_POW2_L2PF") ? std::atoi(std::getenv("SGX_POW2_L2PF")) : -1;
                const int ahead = l2pf >= 0 ? l2pf : (pl.f64 ? 1 : 0);
                q.l2_ahead = ahead * pl.sm_count * std::max(1, pow2_min_blocks(pl.desc.n_fft, pl.f64));
            }
            ck(launch_pow2(q, pl.f64, smem, stream), "kernel launch (r2c_fused_pow2)");
        } else if (pl.mixed && !pl.force_generic) {
            q.FT = mixed_tile_frames();
            q.fd_FT = make_fastdiv(static_cast<unsigned>(q.FT));
            q.vec_ok = (reinterpret_cast<uintptr_t>(q.samples) % (2 * pl.esize) == 0 && clip_stride % 2 == 0 &&
                           pl.desc.hop_size % 2 == 0 && q.pad % 2 == 0) ? 1 : 0;
            // quad epilogue (sparse rows of the n400 family) for f32 mel / loghz spectrogram outputs when its table fits behind the tile
            // without costing the second resident CTA; SGX_MIXED_QUADS=0 keeps the lane = frame epilogue
            static const bool quads_off = std::getenv("SGX_MIXED_QUADS") && std::atoi(std::getenv("SGX_MIXED_QUADS")) == 0;
            q.lane_w_smem = 0;
            q.lane_w_bytes = 0;
            q.sched = nullptr;
            if (!quads_off && !as_linear_power && pl.d_mixed_q && !pl.f64 && !pl.eps_denormal && pl.desc.output == SGX_OUT_SPECTROGRAM &&
                (pl.desc.mapping == SGX_MAP_MEL || pl.desc.mapping == SGX_MAP_LOGHZ) && pl.rows_contig) {
                const size_t base = (mixed_smem_bytes(pl.desc.n_fft, false) + 15) & ~size_t(15), bytes = pl.mixed_q.size() * sizeof(int);
                const size_t ctas = (base + 1024) * 2 <= size_t(228) * 1024 ? 2 : 1;
                if ((base + bytes + 1024) * ctas <= size_t(228) * 1024) {
                    q.lane_w_smem = static_cast<int>(base);
                    q.lane_w_bytes = static_cast<int>(bytes);
                    q.sched = pl.d_mixed_q;
                }
            }
            ck(launch_mixed(q, pl.f64, stream), "kernel launch (r2c_fused_mixed)");
        } else {
            ck(launch_generic(q, pl.f64, pl.smem_bytes, stream), "kernel launch (r2c_fused_generic)");
        }
        pl.last_launches += 1;
    }
}

void ensure_slot(sgx_plan::Slot &s, size_t in_bytes, size_t out_bytes) {
    if (!s.s) ck(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking), "cudaStreamCreate");
    if (s.in_cap < in_bytes) {
        if (s.d_in) { ck(cudaStreamSynchronize(s.s), "sync"); cudaFree(s.d_in); s.d_in = nullptr; }
        ck(cudaMalloc(&s.d_in, in_bytes), "cudaMalloc(staging in)");
        s.in_cap = in_bytes;
    }
    if (s.out_cap < out_bytes) {
        if (s.d_out) { ck(cudaStreamSynchronize(s.s), "sync"); cudaFree(s.d_out); s.d_out = nullptr; }
        ck(cudaMalloc(&s.d_out, out_bytes), "cudaMalloc(staging out)");
        s.out_cap = out_bytes;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char *sgx_last_error_message(void) { return g_err_msg.c_str(); }
void sgx_last_dimension_mismatch(size_t *expected, size_t *got) {
    if (expected) *expected = g_exp;
    if (got) *got = g_got;
}
const char *sgx_version(void) { return "sgx_b200 0.1.0 sm_100a"; }

sgx_status sgx_plan_create(const sgx_plan_desc *desc, sgx_plan **out_plan) {
    return guarded([&] {
        if (!desc || !out_plan) invalid("null argument");
        *out_plan = nullptr;
        validate_desc(*desc);
        std::unique_ptr<sgx_plan> pl(new sgx_plan());
        pl->desc = *desc;
        if (desc->window == SGX_WIN_CUSTOM) {
            pl->custom.assign(desc->custom_window, desc->custom_window + desc->custom_window_len);
            pl->desc.custom_window = pl->custom.data();
        } else {
            pl->desc.custom_window = nullptr;
            pl->desc.custom_window_len = 0;
        }
        pl->f64 = desc->dtype == SGX_F64;
        pl->esize = pl->f64 ? 8 : 4;
        build_tables(pl->desc, pl->tab);
        pl->rows = desc->output == SGX_OUT_COMPLEX_STFT ? pl->tab.out_len
                 : desc->output == SGX_OUT_MFCC ? desc->n_mfcc - ((!desc->include_c0 && desc->n_mfcc > 1) ? 1 : 0)
                                                : pl->tab.n_bins;
        factorise(*pl);
        choose_generic_geometry(*pl);
        select_family(*pl);

        pl->device = desc->device;   // device tables are uploaded on first use (ensure_device)
        *out_plan = pl.release();
    });
}

sgx_status sgx_plan_destroy(sgx_plan *plan) {
    return guarded([&] { delete plan; });
}

sgx_status sgx_plan_output_shape(const sgx_plan *plan, size_t n_samples, size_t *n_rows, size_t *n_frames) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        if (n_samples == 0) invalid("signal length must be non-zero");
        if (n_rows) *n_rows = plan->rows;
        if (n_frames) *n_frames = frame_count(n_samples, plan->desc.n_fft, plan->desc.hop_size, plan->desc.centre != 0);
    });
}

sgx_status sgx_plan_axes(const sgx_plan *plan, size_t n_frames, double *freqs, double *times) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        if (freqs) std::memcpy(freqs, plan->tab.freq_axis.data(), sizeof(double) * plan->tab.freq_axis.size());
        if (times) {
            // frame_period_seconds = hop / sr (:4268-4271); times[i] = i * dt (:2128-2139)
            const double dt = static_cast<double>(plan->desc.hop_size) / plan->desc.sample_rate_hz;
            for (size_t i = 0; i < n_frames; ++i) times[i] = static_cast<double>(i) * dt;
        }
    });
}

sgx_status sgx_plan_window(const sgx_plan *plan, void *out_host) {
    return guarded([&] {
        if (!plan || !out_host) invalid("null argument");
        const size_t n = plan->desc.n_fft;
        if (plan->f64) for (size_t i = 0; i < n; ++i) static_cast<double *>(out_host)[i] = plan->tab.window[i];
        else for (size_t i = 0; i < n; ++i) static_cast<float *>(out_host)[i] = static_cast<float>(plan->tab.window[i]);
    });
}

sgx_status sgx_plan_filterbank(const sgx_plan *plan, double *dense_out, size_t *nnz) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        const HostTables &t = plan->tab;
        const size_t nb = t.n_bins, ol = t.out_len;
        const bool dense_map = plan->desc.mapping == SGX_MAP_ERB || plan->desc.mapping == SGX_MAP_CHROMA;
        if (nnz) *nnz = dense_map ? nb * ol : plan->desc.mapping == SGX_MAP_LINEAR ? ol : t.val.size();
        if (!dense_out) return;
        std::fill(dense_out, dense_out + nb * ol, 0.0);
        if (plan->desc.mapping == SGX_MAP_LINEAR) {
            for (size_t r = 0; r < nb; ++r) dense_out[r * ol + r] = 1.0;
        } else if (dense_map) {
            std::memcpy(dense_out, t.dense.data(), sizeof(double) * nb * ol);
        } else {
            for (size_t r = 0; r < nb; ++r)
                for (int e = t.row_ptr[r]; e < t.row_ptr[r + 1]; ++e) dense_out[r * ol + t.col[e]] = t.val[e];
        }
    });
}

const char *sgx_plan_kernel_name(const sgx_plan *plan) {
    if (!plan) return "";
    if (plan->force_generic) return "r2c_fused_generic";
    if (plan->mfcc_split && plan->tm_mode != 0) return "r2c_fused_n400_tm+dct2_lifter_tc";
    if (plan->dense_split && plan->tc_mode != 0) {
        static thread_local std::string name;
        name = plan->kernel_name + "+dense_rows_tc";
        return name.c_str();
    }
    if (use_tm(*plan)) return "r2c_fused_n400_tm";
    if (use_tc(*plan)) return "r2c_fused_n400_tc";
    return plan->kernel_name.c_str();
}
size_t sgx_plan_last_launch_count(const sgx_plan *plan) { return plan ? plan->last_launches : 0; }
sgx_status sgx_plan_force_generic(sgx_plan *plan, int force) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        plan->force_generic = force != 0;
    });
}

sgx_status sgx_plan_set_tmem_exchange(sgx_plan *plan, int enable) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        plan->tm_mode = enable < 0 ? -1 : (enable != 0 ? 1 : 0);
    });
}

sgx_status sgx_plan_set_tensor_cores(sgx_plan *plan, int enable) {
    return guarded([&] {
        if (!plan) invalid("null plan");
        plan->tc_mode = enable < 0 ? -1 : (enable != 0 ? 1 : 0);
    });
}

sgx_status sgx_plan_compute_batch(sgx_plan *plan, const void *samples, size_t n_clips, size_t n_samples,
                                  size_t clip_stride, void *out, size_t out_rows, size_t out_cols,
                                  size_t out_clip_stride, void *cuda_stream) {
    return guarded([&] {
        if (!plan || !samples || !out) invalid("null argument");
        if (n_samples == 0) invalid("samples must be non-empty");                 // NonEmptySlice
        if (n_clips == 0) invalid("n_clips must be non-zero");
        if (clip_stride < n_samples) invalid("clip_stride must be >= n_samples");
        sgx_plan &pl = *plan;
        const size_t n_frames = frame_count(n_samples, pl.desc.n_fft, pl.desc.hop_size, pl.desc.centre != 0);
        if (out_rows != pl.rows) mismatch(pl.rows, out_rows);                     // rows first (:423-428)
        if (out_cols != n_frames) mismatch(n_frames, out_cols);                   // then columns (:429-434)
        if (out_clip_stride == 0) out_clip_stride = out_rows * out_cols;
        if (out_clip_stride < out_rows * out_cols) invalid("out_clip_stride must be >= out_rows*out_cols");
        ensure_device(pl);
        DeviceGuard g(pl.device);
        pl.last_launches = 0;
        const size_t oes = pl.esize * (pl.desc.output == SGX_OUT_COMPLEX_STFT ? 2 : 1);
        const PtrKind ki = ptr_kind(samples), ko = ptr_kind(out);
        if (ki != ko) invalid("samples and out must both be host pointers or both be device pointers");
        if (ki == PtrKind::Device) {
            run_device(pl, samples, n_clips, n_samples, clip_stride, out, static_cast<long long>(n_frames),
                       static_cast<long long>(out_clip_stride), 0, static_cast<long long>(n_frames),
                       static_cast<cudaStream_t>(cuda_stream));
            return;
        }
        // host pointers: chunk the clips, pipeline H2D / compute / D2H over kStagingSlots private streams
        const size_t in_clip_bytes = n_samples * pl.esize;
        const size_t out_clip_bytes = out_rows * out_cols * oes;
        size_t chunk = std::max<size_t>(1, (size_t(64) << 20) / std::max(in_clip_bytes, out_clip_bytes));
        chunk = std::min(chunk, n_clips);
        int si = 0;
        for (size_t c0 = 0; c0 < n_clips; c0 += chunk, si = (si + 1) % kStagingSlots) {
            const size_t nc = std::min(chunk, n_clips - c0);
            sgx_plan::Slot &s = pl.slot[si];
            ensure_slot(s, chunk * in_clip_bytes, chunk * out_clip_bytes);
            const char *src = static_cast<const char *>(samples) + c0 * clip_stride * pl.esize;
            if (clip_stride == n_samples)      // contiguous clips: one linear copy (faster than the strided 2-D path)
                ck(cudaMemcpyAsync(s.d_in, src, nc * in_clip_bytes, cudaMemcpyHostToDevice, s.s), "H2D copy");
            else
                ck(cudaMemcpy2DAsync(s.d_in, in_clip_bytes, src, clip_stride * pl.esize, in_clip_bytes, nc, cudaMemcpyHostToDevice, s.s), "H2D copy");
            run_device(pl, s.d_in, nc, n_samples, n_samples, s.d_out, static_cast<long long>(n_frames),
                       static_cast<long long>(out_rows * out_cols), 0, static_cast<long long>(n_frames), s.s);
            char *dst = static_cast<char *>(out) + c0 * out_clip_stride * oes;
            if (out_clip_stride == out_rows * out_cols)
                ck(cudaMemcpyAsync(dst, s.d_out, nc * out_clip_bytes, cudaMemcpyDeviceToHost, s.s), "D2H copy");
            else
                ck(cudaMemcpy2DAsync(dst, out_clip_stride * oes, s.d_out, out_clip_bytes, out_clip_bytes, nc, cudaMemcpyDeviceToHost, s.s), "D2H copy");
        }
        for (auto &s : pl.slot) if (s.s) ck(cudaStreamSynchronize(s.s), "cudaStreamSynchronize");
    });
}

sgx_status sgx_plan_compute_frame(sgx_plan *plan, const void *samples, size_t n_samples, size_t frame_idx, void *out,
                                  void *cuda_stream) {
    return guarded([&] {
        if (!plan || !samples || !out) invalid("null argument");
        if (n_samples == 0) invalid("samples must be non-empty");
        sgx_plan &pl = *plan;
        if (frame_idx > (size_t(1) << 62) / pl.desc.hop_size) invalid("frame index overflow");     // checked_mul (:1262-1264)
        ensure_device(pl);
        DeviceGuard g(pl.device);
        pl.last_launches = 0;
        const size_t oes = pl.esize * (pl.desc.output == SGX_OUT_COMPLEX_STFT ? 2 : 1);
        const PtrKind ki = ptr_kind(samples), ko = ptr_kind(out);
        if (ki != ko) invalid("samples and out must both be host pointers or both be device pointers");
        if (ki == PtrKind::Device) {
            run_device(pl, samples, 1, n_samples, n_samples, out, 1, static_cast<long long>(pl.rows),
                       static_cast<long long>(frame_idx), 1, static_cast<cudaStream_t>(cuda_stream));
            return;
        }
        // host pointers: stage only the span [lo, hi) of samples this frame touches and shift the padding so that
        // kernel index (0*hop - pad' + i) addresses the staged span; everything outside it is zero padding anyway.
        sgx_plan::Slot &s = pl.slot[0];
        const long long pad = pl.desc.centre ? static_cast<long long>(pl.desc.n_fft / 2) : 0;
        const long long lo_raw = static_cast<long long>(frame_idx * pl.desc.hop_size) - pad;
        const long long lo = std::min<long long>(std::max<long long>(lo_raw, 0), static_cast<long long>(n_samples));
        const long long hi = std::min<long long>(std::max<long long>(lo_raw + static_cast<long long>(pl.desc.n_fft), lo),
                                                 static_cast<long long>(n_samples));
        const size_t span = static_cast<size_t>(hi - lo);
        ensure_slot(s, std::max<size_t>(span, 1) * pl.esize, pl.rows * oes);
        if (span) ck(cudaMemcpyAsync(s.d_in, static_cast<const char *>(samples) + static_cast<size_t>(lo) * pl.esize,
                                     span * pl.esize, cudaMemcpyHostToDevice, s.s), "H2D copy");
        run_device(pl, s.d_in, 1, span, std::max<size_t>(span, 1), s.d_out, 1, static_cast<long long>(pl.rows), 0, 1, s.s,
                   /*pad_override=*/lo - lo_raw);
        ck(cudaMemcpyAsync(out, s.d_out, pl.rows * oes, cudaMemcpyDeviceToHost, s.s), "D2H copy");
        ck(cudaStreamSynchronize(s.s), "cudaStreamSynchronize");
    });
}

sgx_status sgx_mfcc_from_log_mel(sgx_dtype dtype, const void *log_mel, size_t n_clips, size_t n_mels, size_t n_frames,
                                 size_t n_mfcc, int include_c0, size_t lifter, void *out, int device, void *cuda_stream) {
    return guarded([&] {
        if (!log_mel || !out) invalid("null argument");
        if (dtype != SGX_F32 && dtype != SGX_F64) invalid("dtype must be f32 or f64");
        if (n_mfcc == 0) invalid("n_mfcc must be non-zero");
        if (n_mfcc > n_mels) invalid("n_mfcc must be <= n_mels");                 // src/mfcc.rs:231-233
        if (n_clips == 0 || n_frames == 0) invalid("empty log-mel spectrogram");
        const bool f64 = dtype == SGX_F64;
        const size_t es = f64 ? 8 : 4;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            backend("no CUDA device available (this library has no CPU fallback)");
        }
        int dev = device;
        if (dev < 0) ck(cudaGetDevice(&dev), "cudaGetDevice");
        DeviceGuard g(dev);
        std::vector<double> basis, lift;
        build_dct(n_mfcc, n_mels, lifter, basis, lift);
        const int row0 = (!include_c0 && n_mfcc > 1) ? 1 : 0;
        const size_t rows = n_mfcc - row0;
        void *d_dct = upload(basis, f64), *d_lift = upload(lift, f64);
        const PtrKind ki = ptr_kind(log_mel), ko = ptr_kind(out);
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        float *d_blob = nullptr;
        int sm_count = 148;
        auto cleanup = [&] { cudaFree(d_dct); cudaFree(d_lift); if (d_blob) cudaFree(d_blob); };
        try {
            if (ki != ko) invalid("log_mel and out must both be host pointers or both be device pointers");
            // f32, n_mels <= 128, n_mfcc <= 64: the tensor-core form (dct2_lifter_tc); otherwise the CUDA-core kernel
            const bool tc_ok = !f64 && mfcc_tc_enabled() && mfcc_tc_supported(static_cast<int>(n_mels), static_cast<int>(n_mfcc));
            if (tc_ok) {
                d_blob = upload_floats(build_dct_tc_blob(n_mels, n_mfcc, basis, lift));
                ck(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
            }
            if (ki == PtrKind::Device) {
                if (tc_ok)
                    ck(launch_mfcc_tc(static_cast<const float *>(log_mel), static_cast<long long>(n_frames), static_cast<float *>(out), static_cast<long long>(n_clips), static_cast<int>(n_mels),
                                      static_cast<long long>(n_frames), static_cast<int>(n_mfcc), row0, d_blob, sm_count, st), "kernel launch (dct2_lifter_tc)");
                else
                ck(launch_mfcc(f64, log_mel, out, static_cast<long long>(n_clips), static_cast<int>(n_mels),
                               static_cast<long long>(n_frames), static_cast<int>(n_mfcc), row0, d_dct, d_lift, st), "kernel launch (dct2_lifter)");
                ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");   // tables are freed below
            } else {
                void *d_in = nullptr, *d_out = nullptr;
                const size_t ib = n_clips * n_mels * n_frames * es, ob = n_clips * rows * n_frames * es;
                ck(cudaMalloc(&d_in, ib), "cudaMalloc");
                if (cudaMalloc(&d_out, ob) != cudaSuccess) { cudaFree(d_in); backend("cudaMalloc failed"); }
                cudaError_t e = cudaMemcpyAsync(d_in, log_mel, ib, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess)
                    e = tc_ok ? launch_mfcc_tc(static_cast<const float *>(d_in), static_cast<long long>(n_frames), static_cast<float *>(d_out), static_cast<long long>(n_clips), static_cast<int>(n_mels),
                                               static_cast<long long>(n_frames), static_cast<int>(n_mfcc), row0, d_blob, sm_count, st)
                              : launch_mfcc(f64, d_in, d_out, static_cast<long long>(n_clips), static_cast<int>(n_mels),
                                            static_cast<long long>(n_frames), static_cast<int>(n_mfcc), row0, d_dct, d_lift, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                cudaFree(d_in); cudaFree(d_out);
                ck(e, "mfcc_from_log_mel");
            }
        } catch (...) { cleanup(); throw; }
        cleanup();
    });
}

sgx_status sgx_plan_compute_binaural(sgx_plan *plan, sgx_binaural_cue cue, const void *left, const void *right,
                                     size_t n_pairs, size_t n_samples, size_t clip_stride, double start_freq, double end_freq,
                                     size_t magphase_power, int wrapped, void *out, size_t out_bins, size_t out_frames,
                                     void *cuda_stream) {
    return guarded([&] {
        if (!plan || !left || !right || !out) invalid("null argument");
        sgx_plan &pl = *plan;
        if (pl.desc.output != SGX_OUT_COMPLEX_STFT) invalid("binaural cues need a complex STFT plan");
        if (cue < SGX_CUE_ITD || cue > SGX_CUE_ILR) invalid("unknown binaural cue");
        if (n_samples == 0) invalid("samples must be non-empty");
        if (n_pairs == 0) invalid("n_pairs must be non-zero");
        if (clip_stride < n_samples) invalid("clip_stride must be >= n_samples");
        if (magphase_power == 0 || magphase_power > 0xffffffffu) invalid("magphase_power must be a non-zero u32");
        const double bw = pl.desc.sample_rate_hz / static_cast<double>(pl.desc.n_fft);       // :476
        const double sb = std::round(start_freq / bw), eb = std::round(end_freq / bw);         // f64::round, half away from zero
        if (!(sb >= 0.0) || !(eb > sb)) invalid("Frequency range should have at least one bin");   // the reference's expect() (:549-550)
        const size_t start_bin = static_cast<size_t>(sb), stop_bin = static_cast<size_t>(eb);
        const size_t n_bins = pl.tab.out_len;
        if (stop_bin > n_bins) invalid("End frequency must be less than Nyquist frequency.");
        const size_t n_frames = frame_count(n_samples, pl.desc.n_fft, pl.desc.hop_size, pl.desc.centre != 0);
        if (out_bins != stop_bin - start_bin) mismatch(stop_bin - start_bin, out_bins);
        if (out_frames != n_frames) mismatch(n_frames, out_frames);
        ensure_device(pl);
        DeviceGuard g(pl.device);
        pl.last_launches = 0;
        const PtrKind kl = ptr_kind(left), kr = ptr_kind(right), ko = ptr_kind(out);
        if (kl != kr || kl != ko) invalid("left, right and out must all be host pointers or all be device pointers");
        const bool host = kl != PtrKind::Device;
        const size_t es = pl.esize;
        const size_t stft_bytes = n_bins * n_frames * 2 * es, cue_elems = out_bins * n_frames;
        // pairs per chunk: both channels' STFTs of a chunk stay under ~512 MB each
        size_t chunk = std::max<size_t>(1, (size_t(512) << 20) / stft_bytes);
        chunk = std::min(chunk, n_pairs);
        if (!(pl.pow2 && !pl.force_generic && pl.pow2_ft >= 2) && pl.pair_cap < chunk * stft_bytes) {
            for (void *&q : pl.d_pair) { if (q) { ck(cudaDeviceSynchronize(), "sync"); cudaFree(q); q = nullptr; } }
            pl.pair_cap = 0;
            ck(cudaMalloc(&pl.d_pair[0], chunk * stft_bytes), "cudaMalloc(binaural scratch)");
            ck(cudaMalloc(&pl.d_pair[1], chunk * stft_bytes), "cudaMalloc(binaural scratch)");
            pl.pair_cap = chunk * stft_bytes;
        }
        sgx_plan::Slot &s = pl.slot[0];
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        // Power-of-two plans with at least two frames per tile transform both channels of a pair in ONE launch and
        // compute the cue in its epilogue: the complex spectra never leave the SM.
        const bool fused = pl.pow2 && !pl.force_generic && pl.pow2_ft >= 2;
        if (fused) {
            if (host) {
                size_t hchunk = std::max<size_t>(1, (size_t(256) << 20) / std::max(2 * n_samples * es, cue_elems * es));
                hchunk = std::min(hchunk, n_pairs);
                ensure_slot(s, 2 * hchunk * n_samples * es, hchunk * cue_elems * es);
                st = s.s;
                chunk = hchunk;
            } else {
                chunk = n_pairs;
            }
            for (size_t p0 = 0; p0 < n_pairs; p0 += chunk) {
                const size_t np = std::min(chunk, n_pairs - p0);
                const void *ch[2] = {static_cast<const char *>(left) + p0 * clip_stride * es, static_cast<const char *>(right) + p0 * clip_stride * es};
                size_t stride = clip_stride;
                if (host) {
                    char *d_in = static_cast<char *>(s.d_in);
                    for (int c = 0; c < 2; ++c) {
                        ck(cudaMemcpy2DAsync(d_in + c * np * n_samples * es, n_samples * es, ch[c], clip_stride * es, n_samples * es, np,
                                             cudaMemcpyHostToDevice, st), "H2D copy");
                        ch[c] = d_in + c * np * n_samples * es;
                    }
                    stride = n_samples;
                }
                void *dst = host ? s.d_out : static_cast<char *>(out) + p0 * cue_elems * es;
                KParams q;
                fill_params(pl, q);
                q.samples = ch[0];
                q.samples_b = ch[1];
                q.n_samples = static_cast<long long>(n_samples);
                q.clip_stride = static_cast<long long>(stride);
                q.n_clips = static_cast<int>(np);
                q.frame_begin = 0;
                q.frames_todo = static_cast<long long>(n_frames);
                q.out = dst;
                q.out_row_stride = static_cast<long long>(n_frames);
                q.out_clip_stride = static_cast<long long>(cue_elems);
                q.out_frame_origin = 0;
                q.FT = pl.pow2_ft;
                q.fd_FT = make_fastdiv(static_cast<unsigned>(q.FT));
                q.frame_stride = pl.pow2_frame_stride;
                q.tile_stride = pl.pow2_tile_stride;
                const int hf = pl.pow2_ft / 2;
                q.tiles_per_clip = static_cast<int>((n_frames + hf - 1) / hf);
                q.vec_ok = (reinterpret_cast<uintptr_t>(ch[0]) % (2 * es) == 0 && reinterpret_cast<uintptr_t>(ch[1]) % (2 * es) == 0 &&
                            stride % 2 == 0 && pl.desc.hop_size % 2 == 0 && q.pad % 2 == 0) ? 1 : 0;
                q.cue = cue;
                q.cue_start_bin = static_cast<int>(start_bin);
                q.cue_band = static_cast<int>(out_bins);
                q.cue_power = static_cast<unsigned>(magphase_power);
                q.cue_wrapped = wrapped;
                q.cue_bin_width = bw;
                if (static_cast<long long>(np) * q.tiles_per_clip > 2147483647LL) backend("batch too large for one launch");
                ck(launch_pow2(q, pl.f64, pl.pow2_smem, st), "kernel launch (r2c_fused_pow2, stereo pair)");
                pl.last_launches += 1;
                if (host) ck(cudaMemcpyAsync(static_cast<char *>(out) + p0 * cue_elems * es, s.d_out, np * cue_elems * es, cudaMemcpyDeviceToHost, st), "D2H copy");
            }
            if (host) ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
            return;
        }
        if (host) {
            ensure_slot(s, 2 * chunk * n_samples * es, chunk * cue_elems * es);
            st = s.s;
        }
        pl.scratch_acquire(st);
        for (size_t p0 = 0; p0 < n_pairs; p0 += chunk) {
            const size_t np = std::min(chunk, n_pairs - p0);
            const void *ch[2] = {static_cast<const char *>(left) + p0 * clip_stride * es, static_cast<const char *>(right) + p0 * clip_stride * es};
            size_t stride = clip_stride;
            if (host) {
                char *d_in = static_cast<char *>(s.d_in);
                for (int c = 0; c < 2; ++c) {
                    ck(cudaMemcpy2DAsync(d_in + c * np * n_samples * es, n_samples * es, ch[c], clip_stride * es, n_samples * es, np,
                                         cudaMemcpyHostToDevice, st), "H2D copy");
                    ch[c] = d_in + c * np * n_samples * es;
                }
                stride = n_samples;
            }
            for (int c = 0; c < 2; ++c)
                run_device(pl, ch[c], np, n_samples, stride, pl.d_pair[c], static_cast<long long>(n_frames),
                           static_cast<long long>(n_bins * n_frames), 0, static_cast<long long>(n_frames), st);
            void *dst = host ? s.d_out : static_cast<char *>(out) + p0 * cue_elems * es;
            ck(launch_binaural(pl.f64, cue, pl.d_pair[0], pl.d_pair[1], dst, static_cast<long long>(np), static_cast<int>(n_bins),
                               static_cast<long long>(n_frames), static_cast<int>(start_bin), static_cast<int>(stop_bin), bw,
                               static_cast<unsigned>(magphase_power), wrapped, st), "kernel launch (binaural_cues)");
            pl.last_launches += 1;
            if (host) ck(cudaMemcpyAsync(static_cast<char *>(out) + p0 * cue_elems * es, s.d_out, np * cue_elems * es, cudaMemcpyDeviceToHost, st), "D2H copy");
        }
        pl.scratch_release(st);
        if (host) ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    });
}

sgx_status sgx_binaural_from_stft(sgx_dtype dtype, sgx_binaural_cue cue, const void *left, const void *right, size_t n_pairs,
                                  size_t n_bins, size_t n_frames, size_t start_bin, size_t stop_bin, double bin_width_hz,
                                  size_t magphase_power, int wrapped, void *out, int device, void *cuda_stream) {
    return guarded([&] {
        if (!left || !right || !out) invalid("null argument");
        if (dtype != SGX_F32 && dtype != SGX_F64) invalid("dtype must be f32 or f64");
        if (cue < SGX_CUE_ITD || cue > SGX_CUE_ILR) invalid("unknown binaural cue");
        if (n_pairs == 0 || n_bins == 0 || n_frames == 0) invalid("empty STFT");
        if (!(bin_width_hz > 0.0 && std::isfinite(bin_width_hz))) invalid("bin width must be finite and > 0");
        if (start_bin >= stop_bin) invalid("Frequency range should have at least one bin");    // the reference's expect() (:549-550)
        if (stop_bin > n_bins) mismatch(stop_bin, n_bins);
        if (magphase_power == 0 || magphase_power > 0xffffffffu) invalid("magphase_power must be a non-zero u32");
        const bool f64 = dtype == SGX_F64;
        const size_t es = f64 ? 8 : 4;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            backend("no CUDA device available (this library has no CPU fallback)");
        }
        int dev = device;
        if (dev < 0) ck(cudaGetDevice(&dev), "cudaGetDevice");
        DeviceGuard g(dev);
        const PtrKind kl = ptr_kind(left), kr = ptr_kind(right), ko = ptr_kind(out);
        if (kl != kr || kl != ko) invalid("left, right and out must all be host pointers or all be device pointers");
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        const long long np = static_cast<long long>(n_pairs), nf = static_cast<long long>(n_frames);
        if (kl == PtrKind::Device) {
            ck(launch_binaural(f64, cue, left, right, out, np, static_cast<int>(n_bins), nf, static_cast<int>(start_bin),
                               static_cast<int>(stop_bin), bin_width_hz, static_cast<unsigned>(magphase_power), wrapped, st),
               "kernel launch (binaural_cues)");
            return;
        }
        const size_t ib = n_pairs * n_bins * n_frames * 2 * es, ob = n_pairs * (stop_bin - start_bin) * n_frames * es;
        void *d_l = nullptr, *d_r = nullptr, *d_o = nullptr;
        cudaError_t e = cudaMalloc(&d_l, ib);
        if (e == cudaSuccess) e = cudaMalloc(&d_r, ib);
        if (e == cudaSuccess) e = cudaMalloc(&d_o, ob);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_l, left, ib, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_r, right, ib, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = launch_binaural(f64, cue, d_l, d_r, d_o, np, static_cast<int>(n_bins), nf, static_cast<int>(start_bin),
                                                  static_cast<int>(stop_bin), bin_width_hz, static_cast<unsigned>(magphase_power), wrapped, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, ob, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_l); cudaFree(d_r); cudaFree(d_o);
        ck(e, "binaural_from_stft");
    });
}

sgx_status sgx_chroma_filterbank(double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max, double *dense_out) {
    return guarded([&] {
        if (!dense_out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        validate_chroma(sample_rate_hz, tuning, f_min, f_max);
        std::vector<double> fb;
        build_chroma_filterbank(sample_rate_hz, n_fft, tuning, f_min, f_max, fb);
        std::memcpy(dense_out, fb.data(), sizeof(double) * fb.size());
    });
}

sgx_status sgx_chroma_from_spectrogram(sgx_dtype dtype, const void *spec, size_t n_clips, size_t n_bins, size_t n_frames,
                                       double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max,
                                       sgx_chroma_norm norm, void *out, int device, void *cuda_stream) {
    return guarded([&] {
        if (!spec || !out) invalid("null argument");
        if (dtype != SGX_F32 && dtype != SGX_F64) invalid("dtype must be f32 or f64");
        if (n_fft == 0) invalid("n_fft must be set");
        if (n_clips == 0 || n_frames == 0) invalid("empty spectrogram");
        if (n_bins != n_fft / 2 + 1) mismatch(n_fft / 2 + 1, n_bins);                 // src/chroma.rs:376-379
        validate_chroma(sample_rate_hz, tuning, f_min, f_max);
        if (norm < SGX_CHROMANORM_NONE || norm > SGX_CHROMANORM_MAX) invalid("unknown chroma normalisation");
        const bool f64 = dtype == SGX_F64;
        const size_t es = f64 ? 8 : 4;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            backend("no CUDA device available (this library has no CPU fallback)");
        }
        int dev = device;
        if (dev < 0) ck(cudaGetDevice(&dev), "cudaGetDevice");
        DeviceGuard g(dev);
        std::vector<double> fb, fbT(12 * n_bins);
        build_chroma_filterbank(sample_rate_hz, n_fft, tuning, f_min, f_max, fb);
        for (size_t c = 0; c < 12; ++c)
            for (size_t k = 0; k < n_bins; ++k) fbT[k * 12 + c] = fb[c * n_bins + k];
        auto bin_zero = [&](size_t k) { for (size_t c = 0; c < 12; ++c) if (fbT[k * 12 + c] != 0.0) return false; return true; };
        size_t k0 = 0, k1 = n_bins;
        while (k0 < k1 && bin_zero(k0)) ++k0;
        while (k1 > k0 && bin_zero(k1 - 1)) --k1;
        const PtrKind ki = ptr_kind(spec), ko = ptr_kind(out);
        if (ki != ko) invalid("spec and out must both be host pointers or both be device pointers");
        void *d_w = upload(fbT, f64);
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        void *d_in = nullptr, *d_out = nullptr;
        cudaError_t e = cudaSuccess;
        if (ki == PtrKind::Device) {
            e = launch_chroma(f64, spec, out, static_cast<long long>(n_clips), static_cast<int>(n_bins), static_cast<long long>(n_frames), d_w, norm, static_cast<int>(k0), static_cast<int>(k1), st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);                       // the table is freed below
        } else {
            const size_t ib = n_clips * n_bins * n_frames * es, ob = n_clips * 12 * n_frames * es;
            e = cudaMalloc(&d_in, ib);
            if (e == cudaSuccess) e = cudaMalloc(&d_out, ob);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, spec, ib, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = launch_chroma(f64, d_in, d_out, static_cast<long long>(n_clips), static_cast<int>(n_bins), static_cast<long long>(n_frames), d_w, norm, static_cast<int>(k0), static_cast<int>(k1), st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        cudaFree(d_in); cudaFree(d_out); cudaFree(d_w);
        ck(e, "chroma_from_spectrogram");
    });
}

namespace {
// frames of `nc` clips -> out (device pointers); apply_window = 0 and n_frames = 1 gives a bare irfft
void run_inverse(sgx_plan &pl, const void *d_stft, size_t nc, size_t n_frames, void *d_frames_out, int apply_window, cudaStream_t st) {
    KParams p;
    fill_params(pl, p);
    p.n_clips = static_cast<int>(nc);
    // the register-radix inverse stores (x[2m], x[2m+1]) pairs: the destination must be pair aligned
    if (pl.pow2 && !pl.force_generic && reinterpret_cast<uintptr_t>(d_frames_out) % (2 * pl.esize) == 0) {
        p.FT = pow2_c2r_frames_per_tile(pl.desc.n_fft);
        p.fd_FT = make_fastdiv(static_cast<unsigned>(p.FT));
        p.frame_stride = pl.pow2_frame_stride;
        p.tile_stride = pl.pow2_tile_stride;
        p.tiles_per_clip = static_cast<int>((n_frames + p.FT - 1) / p.FT);
        ck(launch_c2r_pow2(p, pl.f64, d_stft, d_frames_out, static_cast<long long>(nc), static_cast<long long>(n_frames),
                           apply_window, st), "kernel launch (c2r_pow2)");
        pl.last_launches += 1;
        return;
    }
    p.tiles_per_clip = static_cast<int>((n_frames + p.FT - 1) / p.FT);
    ck(launch_c2r_frames(p, pl.f64, pl.smem_bytes, d_stft, d_frames_out, static_cast<long long>(nc), static_cast<long long>(n_frames),
                         apply_window, st), "kernel launch (c2r_frames)");
    pl.last_launches += 1;
}
}  // namespace

sgx_status sgx_plan_istft(sgx_plan *plan, const void *stft, size_t n_clips, size_t n_frames, void *out, size_t *out_len_io,
                          void *cuda_stream) {
    return guarded([&] {
        if (!plan || !out_len_io) invalid("null argument");
        if (n_frames == 0) invalid("stft_matrix must have at least one frame");
        sgx_plan &pl = *plan;
        const size_t n = pl.desc.n_fft, hop = pl.desc.hop_size;
        const size_t pad = pl.desc.centre ? n / 2 : 0;
        const size_t full = (n_frames - 1) * hop + n;                               // :4837
        const size_t unpadded = full > 2 * pad ? full - 2 * pad : 0;                // saturating_sub (:4841)
        const bool trim = pl.desc.centre && unpadded > 0;                           // :4893
        const size_t out_len = trim ? unpadded : full;
        if (!out) { *out_len_io = out_len; return; }
        if (!stft) invalid("null argument");
        if (n_clips == 0) invalid("n_clips must be non-zero");
        if (*out_len_io != out_len) mismatch(out_len, *out_len_io);
        ensure_device(pl);
        DeviceGuard g(pl.device);
        pl.last_launches = 0;
        const PtrKind ki = ptr_kind(stft), ko = ptr_kind(out);
        if (ki != ko) invalid("stft and out must both be host pointers or both be device pointers");
        const bool host = ki != PtrKind::Device;
        const size_t es = pl.esize, bins = pl.tab.out_len;
        const size_t frame_bytes = n_frames * n * es, stft_bytes = bins * n_frames * 2 * es;
        size_t chunk = std::max<size_t>(1, (size_t(512) << 20) / std::max(frame_bytes, stft_bytes));
        chunk = std::min(chunk, n_clips);
        if (pl.frames_cap < chunk * frame_bytes) {
            if (pl.d_frames) { ck(cudaDeviceSynchronize(), "sync"); cudaFree(pl.d_frames); pl.d_frames = nullptr; pl.frames_cap = 0; }
            ck(cudaMalloc(&pl.d_frames, chunk * frame_bytes), "cudaMalloc(istft scratch)");
            pl.frames_cap = chunk * frame_bytes;
        }
        sgx_plan::Slot &s = pl.slot[0];
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        if (host) {
            ensure_slot(s, chunk * stft_bytes, chunk * out_len * es);
            st = s.s;
        }
        pl.scratch_acquire(st);
        for (size_t c0 = 0; c0 < n_clips; c0 += chunk) {
            const size_t nc = std::min(chunk, n_clips - c0);
            const void *src = static_cast<const char *>(stft) + c0 * stft_bytes;
            if (host) {
                ck(cudaMemcpyAsync(s.d_in, src, nc * stft_bytes, cudaMemcpyHostToDevice, st), "H2D copy");
                src = s.d_in;
            }
            void *dst = host ? s.d_out : static_cast<char *>(out) + c0 * out_len * es;
            // power-of-two plans: inverse FFT + window + overlap-add + normalisation in one kernel on a halo tile, as long as
            // the frames that overlap one hop slot fit a tile with slots to spare
            const int halo = static_cast<int>((n + hop - 1) / hop) - 1;
            const int ft = pl.pow2 ? pow2_c2r_frames_per_tile(n) : 0;
            static const bool unfused_env = std::getenv("SGX_ISTFT_UNFUSED") && std::atoi(std::getenv("SGX_ISTFT_UNFUSED")) != 0;
            if (pl.pow2 && !pl.force_generic && !unfused_env && hop <= n && 2 * halo <= ft) {
                KParams q;
                fill_params(pl, q);
                q.n_clips = static_cast<int>(nc);
                const size_t slots = (full + hop - 1) / hop;
                q.tiles_per_clip = static_cast<int>((slots + static_cast<size_t>(ft - halo) - 1) / static_cast<size_t>(ft - halo));
                q.fd_out_len = make_fastdiv(static_cast<unsigned>(hop));        // the kernel's divisor for position -> (slot, offset)
                ck(launch_istft_pow2(q, pl.f64, src, dst, static_cast<long long>(nc), static_cast<long long>(n_frames), halo,
                                     static_cast<long long>(out_len), static_cast<long long>(trim ? pad : 0), st), "kernel launch (istft_pow2)");
                pl.last_launches += 1;
                if (host) ck(cudaMemcpyAsync(static_cast<char *>(out) + c0 * out_len * es, s.d_out, nc * out_len * es, cudaMemcpyDeviceToHost, st), "D2H copy");
                continue;
            }
            run_inverse(pl, src, nc, n_frames, pl.d_frames, 1, st);
            ck(launch_ola_gather(pl.f64, pl.d_frames, pl.d_window, dst, static_cast<long long>(nc), static_cast<long long>(n_frames),
                                 static_cast<int>(n), static_cast<int>(hop), static_cast<long long>(out_len),
                                 static_cast<long long>(trim ? pad : 0), st), "kernel launch (ola_gather)");
            pl.last_launches += 1;
            if (host) ck(cudaMemcpyAsync(static_cast<char *>(out) + c0 * out_len * es, s.d_out, nc * out_len * es, cudaMemcpyDeviceToHost, st), "D2H copy");
        }
        pl.scratch_release(st);
        if (host) ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    });
}

// one C2R transform on an existing complex-STFT plan (rectangular window, hop = n_fft, no centring); host pointers are staged
static void irfft_with_plan(sgx_plan *pl, const void *spectrum, size_t spectrum_len, size_t n_fft, void *out, void *cuda_stream,
                            bool sync_device_path) {
    ensure_device(*pl);
    DeviceGuard g(pl->device);
    const size_t es = pl->esize;
    const PtrKind ki = ptr_kind(spectrum), ko = ptr_kind(out);
    if (ki != ko) invalid("spectrum and out must both be host pointers or both be device pointers");
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (ki == PtrKind::Device) {
        run_inverse(*pl, spectrum, 1, 1, out, 0, s);
        if (sync_device_path) ck(cudaStreamSynchronize(s), "cudaStreamSynchronize");
        return;
    }
    void *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, spectrum_len * 2 * es);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, n_fft * es);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, spectrum, spectrum_len * 2 * es, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        try { run_inverse(*pl, d_in, 1, 1, d_out, 0, s); } catch (...) { cudaFree(d_in); cudaFree(d_out); throw; }
        e = cudaMemcpyAsync(out, d_out, n_fft * es, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_in); cudaFree(d_out);
    ck(e, "irfft");
}

static sgx_plan_desc one_frame_desc(sgx_dtype dtype, size_t n_fft, int device) {
    sgx_plan_desc d;
    std::memset(&d, 0, sizeof d);
    d.dtype = dtype; d.n_fft = n_fft; d.hop_size = n_fft; d.centre = 0; d.window = SGX_WIN_RECTANGULAR;
    d.sample_rate_hz = 1.0; d.mapping = SGX_MAP_LINEAR; d.amp = SGX_AMP_POWER; d.output = SGX_OUT_COMPLEX_STFT; d.device = device;
    return d;
}

sgx_status sgx_irfft(sgx_dtype dtype, const void *spectrum, size_t spectrum_len, size_t n_fft, void *out, int device,
                     void *cuda_stream) {
    sgx_plan *pl = nullptr;
    sgx_status st = guarded([&] {
        if (!spectrum || !out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        if (spectrum_len != n_fft / 2 + 1) mismatch(n_fft / 2 + 1, spectrum_len);   // :4797-4802
    });
    if (st != SGX_OK) return st;
    const sgx_plan_desc d = one_frame_desc(dtype, n_fft, device);
    st = sgx_plan_create(&d, &pl);
    if (st != SGX_OK) return st;
    st = guarded([&] { irfft_with_plan(pl, spectrum, spectrum_len, n_fft, out, cuda_stream, true); });   // the plan's tables are freed below
    sgx_plan_destroy(pl);
    return st;
}

sgx_status sgx_rfft(sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft, void *out, int device, void *cuda_stream) {
    sgx_plan *pl = nullptr;
    sgx_status st = guarded([&] {
        if (!samples || !out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        if (n_in == 0) invalid("samples must be non-empty");
        if (n_in > n_fft) {                                                        // :4494-4500
            char buf[128];
            std::snprintf(buf, sizeof buf, "Input length (%zu) exceeds FFT size (%zu)", n_in, n_fft);
            invalid(buf);
        }
    });
    if (st != SGX_OK) return st;
    sgx_plan_desc d;
    std::memset(&d, 0, sizeof d);
    d.dtype = dtype; d.n_fft = n_fft; d.hop_size = n_fft; d.centre = 0; d.window = SGX_WIN_RECTANGULAR;
    d.sample_rate_hz = 1.0; d.mapping = SGX_MAP_LINEAR; d.amp = SGX_AMP_POWER; d.output = SGX_OUT_COMPLEX_STFT; d.device = device;
    st = sgx_plan_create(&d, &pl);
    if (st != SGX_OK) return st;
    st = sgx_plan_compute_frame(pl, samples, n_in, 0, out, cuda_stream);
    if (st == SGX_OK && ptr_kind(out) == PtrKind::Device) cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream));
    sgx_plan_destroy(pl);
    return st;
}

// ------------------------------------------------------------------------------------------------ FftPlanner
// FftPlanner (src/spectrogram.rs:4977-5235): the reference object owns an inner FFT planner that caches plans by size, so
// repeated single-frame transforms do not rebuild twiddles. Here the cache holds sgx_plan objects (device tables included)
// keyed by everything that shapes a plan.
struct sgx_fft_planner {
    int device = -1;
    std::map<std::tuple<int, size_t, int, int, uint64_t, int>, sgx_plan *> plans;   // dtype, n_fft, output, window, param bits, amp
    ~sgx_fft_planner() { for (auto &kv : plans) sgx_plan_destroy(kv.second); }
};

// cached plan for the key, created on first use; a creation failure leaves sgx_plan_create's status and message in place
static sgx_status planner_get(sgx_fft_planner *pp, sgx_dtype dtype, size_t n_fft, int output, int window, double wparam, int amp,
                              sgx_plan **out) {
    uint64_t bits;
    std::memcpy(&bits, &wparam, 8);
    const auto key = std::make_tuple(static_cast<int>(dtype), n_fft, output, window, bits, amp);
    auto it = pp->plans.find(key);
    if (it != pp->plans.end()) { *out = it->second; return SGX_OK; }
    sgx_plan_desc d = one_frame_desc(dtype, n_fft, pp->device);
    d.output = static_cast<sgx_output>(output);
    d.window = static_cast<sgx_window>(window);
    d.window_param = wparam;
    d.amp = static_cast<sgx_amp>(amp);
    sgx_plan *pl = nullptr;
    const sgx_status st = sgx_plan_create(&d, &pl);
    if (st != SGX_OK) return st;
    pp->plans.emplace(key, pl);
    *out = pl;
    return SGX_OK;
}

sgx_status sgx_fft_planner_create(int device, sgx_fft_planner **out) {
    return guarded([&] {
        if (!out) invalid("null argument");
        *out = new sgx_fft_planner();
        (*out)->device = device;
    });
}
sgx_status sgx_fft_planner_destroy(sgx_fft_planner *planner) {
    return guarded([&] { delete planner; });
}
size_t sgx_fft_planner_cached_plans(const sgx_fft_planner *planner) { return planner ? planner->plans.size() : 0; }

sgx_status sgx_fft_planner_rfft(sgx_fft_planner *planner, sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft, void *out,
                                void *cuda_stream) {
    sgx_plan *pl = nullptr;
    sgx_status st = guarded([&] {
        if (!planner || !samples || !out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        if (n_in == 0) invalid("samples must be non-empty");
        if (n_in > n_fft) {                                                        // :5034-5040
            char buf[128];
            std::snprintf(buf, sizeof buf, "Input length (%zu) exceeds FFT size (%zu)", n_in, n_fft);
            invalid(buf);
        }
    });
    if (st == SGX_OK) st = planner_get(planner, dtype, n_fft, SGX_OUT_COMPLEX_STFT, SGX_WIN_RECTANGULAR, 0.0, SGX_AMP_POWER, &pl);
    if (st != SGX_OK) return st;
    return sgx_plan_compute_frame(pl, samples, n_in, 0, out, cuda_stream);
}

sgx_status sgx_fft_planner_irfft(sgx_fft_planner *planner, sgx_dtype dtype, const void *spectrum, size_t spectrum_len, size_t n_fft,
                                 void *out, void *cuda_stream) {
    sgx_plan *pl = nullptr;
    sgx_status st = guarded([&] {
        if (!planner || !spectrum || !out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        if (spectrum_len != n_fft / 2 + 1) mismatch(n_fft / 2 + 1, spectrum_len);   // :5120-5126
    });
    if (st == SGX_OK) st = planner_get(planner, dtype, n_fft, SGX_OUT_COMPLEX_STFT, SGX_WIN_RECTANGULAR, 0.0, SGX_AMP_POWER, &pl);
    if (st != SGX_OK) return st;
    return guarded([&] { irfft_with_plan(pl, spectrum, spectrum_len, n_fft, out, cuda_stream, false); });
}

sgx_status sgx_fft_planner_power_spectrum(sgx_fft_planner *planner, sgx_dtype dtype, const void *samples, size_t n_in, size_t n_fft,
                                          sgx_window window, double window_param, int magnitude, void *out, void *cuda_stream) {
    sgx_plan *pl = nullptr;
    sgx_status st = guarded([&] {
        if (!planner || !samples || !out) invalid("null argument");
        if (n_fft == 0) invalid("n_fft must be set");
        if (n_in == 0) invalid("samples must be non-empty");
        if (n_in > n_fft) {                                                        // :5169-5175
            char buf[128];
            std::snprintf(buf, sizeof buf, "Input length (%zu) exceeds FFT size (%zu)", n_in, n_fft);
            invalid(buf);
        }
        if (window == SGX_WIN_CUSTOM) invalid("custom windows are not cached by the planner: build a linear plan instead");
    });
    if (st == SGX_OK)
        st = planner_get(planner, dtype, n_fft, SGX_OUT_SPECTROGRAM, window, window_param, magnitude ? SGX_AMP_MAGNITUDE : SGX_AMP_POWER, &pl);
    if (st != SGX_OK) return st;
    return sgx_plan_compute_frame(pl, samples, n_in, 0, out, cuda_stream);
}

}  // extern "C"
