// fastdiv.hpp -- division by a launch-uniform run-time constant as a multiply-high and a shift (host + device).
#pragma once

namespace sgx {

// Division by a run-time constant without the ~20-instruction integer divide: q = umulhi(n, mul) >> shr for 0 <= n < 2^31
// (the round-up multiplier of Granlund & Montgomery). mul == 0 encodes "divide by 1". Built on the host by make_fastdiv.
struct FastDiv {
    unsigned mul, shr;
};
constexpr int kFdStages = 12;      // stages with a precomputed divisor (more stages than this fall back to '/')
#ifdef __CUDACC__
__device__ __forceinline__ int fd_div(int n, const FastDiv &fd) {
    return fd.mul == 0 ? n : static_cast<int>(__umulhi(static_cast<unsigned>(n), fd.mul) >> fd.shr);
}
#endif
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv fd{0u, 0u};
    if (d <= 1) return fd;
    unsigned lg = 0;                                   // ceil(log2 d)
    while ((1ull << lg) < d) ++lg;
    const unsigned p = 31 + lg;
    fd.mul = static_cast<unsigned>(((1ull << p) + d - 1) / d);
    fd.shr = p - 32;
    return fd;
}

}  // namespace sgx
