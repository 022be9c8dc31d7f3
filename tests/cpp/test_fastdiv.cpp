// test_fastdiv.cpp -- the multiply-shift divisors the generic kernel family uses for its index math (fastdiv.hpp,
// make_fastdiv / fd_div) against exact integer division, on the host. fd_div's device form is umulhi(n, mul) >> shr;
// the same arithmetic is restated here with a 64-bit product.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../spectrograms_b200/csrc/fastdiv.hpp"

static int host_fd_div(int n, const sgx::FastDiv &fd) {
    return fd.mul == 0 ? n : static_cast<int>((static_cast<uint64_t>(static_cast<uint32_t>(n)) * fd.mul >> 32) >> fd.shr);
}

int main() {
    std::mt19937_64 rng(7);
    long checked = 0;
    auto check = [&](unsigned d, int n) {
        const sgx::FastDiv fd = sgx::make_fastdiv(d);
        if (host_fd_div(n, fd) != n / static_cast<int>(d)) { std::fprintf(stderr, "FAILED d=%u n=%d\n", d, n); std::exit(1); }
        ++checked;
    };
    for (unsigned d = 1; d <= 5000; ++d) {
        for (int n : {0, 1, static_cast<int>(d) - 1, static_cast<int>(d), static_cast<int>(d) + 1, 2 * static_cast<int>(d) - 1, 0x7fffffff, 0x40000007}) if (n >= 0) check(d, n);
        for (int k = 0; k < 50; ++k) check(d, static_cast<int>(rng() & 0x7fffffff));
    }
    for (int k = 0; k < 200000; ++k) check(static_cast<unsigned>(rng() % 0x1000000) + 1, static_cast<int>(rng() & 0x7fffffff));
    std::printf("FASTDIV_OK %ld\n", checked);
    return 0;
}
