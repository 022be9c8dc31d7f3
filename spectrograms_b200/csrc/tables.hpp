// tables.hpp -- host-side (f64) construction of everything a plan precomputes: window, frequency mapping,
// frequency axis, DCT-II basis, lifter. Mirrors what the reference builds once per plan
// (StftPlan::new src/spectrogram.rs:1204-1228, FrequencyMapping::new_* :1676-1783). Compiled with
// -ffp-contract=off so that only the reference's explicit mul_add sites are fused.
#pragma once

#include <cstddef>
#include <string>
#include <vector>

#include "../../include/sgx_b200.h"

namespace sgx {

struct Error {
    sgx_status code;
    std::string msg;          // payload (without the Display prefix)
    size_t expected = 0, got = 0;
};

struct HostTables {
    size_t out_len = 0;       // n_fft/2 + 1
    size_t n_bins = 0;        // rows of the frequency mapping
    std::vector<double> window;                 // n_fft
    // sparse mapping (mel / loghz), CSR, ascending column order per row
    std::vector<int> row_ptr, col;
    std::vector<double> val;
    // dense mapping (erb) (n_bins, out_len)
    std::vector<double> dense;
    std::vector<double> freq_axis;              // n_bins
    // mfcc
    std::vector<double> dct;                    // (n_mfcc, n_mels) cos(pi k (i+0.5)/n_mels)
    std::vector<double> lifter;                 // n_mfcc (all ones when lifter == 0)
};

// frame_count (src/spectrogram.rs:1230-1250)
size_t frame_count(size_t n_samples, size_t n_fft, size_t hop, bool centre);

// validation + construction; throws sgx::Error
void validate_desc(const sgx_plan_desc &d);
void build_tables(const sgx_plan_desc &d, HostTables &t);

// chroma (src/chroma.rs:81-112, :279-346): validation (throws) and the dense [12][n_fft/2 + 1] filterbank
void validate_chroma(double sample_rate_hz, double tuning, double f_min, double f_max);
void build_chroma_filterbank(double sample_rate_hz, size_t n_fft, double tuning, double f_min, double f_max,
                             std::vector<double> &dense);

// DCT-II basis and lifter weights (src/mfcc.rs:278-316)
void build_dct(size_t n_mfcc, size_t n_mels, size_t lifter, std::vector<double> &basis, std::vector<double> &lift);

}  // namespace sgx
