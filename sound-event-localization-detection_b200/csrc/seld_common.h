// Shared host/device declarations of libseld_cuda (internal; the public C ABI is include/seld_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/seld_cuda.h"

namespace seld {

constexpr int kMaxMels = 64;
constexpr int kFeatWarps = 12;  // warps per CTA of the feature kernel (1 CTA / SM, smem-limited)

// Device-side view of a plan (passed by value to kernels).
struct PlanDev {
    int n_fft, r1, hop, n_bins, n_mels;
    int la, lb;                // trip counts of the two mel gather loops (multiples of 4)
    const float* window;       // [n_fft], pre-scaled by 1/2 (exact) so the channel split needs no factor
    const float2* twiddle;     // [r1][32]: W_N^{lane*k_lo}
    const int2* mel_entries;   // [la+lb][32]: {.x = byte offset of the bin row (bin*16), .y = float bits of weight}
    const int* mel_idx;        // [2][32]: mel index of slot A / slot B per lane, -1 = none
};

struct FeatArgs {
    const float* audio;
    long long clip_stride, chan_stride, n_samples;
    const long long* lengths;   // device, may be null
    int B, C, G;                // G = channel groups of 4
    float* out;
    long long T_out;
    int C_out, c_off, n_out;     // n_out = channels this call writes
    double* stats;              // may be null
    const int* stat_frames;     // may be null
    float2* spec;               // may be null
    long long n_items;          // B * G * T_out
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

}  // namespace seld

struct seld_plan {
    int device;
    int num_sms;
    seld::PlanDev dev;
    void* d_blob;       // one allocation holding all tables
    size_t table_bytes; // constant tables at the start of the feature kernel's dynamic smem
    size_t warp_smem;   // + this many bytes per warp (Q rows + max(R rows, transpose tile))
    bool v3_ok;         // the filterbank equals the baked one of mel_baked.h: the v3 kernel may be used
};

#define SELD_CUDA_TRY(expr)                                            \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return seld::cuda_fail(_e, #expr);      \
    } while (0)
