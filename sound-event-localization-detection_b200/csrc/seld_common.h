// Shared host/device declarations of libseld_cuda (internal; the public C ABI is include/seld_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/seld_cuda.h"

namespace seld {

constexpr int kMaxMels = 64;
constexpr int kFeatWarps = 12;  // most warps per CTA of the feature kernels (1 CTA / SM, smem-limited)
constexpr int kMaxSmemOptin = 232448;  // 227 KB: the per-CTA opt-in limit of sm_100
constexpr int kRedoSlots = 16;         // streams per plan with a redo list of their own (features_fast.cuh)

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
    const void* audio;          // float32, or int16 PCM when in_i16 (fast path only)
    int in_i16;
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
    // epilogue options of the fast path (seld_feat_opts): normalisation tables, model-stem layout, bf16
    const float* mean;          // [C_out * n_mels] or null
    const float* inv_std;       // [C_out * n_mels] or null
    int out_ctf;                // 0: (B, T, C, F)   1: (B, C, T, F)
    int out_bf16;               // 0: float32        1: bfloat16
    int* status;                // plan-owned device status word (bit 0: a clip was too short for reflect padding)
    unsigned* redo;             // fast path: redo list of this stream's slot: [0] count, [1] CTAs done, [4..] frame indices
    int redo_mode;              // block-floating kernel: 1 = run over the redo list, 0 = over all frames
    float* sink;                // always null in the library (keeps the stripped variants of tools/featbench alive)
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

}  // namespace seld

struct seld_plan {
    int device;
    int num_sms;
    seld::PlanDev dev;
    void* d_blob;       // one allocation holding all tables (+ the status word)
    int* d_status;      // device status word, see seld_plan_status()
    size_t table_bytes; // constant tables at the start of the generic feature kernel's dynamic smem
    size_t warp_smem;   // + this many bytes per warp (Q rows + max(R rows, transpose tile))
    int generic_warps;  // warps per CTA of the generic kernel that fit the 227 KB of shared memory (<= 12)
    bool v3_ok;         // the filterbank equals the baked one of mel_baked.h: the fast kernel may be used
    bool force_generic; // SELD_FEAT_IMPL=v2 at plan creation (A/B measurements)
    bool force_bf;      // SELD_FEAT_IMPL=bf: the block-floating kernel over every frame (tests, A/B measurements)
    // redo lists of the fast path: one slot per CUDA stream that has used the plan (calls on one stream are ordered)
    unsigned* d_redo;   // kRedoSlots x (4 + kRedoCap) words
    void* slot_stream[seld::kRedoSlots];
    int n_slots;
    void* slot_mutex;   // std::mutex*
    int fast_warps;     // 12, or 8 when SELD_V3_CFG=8 at plan creation (second resource configuration)
};

namespace seld {
// RAII: make `device` current for the duration of an ABI call and restore the caller's device afterwards
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace seld

#define SELD_CUDA_TRY(expr)                                            \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return seld::cuda_fail(_e, #expr);      \
    } while (0)
