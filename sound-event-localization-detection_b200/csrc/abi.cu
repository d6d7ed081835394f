// extern "C" surface of libseld_cuda (see include/seld_cuda.h) + plan construction.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "seld_common.h"
#include "mel_tables.h"

namespace seld { constexpr int kRedoCapAbi = 1 << 16; }  // == kRedoCap of features_fast.cuh (static_assert in features.cu)

namespace seld {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
    return SELD_ERR_CUDA;
}
static int bad_arg(const std::string& msg) {
    set_error(msg);
    return SELD_ERR_BAD_ARG;
}
static int unsupported(const std::string& msg) {
    set_error(msg);
    return SELD_ERR_UNSUPPORTED;
}
// device that owns a device pointer (the plan-less entry points run where their output lives, whatever the
// caller's current device is); -1 if the pointer is not device memory
static int device_of(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}
#define SELD_GUARD_PTR(ptr, what)                                                          \
    const int _dev = device_of(ptr);                                                       \
    if (_dev < 0) return bad_arg(std::string(what) + ": not a device pointer");            \
    DeviceGuard guard(_dev);                                                               \
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, what)

int launch_features(const seld_plan* plan, bool iv, const FeatArgs& a, cudaStream_t stream);
bool fast_filterbank_matches(int n_fft, const float* fb, int n_mels);
bool fast_path_ok(const seld_plan* plan, const FeatArgs& a);
int configure_feature_kernels(const seld_plan* plan);
int configure_gcc_kernels(const seld_plan* plan);
int launch_gcc(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream);
int launch_feature_stats(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream);
int launch_labels_fill(float* out, long long rows, int cells, int M, cudaStream_t st);
int launch_labels_paint(float* out, long long rows, int I, int J, int M, const int* events, const double* centres,
                        int n_events, double sigma_az, double sigma_el, cudaStream_t st);
int launch_window_gather(const float* src, long long rows, long long row_len, const long long* starts, int n_win,
                         int win_len, const float* pad_row, float* out, cudaStream_t st);
int launch_scaler_apply(float* x, long long rows, int n_feat, const float* mean, const float* inv_std, cudaStream_t st);
int launch_pcm16_to_float(const short* in, float* out, long long n, cudaStream_t st);
int launch_batch_class_mask(const int* order, int first, int n_win, const int* win_start, const int* win_lo, const int* win_hi,
                            int win_len, const int* events, const double* centres, int I, int J, int M, double sigma_az,
                            double sigma_el, unsigned short* mask, cudaStream_t st);
int launch_aux_losses(const float* logits, const unsigned short* mask, long long n_frames, int I, int J, int M, double* sums,
                      float* grad, const float* gscale, cudaStream_t st);
int launch_class_loss(int mode, const float* logits, const unsigned short* mask, long long n_cells, int M, const float* weight,
                      double* sums, float* grad, const float* gscale, cudaStream_t st);
int launch_loader_batch(const float* feat, long long rows, int row_len, const int* order, int first, int n_win,
                        const int* win_start, const int* win_lo, const int* win_hi, int win_len, float* out_spec,
                        const int* events, const double* centres, int I, int J, int M, double sigma_az, double sigma_el,
                        float* out_lab, cudaStream_t st);

}  // namespace seld

using namespace seld;

extern "C" {

int seld_version(void) { return 201; }
const char* seld_last_error(void) { return g_last_error.c_str(); }
int64_t seld_num_frames(int64_t n_samples, int hop) { return hop > 0 ? 1 + n_samples / hop : 0; }
int seld_out_channels(int mode, int n_channels) {
    switch (mode) {
        case SELD_MODE_LOGMEL: return n_channels;
        case SELD_MODE_LOGMEL_IV: return 7;
        case SELD_MODE_LOGMEL_GCC: return 10;
        default: return -1;
    }
}

int seld_plan_create(seld_plan** out, int device, int n_fft, int hop, int n_mels, const float* h_window,
                     const float* h_fb) {
    if (!out || !h_window || !h_fb) return bad_arg("seld_plan_create: null argument");
    *out = nullptr;
    if (n_fft != 1024 && n_fft != 960) return unsupported("seld_plan_create: n_fft must be 960 or 1024");
    if (hop <= 0) return bad_arg("seld_plan_create: hop must be positive");
    if (n_mels < 1 || n_mels > kMaxMels) return unsupported("seld_plan_create: n_mels must be in [1, 64]");
    DeviceGuard guard(device);  // the caller's current device is restored on return
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "seld_plan_create: cudaSetDevice");
    const int r1 = n_fft / 32, n_bins = n_fft / 2 + 1;

    std::vector<float> win(n_fft);
    for (int i = 0; i < n_fft; ++i) win[i] = 0.5f * h_window[i];
    std::vector<float2> tw((size_t)r1 * 32);
    for (int k = 0; k < r1; ++k)
        for (int l = 0; l < 32; ++l) {
            const double a = -2.0 * M_PI * double((long long)k * l % n_fft) / double(n_fft);
            tw[(size_t)k * 32 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    MelTables mt = build_mel_tables(h_fb, n_bins, n_mels);

    const size_t off_win = 0;
    const size_t off_tw = off_win + sizeof(float) * n_fft;
    const size_t off_mel = off_tw + sizeof(float2) * tw.size();
    const size_t off_idx = off_mel + sizeof(int2) * mt.entries.size();
    const size_t total = off_idx + sizeof(int) * 64;      // what the generic kernel copies into shared memory
    const size_t off_status = (total + 15) & ~size_t(15);  // + the device status word
    std::vector<unsigned char> blob(off_status + 16, 0);
    std::memcpy(blob.data() + off_win, win.data(), sizeof(float) * n_fft);
    std::memcpy(blob.data() + off_tw, tw.data(), sizeof(float2) * tw.size());
    std::memcpy(blob.data() + off_mel, mt.entries.data(), sizeof(int2) * mt.entries.size());
    std::memcpy(blob.data() + off_idx, mt.idx.data(), sizeof(int) * 64);

    // generic kernel: as many warps per CTA as the 227 KB of shared memory hold next to this plan's tables (the mel
    // gather table grows when n_mels shrinks: wider filters)
    const size_t warp_smem = (size_t)(n_bins + std::max(n_bins, 528)) * sizeof(float4);
    int warps = total < (size_t)kMaxSmemOptin ? (int)(((size_t)kMaxSmemOptin - total) / warp_smem) : 0;
    if (warps > kFeatWarps) warps = kFeatWarps;
    if (warps < 1) return unsupported("seld_plan_create: the tables of this filterbank do not fit in shared memory");

    seld_plan* p = new (std::nothrow) seld_plan();
    if (!p) {
        set_error("seld_plan_create: out of host memory");
        return SELD_ERR_ALLOC;
    }
    p->device = device;
    p->d_blob = nullptr;
    cudaError_t e = cudaMalloc(&p->d_blob, blob.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        if (p->d_blob) cudaFree(p->d_blob);
        delete p;
        return cuda_fail(e, "seld_plan_create");
    }
    unsigned char* d = static_cast<unsigned char*>(p->d_blob);
    p->dev.n_fft = n_fft;
    p->dev.r1 = r1;
    p->dev.hop = hop;
    p->dev.n_bins = n_bins;
    p->dev.n_mels = n_mels;
    p->dev.la = mt.la;
    p->dev.lb = mt.lb;
    p->dev.window = reinterpret_cast<const float*>(d + off_win);
    p->dev.twiddle = reinterpret_cast<const float2*>(d + off_tw);
    p->dev.mel_entries = reinterpret_cast<const int2*>(d + off_mel);
    p->dev.mel_idx = reinterpret_cast<const int*>(d + off_idx);
    p->d_status = reinterpret_cast<int*>(d + off_status);
    p->table_bytes = total;
    p->warp_smem = warp_smem;
    p->generic_warps = warps;
    p->v3_ok = fast_filterbank_matches(n_fft, h_fb, n_mels);
    // A/B switches, read ONCE here (never on the hot calls): SELD_FEAT_IMPL=v2 forces the generic kernel,
    // SELD_V3_CFG=8 selects the 8-warp resource configuration of the fast kernel
    const char* impl = getenv("SELD_FEAT_IMPL");
    p->force_generic = impl && impl[0] == 'v' && impl[1] == '2';
    p->force_bf = impl && impl[0] == 'b' && impl[1] == 'f';
    p->n_slots = 0;
    p->slot_mutex = new std::mutex();
    p->d_redo = nullptr;
    if (p->v3_ok) {
        const size_t redo_bytes = sizeof(unsigned) * (size_t)kRedoSlots * (4 + kRedoCapAbi);
        e = cudaMalloc(&p->d_redo, redo_bytes);
        if (e == cudaSuccess) e = cudaMemset(p->d_redo, 0, redo_bytes);
        if (e != cudaSuccess) {
            if (p->d_redo) cudaFree(p->d_redo);
            cudaFree(p->d_blob);
            delete static_cast<std::mutex*>(p->slot_mutex);
            delete p;
            return cuda_fail(e, "seld_plan_create: redo lists");
        }
    }
    const char* cfg = getenv("SELD_V3_CFG");
    p->fast_warps = (cfg && cfg[0] == '8') ? 8 : 12;
    int rc = configure_feature_kernels(p);
    if (rc == SELD_OK) rc = configure_gcc_kernels(p);
    if (rc != SELD_OK) {
        cudaFree(p->d_blob);
        if (p->d_redo) cudaFree(p->d_redo);
        delete static_cast<std::mutex*>(p->slot_mutex);
        delete p;
        return rc;
    }
    *out = p;
    return SELD_OK;
}

int seld_plan_destroy(seld_plan* plan) {
    if (!plan) return SELD_OK;
    DeviceGuard guard(plan->device);
    if (plan->d_blob) cudaFree(plan->d_blob);
    if (plan->d_redo) cudaFree(plan->d_redo);
    delete static_cast<std::mutex*>(plan->slot_mutex);
    delete plan;
    return SELD_OK;
}

int seld_plan_has_fast_path(const seld_plan* plan) { return plan && plan->v3_ok && !plan->force_generic ? 1 : 0; }

int seld_plan_status(seld_plan* plan, void* stream, int* h_status) {
    if (!plan || !h_status) return bad_arg("seld_plan_status: null argument");
    DeviceGuard guard(plan->device);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "seld_plan_status: cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int v = 0;
    SELD_CUDA_TRY(cudaMemcpyAsync(&v, plan->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SELD_CUDA_TRY(cudaMemsetAsync(plan->d_status, 0, sizeof(int), st));
    SELD_CUDA_TRY(cudaStreamSynchronize(st));
    *h_status = v;
    if (v & 1) {
        set_error("seld_features: a clip in d_lengths has <= n_fft/2 samples (reflect padding is undefined; torch.stft "
                  "raises there): its rows were written as 0");
        return SELD_ERR_BAD_ARG;
    }
    return SELD_OK;
}

int seld_features_ex(seld_plan* plan, int mode, const void* d_audio, int64_t clip_stride, int64_t chan_stride,
                     int64_t n_samples, const int64_t* d_lengths, int B, int C, void* d_out, int64_t T_out, int C_out,
                     int c_off, double* d_stats, const int32_t* d_stat_frames, float* d_spec, const seld_feat_opts* opts,
                     void* stream) {
    if (!plan) return bad_arg("seld_features: null plan");
    if (!d_audio || !d_out) return bad_arg("seld_features: null device pointer");
    if (B < 0 || C < 1 || T_out < 0) return bad_arg("seld_features: negative size");
    const int n_out = seld_out_channels(mode, C);
    if (n_out < 0) return bad_arg("seld_features: unknown mode");
    if ((mode == SELD_MODE_LOGMEL_IV || mode == SELD_MODE_LOGMEL_GCC) && C != 4)
        return bad_arg("seld_features: IV / GCC-PHAT modes need exactly 4 channels");
    if (mode == SELD_MODE_LOGMEL_GCC && plan->dev.n_mels != 64)
        return unsupported("seld_features: GCC-PHAT mode needs n_mels == 64 lags");
    if (c_off < 0 || c_off + n_out > C_out) return bad_arg("seld_features: output channels out of range");
    if (n_samples < 0 || n_samples > 0x7fffffffll) return bad_arg("seld_features: n_samples must be in [0, 2^31)");
    if (!d_lengths && n_samples <= plan->dev.n_fft / 2)
        return bad_arg("seld_features: reflect padding needs more than n_fft/2 samples");
    if (opts) {
        if (opts->in_dtype != SELD_DTYPE_F32 && opts->in_dtype != SELD_DTYPE_I16)
            return bad_arg("seld_features_ex: in_dtype must be SELD_DTYPE_F32 or SELD_DTYPE_I16");
        if (opts->out_dtype != SELD_DTYPE_F32 && opts->out_dtype != SELD_DTYPE_BF16)
            return bad_arg("seld_features_ex: out_dtype must be SELD_DTYPE_F32 or SELD_DTYPE_BF16");
        if (opts->out_layout != SELD_LAYOUT_TCF && opts->out_layout != SELD_LAYOUT_CTF)
            return bad_arg("seld_features_ex: out_layout must be SELD_LAYOUT_TCF or SELD_LAYOUT_CTF");
        if ((opts->d_mean == nullptr) != (opts->d_inv_std == nullptr))
            return bad_arg("seld_features_ex: d_mean and d_inv_std go together");
    }
    if (B == 0 || T_out == 0) return SELD_OK;
    FeatArgs a{};
    a.audio = d_audio;
    a.in_i16 = opts && opts->in_dtype == SELD_DTYPE_I16;
    a.clip_stride = clip_stride;
    a.chan_stride = chan_stride;
    a.n_samples = n_samples;
    a.lengths = reinterpret_cast<const long long*>(d_lengths);
    a.B = B;
    a.C = C;
    a.G = (C + 3) / 4;
    a.out = static_cast<float*>(d_out);
    a.T_out = T_out;
    a.C_out = C_out;
    a.c_off = c_off;
    a.n_out = n_out;
    a.stats = d_stats;
    a.stat_frames = d_stat_frames;
    a.spec = reinterpret_cast<float2*>(d_spec);
    a.mean = opts ? opts->d_mean : nullptr;
    a.inv_std = opts ? opts->d_inv_std : nullptr;
    a.out_ctf = opts && opts->out_layout == SELD_LAYOUT_CTF;
    a.out_bf16 = opts && opts->out_dtype == SELD_DTYPE_BF16;
    a.status = plan->d_status;
    a.redo = nullptr;
    a.redo_mode = 0;
    a.sink = nullptr;
    const long long n_items = (long long)B * a.G * T_out;
    if (n_items >= (1ll << 31)) return bad_arg("seld_features: B * ceil(C/4) * T_out must be < 2^31 per call");
    a.n_items = n_items;
    DeviceGuard guard(plan->device);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "seld_features: cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == SELD_MODE_LOGMEL_GCC) {
        if (a.in_i16 || a.mean || a.out_ctf || a.out_bf16)
            return unsupported("seld_features_ex: the GCC-PHAT mode takes float32 input and writes plain float32 rows");
        const int rc = launch_gcc(plan, a, st);  // ONE launch: channels [c_off, c_off+4) log-mel, [c_off+4, c_off+10) GCC-PHAT
        if (rc != SELD_OK) return rc;
        return a.stats ? launch_feature_stats(plan, a, st) : SELD_OK;
    }
    return launch_features(plan, mode == SELD_MODE_LOGMEL_IV, a, st);
}

int seld_features(seld_plan* plan, int mode, const float* d_audio, int64_t clip_stride, int64_t chan_stride,
                  int64_t n_samples, const int64_t* d_lengths, int B, int C, float* d_out, int64_t T_out, int C_out,
                  int c_off, double* d_stats, const int32_t* d_stat_frames, float* d_spec, void* stream) {
    return seld_features_ex(plan, mode, d_audio, clip_stride, chan_stride, n_samples, d_lengths, B, C, d_out, T_out, C_out,
                            c_off, d_stats, d_stat_frames, d_spec, nullptr, stream);
}

int seld_feature_stats(seld_plan* plan, const float* d_feat, int B, int64_t T_out, int C_out, int c_off, int n_channels,
                       int64_t n_samples, const int64_t* d_lengths, const int32_t* d_stat_frames, double* d_stats,
                       void* stream) {
    if (!plan) return bad_arg("seld_feature_stats: null plan");
    if (!d_feat || !d_stats) return bad_arg("seld_feature_stats: null device pointer");
    if (B < 0 || T_out < 0 || C_out < 1 || c_off < 0 || n_channels < 1 || c_off + n_channels > C_out)
        return bad_arg("seld_feature_stats: bad size");
    if (B == 0 || T_out == 0) return SELD_OK;
    FeatArgs a{};
    a.out = const_cast<float*>(d_feat);
    a.B = B;
    a.T_out = T_out;
    a.C_out = C_out;
    a.c_off = c_off;
    a.n_out = n_channels;
    a.n_samples = n_samples;
    a.lengths = reinterpret_cast<const long long*>(d_lengths);
    a.stat_frames = d_stat_frames;
    a.stats = d_stats;
    DeviceGuard guard(plan->device);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "seld_feature_stats: cudaSetDevice");
    return launch_feature_stats(plan, a, static_cast<cudaStream_t>(stream));
}

int seld_scaler_apply(float* d_x, int64_t rows, int n_feat, const float* d_mean, const float* d_inv_std,
                      void* stream) {
    if (!d_x || !d_mean || !d_inv_std) return bad_arg("seld_scaler_apply: null pointer");
    if (rows < 0 || n_feat < 1) return bad_arg("seld_scaler_apply: bad size");
    SELD_GUARD_PTR(d_x, "seld_scaler_apply");
    return launch_scaler_apply(d_x, rows, n_feat, d_mean, d_inv_std, static_cast<cudaStream_t>(stream));
}

int seld_pcm16_to_float(const int16_t* d_pcm, float* d_out, int64_t n, void* stream) {
    if (n < 0) return bad_arg("seld_pcm16_to_float: negative size");
    if (n == 0) return SELD_OK;
    if (!d_pcm || !d_out) return bad_arg("seld_pcm16_to_float: null pointer");
    SELD_GUARD_PTR(d_out, "seld_pcm16_to_float");
    return launch_pcm16_to_float(d_pcm, d_out, n, static_cast<cudaStream_t>(stream));
}

int seld_labels_fill(float* d_out, int64_t rows, int cells, int n_classes, void* stream) {
    if (!d_out) return bad_arg("seld_labels_fill: null pointer");
    if (rows < 0 || cells < 1 || n_classes < 1) return bad_arg("seld_labels_fill: bad size");
    if (rows == 0) return SELD_OK;
    SELD_GUARD_PTR(d_out, "seld_labels_fill");
    return launch_labels_fill(d_out, rows, cells, n_classes, static_cast<cudaStream_t>(stream));
}

int seld_labels_paint(float* d_out, int64_t rows, int I, int J, int n_classes, const int32_t* d_events,
                      const double* d_centres, int n_events, double sigma_az, double sigma_el, void* stream) {
    if (n_events < 0) return bad_arg("seld_labels_paint: negative event count");
    if (n_events == 0) return SELD_OK;
    if (!d_out || !d_events) return bad_arg("seld_labels_paint: null pointer");
    if (I < 1 || J < 1 || n_classes < 1 || rows < 0) return bad_arg("seld_labels_paint: bad size");
    SELD_GUARD_PTR(d_out, "seld_labels_paint");
    return launch_labels_paint(d_out, rows, I, J, n_classes, d_events, d_centres, n_events, sigma_az, sigma_el,
                               static_cast<cudaStream_t>(stream));
}

int seld_window_gather(const float* d_src, int64_t rows, int64_t row_len, const int64_t* d_starts, int n_win,
                       int win_len, const float* d_pad_row, float* d_out, void* stream) {
    if (n_win < 0 || win_len < 0 || row_len < 0 || rows < 0) return bad_arg("seld_window_gather: negative size");
    if (n_win == 0 || win_len == 0 || row_len == 0) return SELD_OK;
    if (!d_src || !d_starts || !d_pad_row || !d_out) return bad_arg("seld_window_gather: null pointer");
    SELD_GUARD_PTR(d_out, "seld_window_gather");
    return launch_window_gather(d_src, rows, row_len, reinterpret_cast<const long long*>(d_starts), n_win, win_len,
                                d_pad_row, d_out, static_cast<cudaStream_t>(stream));
}

int seld_loader_batch(const float* d_feat, int64_t rows, int row_len, const int32_t* d_order, int first, int n_win,
                      const int32_t* d_win_start, const int32_t* d_win_lo, const int32_t* d_win_hi, int win_len,
                      float* d_spec_out, const int32_t* d_events, const double* d_centres, int I, int J, int n_classes,
                      double sigma_az, double sigma_el, float* d_labels_out, void* stream) {
    if (n_win < 0 || win_len < 0 || rows < 0 || row_len < 1 || first < 0) return bad_arg("seld_loader_batch: bad size");
    if (n_win == 0 || win_len == 0) return SELD_OK;
    if (!d_feat || !d_win_start || !d_spec_out) return bad_arg("seld_loader_batch: null pointer");
    if (d_labels_out && (!d_win_lo || !d_win_hi || I < 1 || J < 1 || n_classes < 1))
        return bad_arg("seld_loader_batch: labels need the per-window event ranges and the grid size");
    SELD_GUARD_PTR(d_spec_out, "seld_loader_batch");
    return launch_loader_batch(d_feat, rows, row_len, d_order, first, n_win, d_win_start, d_win_lo, d_win_hi, win_len,
                               d_spec_out, d_events, d_centres, I, J, n_classes, sigma_az, sigma_el, d_labels_out,
                               static_cast<cudaStream_t>(stream));
}

int seld_batch_class_mask(const int32_t* d_order, int first, int n_win, const int32_t* d_win_start, const int32_t* d_win_lo,
                          const int32_t* d_win_hi, int win_len, const int32_t* d_events, const double* d_centres, int I, int J,
                          int n_classes, double sigma_az, double sigma_el, uint16_t* d_mask, void* stream) {
    if (n_win < 0 || win_len < 0 || first < 0 || I < 1 || J < 1 || n_classes < 1) return bad_arg("seld_batch_class_mask: bad size");
    if (n_win == 0 || win_len == 0) return SELD_OK;
    if (!d_win_start || !d_win_lo || !d_win_hi || !d_mask) return bad_arg("seld_batch_class_mask: null pointer");
    SELD_GUARD_PTR(d_mask, "seld_batch_class_mask");
    return launch_batch_class_mask(d_order, first, n_win, d_win_start, d_win_lo, d_win_hi, win_len, d_events, d_centres, I, J,
                                   n_classes, sigma_az, sigma_el, d_mask, static_cast<cudaStream_t>(stream));
}

int seld_class_loss(int loss_type, const float* d_logits, const uint16_t* d_mask, int64_t n_cells, int n_classes,
                    const float* d_class_weight, double* d_sums, float* d_grad, const float* d_grad_scale, void* stream) {
    if (loss_type != SELD_LOSS_MSE && loss_type != SELD_LOSS_CE) return bad_arg("seld_class_loss: unknown loss type");
    if (n_cells < 0 || n_classes < 1) return bad_arg("seld_class_loss: bad size");
    if (n_cells == 0) return SELD_OK;
    if (!d_logits || !d_mask || (!d_sums && !d_grad)) return bad_arg("seld_class_loss: null pointer");
    if (d_grad && !d_grad_scale) return bad_arg("seld_class_loss: d_grad needs d_grad_scale");
    SELD_GUARD_PTR(d_logits, "seld_class_loss");
    return launch_class_loss(loss_type, d_logits, d_mask, n_cells, n_classes, d_class_weight, d_sums, d_grad, d_grad_scale,
                             static_cast<cudaStream_t>(stream));
}

int seld_aux_losses(const float* d_logits, const uint16_t* d_mask, int64_t n_frames, int I, int J, int n_classes,
                    double* d_sums, float* d_grad, const float* d_grad_scale, void* stream) {
    if (n_frames < 0 || I < 1 || J < 1 || n_classes < 1) return bad_arg("seld_aux_losses: bad size");
    if (n_frames == 0) return SELD_OK;
    if (!d_logits || !d_mask || (!d_sums && !d_grad)) return bad_arg("seld_aux_losses: null pointer");
    if (d_grad && !d_grad_scale) return bad_arg("seld_aux_losses: d_grad needs d_grad_scale");
    SELD_GUARD_PTR(d_logits, "seld_aux_losses");
    return launch_aux_losses(d_logits, d_mask, n_frames, I, J, n_classes, d_sums, d_grad, d_grad_scale,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
