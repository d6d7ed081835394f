// extern "C" surface of libseld_cuda (see include/seld_cuda.h) + plan construction.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "seld_common.h"
#include "mel_tables.h"

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

int launch_features(const seld_plan* plan, bool iv, const FeatArgs& a, cudaStream_t stream);
bool v3_filterbank_matches(int n_fft, const float* fb, int n_mels);
int launch_gcc(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream);
int launch_feature_stats(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream);
int launch_labels_fill(float* out, long long rows, int cells, int M, cudaStream_t st);
int launch_labels_paint(float* out, long long rows, int I, int J, int M, const int* events, const double* centres,
                        int n_events, double sigma_az, double sigma_el, cudaStream_t st);
int launch_window_gather(const float* src, long long rows, long long row_len, const long long* starts, int n_win,
                         int win_len, const float* pad_row, float* out, cudaStream_t st);
int launch_scaler_apply(float* x, long long rows, int n_feat, const float* mean, const float* inv_std, cudaStream_t st);
int launch_pcm16_to_float(const short* in, float* out, long long n, cudaStream_t st);

}  // namespace seld

using namespace seld;

extern "C" {

int seld_version(void) { return 100; }
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
    SELD_CUDA_TRY(cudaSetDevice(device));
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
    const size_t total = off_idx + sizeof(int) * 64;
    std::vector<unsigned char> blob(total);
    std::memcpy(blob.data() + off_win, win.data(), sizeof(float) * n_fft);
    std::memcpy(blob.data() + off_tw, tw.data(), sizeof(float2) * tw.size());
    std::memcpy(blob.data() + off_mel, mt.entries.data(), sizeof(int2) * mt.entries.size());
    std::memcpy(blob.data() + off_idx, mt.idx.data(), sizeof(int) * 64);

    seld_plan* p = new (std::nothrow) seld_plan();
    if (!p) {
        set_error("seld_plan_create: out of host memory");
        return SELD_ERR_ALLOC;
    }
    p->device = device;
    cudaError_t e = cudaMalloc(&p->d_blob, total);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, blob.data(), total, cudaMemcpyHostToDevice);
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
    p->table_bytes = total;
    p->warp_smem = (size_t)(n_bins + std::max(n_bins, 528)) * sizeof(float4);
    p->v3_ok = v3_filterbank_matches(n_fft, h_fb, n_mels);
    *out = p;
    return SELD_OK;
}

int seld_plan_destroy(seld_plan* plan) {
    if (!plan) return SELD_OK;
    cudaSetDevice(plan->device);
    if (plan->d_blob) cudaFree(plan->d_blob);
    delete plan;
    return SELD_OK;
}

int seld_features(seld_plan* plan, int mode, const float* d_audio, int64_t clip_stride, int64_t chan_stride,
                  int64_t n_samples, const int64_t* d_lengths, int B, int C, float* d_out, int64_t T_out, int C_out,
                  int c_off, double* d_stats, const int32_t* d_stat_frames, float* d_spec, void* stream) {
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
    if (!d_lengths && n_samples <= plan->dev.n_fft / 2)
        return bad_arg("seld_features: reflect padding needs more than n_fft/2 samples");
    if (B == 0 || T_out == 0) return SELD_OK;
    FeatArgs a;
    a.audio = d_audio;
    a.clip_stride = clip_stride;
    a.chan_stride = chan_stride;
    a.n_samples = n_samples;
    a.lengths = reinterpret_cast<const long long*>(d_lengths);
    a.B = B;
    a.C = C;
    a.G = (C + 3) / 4;
    a.out = d_out;
    a.T_out = T_out;
    a.C_out = C_out;
    a.c_off = c_off;
    a.n_out = n_out;
    a.stats = d_stats;
    a.stat_frames = d_stat_frames;
    a.spec = reinterpret_cast<float2*>(d_spec);
    const long long n_items = (long long)B * a.G * T_out;
    if (n_items >= (1ll << 31)) return bad_arg("seld_features: B * ceil(C/4) * T_out must be < 2^31 per call");
    a.n_items = n_items;
    SELD_CUDA_TRY(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == SELD_MODE_LOGMEL_GCC) {  // channels [c_off, c_off+4): log-mel, [c_off+4, c_off+10): GCC-PHAT
        FeatArgs lm = a;
        lm.n_out = 4;
        lm.stats = nullptr;
        lm.spec = a.spec;
        int rc = launch_features(plan, false, lm, st);
        if (rc != SELD_OK) return rc;
        FeatArgs g = a;
        g.c_off = c_off + 4;
        rc = launch_gcc(plan, g, st);
        if (rc != SELD_OK) return rc;
        return a.stats ? launch_feature_stats(plan, a, st) : SELD_OK;
    }
    return launch_features(plan, mode == SELD_MODE_LOGMEL_IV, a, st);
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
    SELD_CUDA_TRY(cudaSetDevice(plan->device));
    return launch_feature_stats(plan, a, static_cast<cudaStream_t>(stream));
}

int seld_scaler_apply(float* d_x, int64_t rows, int n_feat, const float* d_mean, const float* d_inv_std,
                      void* stream) {
    if (!d_x || !d_mean || !d_inv_std) return bad_arg("seld_scaler_apply: null pointer");
    if (rows < 0 || n_feat < 1) return bad_arg("seld_scaler_apply: bad size");
    return launch_scaler_apply(d_x, rows, n_feat, d_mean, d_inv_std, static_cast<cudaStream_t>(stream));
}

int seld_pcm16_to_float(const int16_t* d_pcm, float* d_out, int64_t n, void* stream) {
    if (n < 0) return bad_arg("seld_pcm16_to_float: negative size");
    if (n == 0) return SELD_OK;
    if (!d_pcm || !d_out) return bad_arg("seld_pcm16_to_float: null pointer");
    return launch_pcm16_to_float(d_pcm, d_out, n, static_cast<cudaStream_t>(stream));
}

int seld_labels_fill(float* d_out, int64_t rows, int cells, int n_classes, void* stream) {
    if (!d_out) return bad_arg("seld_labels_fill: null pointer");
    if (rows < 0 || cells < 1 || n_classes < 1) return bad_arg("seld_labels_fill: bad size");
    return launch_labels_fill(d_out, rows, cells, n_classes, static_cast<cudaStream_t>(stream));
}

int seld_labels_paint(float* d_out, int64_t rows, int I, int J, int n_classes, const int32_t* d_events,
                      const double* d_centres, int n_events, double sigma_az, double sigma_el, void* stream) {
    if (n_events < 0) return bad_arg("seld_labels_paint: negative event count");
    if (n_events == 0) return SELD_OK;
    if (!d_out || !d_events) return bad_arg("seld_labels_paint: null pointer");
    if (I < 1 || J < 1 || n_classes < 1 || rows < 0) return bad_arg("seld_labels_paint: bad size");
    return launch_labels_paint(d_out, rows, I, J, n_classes, d_events, d_centres, n_events, sigma_az, sigma_el,
                               static_cast<cudaStream_t>(stream));
}

int seld_window_gather(const float* d_src, int64_t rows, int64_t row_len, const int64_t* d_starts, int n_win,
                       int win_len, const float* d_pad_row, float* d_out, void* stream) {
    if (n_win < 0 || win_len < 0 || row_len < 0 || rows < 0) return bad_arg("seld_window_gather: negative size");
    if (n_win == 0 || win_len == 0 || row_len == 0) return SELD_OK;
    if (!d_src || !d_starts || !d_pad_row || !d_out) return bad_arg("seld_window_gather: null pointer");
    return launch_window_gather(d_src, rows, row_len, reinterpret_cast<const long long*>(d_starts), n_win, win_len,
                                d_pad_row, d_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
