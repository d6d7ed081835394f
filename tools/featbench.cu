// Standalone A/B harness for the feature kernels through the C ABI (no Python): times the generic (v2) and
// the v3 kernel on synthetic audio resident in HBM and reports their largest difference.
//   featbench [B=64] [seconds=60] [n_fft=1024] [mode=1] [iters=5] [impl=both|v2|v3]
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../include/seld_cuda.h"
#include "../sound-event-localization-detection_b200/csrc/mel_baked.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_noise(float* x, long long n, unsigned seed) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        float u1 = ((z & 0xffffff) + 1) * (1.0f / 16777217.0f), u2 = ((z >> 24) & 0xffffff) * (1.0f / 16777216.0f);
        x[i] = 0.1f * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
}

template <int NFFT>
static std::vector<float> baked_fb() {
    using MB = seld::MelBaked<NFFT>;
    std::vector<float> fb((size_t)MB::NB * 64, 0.f);
    for (int k = 0; k < MB::NB; ++k) {
        if (MB::m0[k] >= 0) fb[(size_t)k * 64 + MB::m0[k]] = MB::w0[k];
        if (MB::m1[k] >= 0) fb[(size_t)k * 64 + MB::m1[k]] = MB::w1[k];
    }
    return fb;
}

int main(int argc, char** argv) {
    int B = argc > 1 ? atoi(argv[1]) : 64;
    int seconds = argc > 2 ? atoi(argv[2]) : 60;
    int n_fft = argc > 3 ? atoi(argv[3]) : 1024;
    int mode = argc > 4 ? atoi(argv[4]) : 1;
    int iters = argc > 5 ? atoi(argv[5]) : 5;
    const char* impl = argc > 6 ? argv[6] : "both";
    const int hop = 480, C = 4;
    const long long N = 24000ll * seconds;
    const long long T = 1 + N / hop;
    const int C_out = seld_out_channels(mode, C);
    std::vector<float> win(n_fft);
    for (int i = 0; i < n_fft; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / n_fft));
    std::vector<float> fb = n_fft == 1024 ? baked_fb<1024>() : baked_fb<960>();
    seld_plan* plan = nullptr;
    if (seld_plan_create(&plan, 0, n_fft, hop, 64, win.data(), fb.data()) != 0) { printf("plan: %s\n", seld_last_error()); return 1; }
    float *audio, *out[2];
    CK(cudaMalloc(&audio, sizeof(float) * B * C * N));
    const size_t out_n = (size_t)B * T * C_out * 64;
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&out[i], sizeof(float) * out_n)); CK(cudaMemset(out[i], 0xff, sizeof(float) * out_n)); }
    fill_noise<<<1184, 256>>>(audio, (long long)B * C * N, 1234u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)B * seconds * (4.0 * 24000 * 4 + (double)C_out * 64 * 50 * 4);
    const char* names[2] = {"v2", "v3"};
    bool ran[2] = {false, false};
    for (int which = 0; which < 2; ++which) {
        if (strcmp(impl, "both") != 0 && strcmp(impl, names[which]) != 0) continue;
        setenv("SELD_FEAT_IMPL", names[which], 1);
        for (int w = 0; w < 2; ++w)
            if (seld_features(plan, mode, audio, C * N, N, N, nullptr, B, C, out[which], T, C_out, 0, nullptr, nullptr, nullptr, nullptr) != 0) { printf("features: %s\n", seld_last_error()); return 1; }
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int it = 0; it < iters; ++it)
            seld_features(plan, mode, audio, C * N, N, N, nullptr, B, C, out[which], T, C_out, 0, nullptr, nullptr, nullptr, nullptr);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
        printf("%s n_fft=%d mode=%d B=%d x %ds: %.3f ms/step  %.3f M clip-s/s  %.1f GB/s (%.1f%% of 6551.7)\n", names[which], n_fft, mode, B,
               seconds, ms, B * seconds / ms / 1e3, bytes / ms / 1e6, 100.0 * bytes / ms / 1e6 / 6551.7);
        ran[which] = true;
    }
    if (ran[0] && ran[1]) {
        std::vector<float> h0(out_n), h1(out_n);
        CK(cudaMemcpy(h0.data(), out[0], sizeof(float) * out_n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h1.data(), out[1], sizeof(float) * out_n, cudaMemcpyDeviceToHost));
        double worst_db = 0, worst_iv = 0; size_t bad = 0, where = 0;
        for (size_t i = 0; i < out_n; ++i) {
            const int c = (i / 64) % C_out;
            const double d = fabs((double)h0[i] - (double)h1[i]);
            if (!(d == d)) { if (!bad) where = i; ++bad; continue; }
            if (c < 4) { if (d > worst_db) { worst_db = d; where = i; } } else if (d > worst_iv) worst_iv = d;
        }
        printf("v2 vs v3: max |d| log-mel %.3e dB, IV/other %.3e, NaN/unwritten %zu (first/worst at %zu: %g vs %g)\n", worst_db, worst_iv, bad,
               where, h0[where], h1[where]);
    }
    seld_plan_destroy(plan);
    return 0;
}
