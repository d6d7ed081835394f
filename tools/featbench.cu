// Standalone timing harness for the feature kernels (no Python).
//   * ABI cases: any number of libseld_cuda.so builds, dlopen-ed side by side (--lib name=path), timed through the C ABI
//     — what the library really runs.  Older builds lack seld_features_ex: their extra cases are skipped.
//   * in-binary cases: the fast kernel compiled HERE (features_fast.cuh) with stages removed from the end of its
//     pipeline (template parameter STRIP: the measured ceiling study of DESIGN.md) or with experiment switches (EXP).
// All cases run round-robin for several rounds in ONE process on ONE GPU; min and median per case are reported, so
// that clock / thermal drift between processes and boxes cannot pass for a kernel difference.
// usage: featbench [--B 256] [--seconds 60] [--n_fft 1024] [--iters 5] [--rounds 5] [--lib name=path]... [--strip] [--exp]
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

// The kernels compiled HERE live in their own namespace: a kernel instantiated both in this binary and in a dlopen-ed
// library under the same mangled name was seen to run ONE of the copies for all of them (equal times for builds that
// differ), so nothing in this binary may share a kernel name with the libraries under test.
#define seld seld_fb
#include "../include/seld_cuda.h"
#include "../sound-event-localization-detection_b200/csrc/features_fast.cuh"

namespace seld {  // the two internal helpers the kernel header's launch wrappers use (hidden symbols of the library)
void set_error(const std::string&) {}
int cuda_fail(cudaError_t e, const char* what) { printf("%s: %s\n", what, cudaGetErrorString(e)); return SELD_ERR_CUDA; }
}  // namespace seld

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
__global__ void to_i16(const float* x, short* y, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = (short)fminf(fmaxf(rintf(x[i] * 32768.0f), -32768.f), 32767.f);
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

struct Lib {
    std::string name;
    void* h = nullptr;
    seld_plan* plan = nullptr;
    decltype(&seld_plan_create) plan_create = nullptr;
    decltype(&seld_features) features = nullptr;
    decltype(&seld_features_ex) features_ex = nullptr;
    decltype(&seld_last_error) last_error = nullptr;
};

struct Case {
    std::string name;
    double bytes;
    std::function<void()> run;
    std::vector<float> ms;
};

int main(int argc, char** argv) {
    int B = 256, seconds = 60, n_fft = 1024, iters = 5, rounds = 5;
    bool strip = false, exp = false;
    std::vector<std::pair<std::string, std::string>> libs;
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char* s) { return !strcmp(argv[i], s) && i + 1 < argc; };
        if (is("--B")) B = atoi(argv[++i]);
        else if (is("--seconds")) seconds = atoi(argv[++i]);
        else if (is("--n_fft")) n_fft = atoi(argv[++i]);
        else if (is("--iters")) iters = atoi(argv[++i]);
        else if (is("--rounds")) rounds = atoi(argv[++i]);
        else if (is("--lib")) { std::string a = argv[++i]; auto p = a.find('='); libs.push_back({a.substr(0, p), a.substr(p + 1)}); }
        else if (!strcmp(argv[i], "--strip")) strip = true;
        else if (!strcmp(argv[i], "--exp")) exp = true;
        else { printf("unknown argument %s\n", argv[i]); return 2; }
    }
    const int hop = 480, C = 4;
    const long long N = 24000ll * seconds;
    const long long T = 1 + N / hop;
    std::vector<float> win(n_fft);
    for (int i = 0; i < n_fft; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / n_fft));
    std::vector<float> fb = n_fft == 1024 ? baked_fb<1024>() : baked_fb<960>();

    float *audio, *out;
    CK(cudaMalloc(&audio, sizeof(float) * B * C * N));
    const size_t out_n = (size_t)B * T * 10 * 64;
    CK(cudaMalloc(&out, sizeof(float) * out_n));
    CK(cudaMemset(out, 0xff, sizeof(float) * out_n));
    fill_noise<<<1184, 256>>>(audio, (long long)B * C * N, 1234u);
    short* pcm;
    CK(cudaMalloc(&pcm, sizeof(short) * B * C * N));
    to_i16<<<1184, 256>>>(audio, pcm, (long long)B * C * N);
    double* stats;
    CK(cudaMalloc(&stats, sizeof(double) * 2 * 10 * 64));
    CK(cudaMemset(stats, 0, sizeof(double) * 2 * 10 * 64));
    float *mean, *istd;
    CK(cudaMalloc(&mean, 4 * 448)); CK(cudaMalloc(&istd, 4 * 448));
    CK(cudaMemset(mean, 0, 4 * 448));
    std::vector<float> ones(448, 1.f);
    CK(cudaMemcpy(istd, ones.data(), 4 * 448, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());

    auto bytes_of = [&](int c_out, int in_bytes, int out_bytes) {
        return (double)B * seconds * ((double)in_bytes * 24000 * 4 + (double)c_out * 64 * 50 * out_bytes);
    };
    std::vector<Case> cases;
    std::vector<Lib> L(libs.size());
    for (size_t i = 0; i < libs.size(); ++i) {
        Lib& l = L[i];
        l.name = libs[i].first;
        l.h = dlopen(libs[i].second.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!l.h) { printf("dlopen %s: %s\n", libs[i].second.c_str(), dlerror()); return 1; }
        l.plan_create = (decltype(l.plan_create))dlsym(l.h, "seld_plan_create");
        l.features = (decltype(l.features))dlsym(l.h, "seld_features");
        l.features_ex = (decltype(l.features_ex))dlsym(l.h, "seld_features_ex");
        l.last_error = (decltype(l.last_error))dlsym(l.h, "seld_last_error");
        if (l.plan_create(&l.plan, 0, n_fft, hop, 64, win.data(), fb.data()) != 0) { printf("plan (%s): %s\n", l.name.c_str(), l.last_error()); return 1; }
        Lib* lp = &L[i];
        auto add = [&](const char* nm, int mode, int c_out, double* st) {
            cases.push_back({l.name + ": " + nm, bytes_of(c_out, 4, 4), [=] {
                const int rc = lp->features(lp->plan, mode, audio, C * N, N, N, nullptr, B, C, out, T, c_out, 0, st, nullptr, nullptr, nullptr);
                if (rc != 0) { printf("features (%s, mode %d): rc %d: %s | %s\n", lp->name.c_str(), mode, rc, lp->last_error(), cudaGetErrorString(cudaGetLastError())); exit(1); }
            }, {}});
        };
        add("foa (7 ch)", 1, 7, nullptr);
        add("logmel (4 ch)", 0, 4, nullptr);
        add("foa + scaler partials", 1, 7, stats);
        if (n_fft == 1024 || l.features_ex) add("mic (4 log-mel + 6 GCC-PHAT)", 2, 10, nullptr);
        if (l.features_ex) {
            cases.push_back({l.name + ": foa, int16 PCM input", bytes_of(7, 2, 4), [=] {
                seld_feat_opts o{};
                o.in_dtype = SELD_DTYPE_I16;
                if (lp->features_ex(lp->plan, 1, pcm, C * N, N, N, nullptr, B, C, out, T, 7, 0, nullptr, nullptr, nullptr, &o, nullptr) != 0) { printf("features_ex: %s\n", lp->last_error()); exit(1); }
            }, {}});
            cases.push_back({l.name + ": foa, normalised (B,C,T,F) bf16 out", bytes_of(7, 4, 2), [=] {
                seld_feat_opts o{};
                o.out_layout = SELD_LAYOUT_CTF;
                o.out_dtype = SELD_DTYPE_BF16;
                o.d_mean = mean;
                o.d_inv_std = istd;
                if (lp->features_ex(lp->plan, 1, audio, C * N, N, N, nullptr, B, C, out, T, 7, 0, nullptr, nullptr, nullptr, &o, nullptr) != 0) { printf("features_ex: %s\n", lp->last_error()); exit(1); }
            }, {}});
        }
    }
    // in-binary variants of the fast kernel (n_fft 1024, 7 channels) need a plan struct: built by the in-tree library
    seld_plan* plan = nullptr;
    seld::FeatArgs a{};
    if (strip || exp) {
        if (n_fft != 1024) { printf("--strip / --exp: n_fft 1024 only\n"); return 2; }
        if (seld_plan_create(&plan, 0, n_fft, hop, 64, win.data(), fb.data()) != 0) { printf("plan: %s\n", seld_last_error()); return 1; }
        a.audio = audio; a.clip_stride = C * N; a.chan_stride = N; a.n_samples = N; a.B = B; a.C = 4; a.G = 1;
        a.out = out; a.T_out = T; a.C_out = 7; a.c_off = 0; a.n_out = 7; a.n_items = (long long)B * T;
        a.status = plan->d_status;
    }
    unsigned* redo = nullptr;
    if (strip || exp) {
        CK(cudaMalloc(&redo, sizeof(unsigned) * (4 + seld::kRedoCap)));
        CK(cudaMemset(redo, 0, sizeof(unsigned) * (4 + seld::kRedoCap)));
        a.redo = redo;
    }
#define INBIN(S, BFK, NAME)                                                                                   \
    {                                                                                                         \
        if (seld::fast_configure<32, true, false, 0, 12, BFK, S>() != 0) return 1;                             \
        cases.push_back({NAME, bytes_of(7, 4, 4), [=] { seld::fast_launch<32, true, false, 0, 12, BFK, S>(plan, a, nullptr); }, {}}); \
    }
    if (strip) {
        INBIN(5, false, "strip5: loads + window only")
        INBIN(4, false, "strip4: + 2 packed FFTs")
        INBIN(3, false, "strip3: + channel split (shuffles)")
        INBIN(2, false, "strip2: + per-bin features -> planes, barriers")
        INBIN(1, false, "strip1: + mel phase")
        INBIN(0, false, "strip0: + log + row copy-out (= lean kernel)")
    }
    if (exp) {
        INBIN(0, false, "lean kernel alone (in-binary build)")
        INBIN(0, true, "block-floating kernel over every frame")
    }

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // warm the clocks with ~1.5 s of real work, then every case once
    for (int w = 0; w < 60 && !cases.empty(); ++w) cases[0].run();
    for (auto& c : cases) c.run();
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < rounds; ++r)
        for (auto& c : cases) {
            c.run();
            cudaEventRecord(e0);
            for (int it = 0; it < iters; ++it) c.run();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float t; cudaEventElapsedTime(&t, e0, e1);
            c.ms.push_back(t / iters);
        }
    printf("# B=%d x %d s, n_fft=%d, iters=%d, rounds=%d (min / median ms; %% of 6551.7 GB/s at the min)\n", B, seconds, n_fft, iters, rounds);
    for (auto& c : cases) {
        std::sort(c.ms.begin(), c.ms.end());
        const float mn = c.ms.front(), md = c.ms[c.ms.size() / 2];
        printf("%-52s %7.3f / %7.3f ms  %7.3f M clip-s/s  %6.1f GB/s  %5.1f%%\n", c.name.c_str(), mn, md, B * seconds / mn / 1e3,
               c.bytes / mn / 1e6, 100.0 * c.bytes / mn / 1e6 / 6551.7);
    }
    return 0;
}
