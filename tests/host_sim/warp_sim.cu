// CPU emulation of the feature kernel's per-warp algorithm, lane by lane, using the SAME __host__ __device__
// arithmetic (dft_inreg.cuh / warp_fft.cuh) and the same mel tables (mel_tables.h) as the CUDA kernel.
// The build container has no GPU; this is how the index logic (transpose swizzle, mirror-bin exchange,
// Q/R layout, mel schedule) is validated before spending GPU time.  Test infrastructure only.
//
// usage: warp_sim <n_fft> <C> <N> <iv> in.bin out.bin
//   in.bin : float32 window[n_fft], fb[n_bins*64], audio[C*N]
//   out.bin: float32 (T, C_out, 64) features, then float32 complex (C, T, n_bins) spectra
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../sound-event-localization-detection_b200/csrc/warp_fft.cuh"
#include "../../sound-event-localization-detection_b200/csrc/mel_tables.h"
using namespace seld;

template <int R1>
static void fft_pair_sim(float2 (*u)[32], const float* xa, const float* xb, long long start, long long len,
                         const float* win, const float2* tw, float2* T) {
    using F = WarpFft<R1>;
    for (int lane = 0; lane < 32; ++lane) {
        float2 v[R1];
        for (int j = 0; j < R1; ++j) {
            long long idx = F::reflect(start + lane + 32 * j, len);
            float w = win[lane + 32 * j];
            v[j] = make_float2((xa ? xa[idx] : 0.f) * w, (xb ? xb[idx] : 0.f) * w);
        }
        F::pass1(v, tw + lane);
        F::t_store(v, T, lane);
    }
    for (int lane = 0; lane < 32; ++lane) {
        F::t_load(u[lane], T, lane);
        F::pass2(u[lane]);
    }
}

template <int R1>
static int run(int C, long long N, bool iv, const float* window, const float* fb, const float* audio, FILE* fo) {
    using F = WarpFft<R1>;
    constexpr int NFFT = F::N, NB = F::NB;
    const int hop = 480, n_mels = 64;
    std::vector<float> win(NFFT);
    for (int i = 0; i < NFFT; ++i) win[i] = 0.5f * window[i];
    std::vector<float2> tw(R1 * 32);
    for (int k = 0; k < R1; ++k)
        for (int l = 0; l < 32; ++l) {
            double a = -2.0 * M_PI * double((long long)k * l % NFFT) / double(NFFT);
            tw[k * 32 + l] = make_float2((float)cos(a), (float)sin(a));
        }
    MelTables mt = build_mel_tables(fb, NB, n_mels);
    fprintf(stderr, "mel tables: la=%d lb=%d\n", mt.la, mt.lb);
    // bank-conflict estimate of the schedule: wavefronts per quarter-warp LDS.128
    {
        double wf = 0; long cnt = 0;
        for (int it = 0; it < mt.la + mt.lb; ++it)
            for (int q = 0; q < 4; ++q) {
                int c[8] = {0};
                int mx = 0;
                for (int l = 0; l < 8; ++l) { int r = (mt.entries[it * 32 + q * 8 + l].x >> 4) & 7; mx = std::max(mx, ++c[r]); }
                wf += mx; ++cnt;
            }
        fprintf(stderr, "mel schedule: %.3f wavefronts per quarter-warp load (1.0 = conflict-free)\n", wf / cnt);
    }
    const long long T = 1 + N / hop;
    const int G = (C + 3) / 4;
    const int C_out = iv ? 7 : C;
    std::vector<float> out((size_t)T * C_out * n_mels, 0.f);
    std::vector<float2> spec((size_t)C * T * NB);
    std::vector<float4> Q(NB), R(NB > 528 ? NB : 528);  // R also hosts the 32x33 transpose tile
    float2* Tt = reinterpret_cast<float2*>(R.data());
    static float2 u[32][32];
    for (long long t = 0; t < T; ++t)
        for (int g = 0; g < G; ++g) {
            const int c0 = 4 * g, nch = std::min(4, C - c0);
            const float* x = audio + (size_t)c0 * N;
            const long long start = t * hop - F::HALF;
            fft_pair_sim<R1>(u, x, nch > 1 ? x + N : nullptr, start, N, win.data(), tw.data(), Tt);
            for (int lane = 0; lane < R1; ++lane)
                for (int kh = 0; kh <= 16; ++kh) {
                    if (kh == 16 && lane != 0) break;
                    float2 z = u[lane][kh], pz;
                    if (kh < 16) {
                        const int src = F::partner_lane(lane);
                        pz = u[src][31 - kh];
                        if (lane == 0) pz = u[0][(32 - kh) & 31];
                    } else pz = z;
                    float2 x0, x1;
                    F::unpack(z, pz, x0, x1);
                    const int k = F::bin_of(lane, kh);
                    Q[k] = make_float4(x0.x, x0.y, x1.x, x1.y);
                    spec[((size_t)(c0) * T + t) * NB + k] = x0;
                    if (nch > 1) spec[((size_t)(c0 + 1) * T + t) * NB + k] = x1;
                }
            const bool have_b = nch > 2;
            if (have_b)
                fft_pair_sim<R1>(u, x + 2 * N, nch > 3 ? x + 3 * N : nullptr, start, N, win.data(), tw.data(), Tt);
            std::vector<float4> Rn(NB > 528 ? NB : 528);
            for (int lane = 0; lane < R1; ++lane)
                for (int kh = 0; kh <= 16; ++kh) {
                    if (kh == 16 && lane != 0) break;
                    float2 x2 = make_float2(0, 0), x3 = x2;
                    if (have_b) {
                        float2 z = u[lane][kh], pz;
                        if (kh < 16) {
                            const int src = F::partner_lane(lane);
                            pz = u[src][31 - kh];
                            if (lane == 0) pz = u[0][(32 - kh) & 31];
                        } else pz = z;
                        F::unpack(z, pz, x2, x3);
                    }
                    const int k = F::bin_of(lane, kh);
                    float4 s = Q[k], q, r;
                    if (iv) bin_features<true>(make_float2(s.x, s.y), make_float2(s.z, s.w), x2, x3, q, r);
                    else bin_features<false>(make_float2(s.x, s.y), make_float2(s.z, s.w), x2, x3, q, r);
                    Q[k] = q; Rn[k] = r;  // (device writes R in place: T has been consumed by then)
                    if (have_b) {
                        spec[((size_t)(c0 + 2) * T + t) * NB + k] = x2;
                        if (nch > 3) spec[((size_t)(c0 + 3) * T + t) * NB + k] = x3;
                    }
                }
            R = Rn; Tt = reinterpret_cast<float2*>(R.data());
            for (int lane = 0; lane < 32; ++lane) {
                float acc[2][7] = {{0}};
                for (int i = 0; i < mt.la + mt.lb; ++i) {
                    const int s = i < mt.la ? 0 : 1;
                    const int2 en = mt.entries[(size_t)i * 32 + lane];
                    float w; memcpy(&w, &en.y, 4);
                    const float4 q = Q[en.x >> 4], r = R[en.x >> 4];
                    acc[s][0] = fmaf(w, q.x, acc[s][0]); acc[s][1] = fmaf(w, q.y, acc[s][1]);
                    acc[s][2] = fmaf(w, r.x, acc[s][2]); acc[s][3] = fmaf(w, r.y, acc[s][3]);
                    acc[s][4] = fmaf(w, q.z, acc[s][4]); acc[s][5] = fmaf(w, q.w, acc[s][5]);
                    acc[s][6] = fmaf(w, r.z, acc[s][6]);
                }
                for (int s = 0; s < 2; ++s) {
                    const int m = mt.idx[s * 32 + lane];
                    if (m < 0) continue;
                    for (int c = 0; c < (iv ? 7 : 4); ++c) {
                        if (c < 4 && c >= nch) continue;
                        const float v = c < 4 ? power_to_db(acc[s][c]) : acc[s][c];
                        out[((size_t)t * C_out + c0 + c) * n_mels + m] = v;
                    }
                }
            }
        }
    fwrite(out.data(), sizeof(float), out.size(), fo);
    fwrite(spec.data(), sizeof(float2), spec.size(), fo);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage\n"); return 2; }
    const int n_fft = atoi(argv[1]), C = atoi(argv[2]);
    const long long N = atoll(argv[3]);
    const bool iv = atoi(argv[4]) != 0;
    const int NB = n_fft / 2 + 1;
    FILE* fi = fopen(argv[5], "rb");
    FILE* fo = fopen(argv[6], "wb");
    if (!fi || !fo) return 3;
    std::vector<float> window(n_fft), fb((size_t)NB * 64), audio((size_t)C * N);
    if (fread(window.data(), 4, window.size(), fi) != window.size()) return 4;
    if (fread(fb.data(), 4, fb.size(), fi) != fb.size()) return 4;
    if (fread(audio.data(), 4, audio.size(), fi) != audio.size()) return 4;
    int rc = n_fft == 1024 ? run<32>(C, N, iv, window.data(), fb.data(), audio.data(), fo)
                           : run<30>(C, N, iv, window.data(), fb.data(), audio.data(), fo);
    fclose(fi); fclose(fo);
    return rc;
}
