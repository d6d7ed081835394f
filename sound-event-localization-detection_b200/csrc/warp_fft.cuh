// One warp = one complex FFT of N = R1*32 points (R1 = 32 -> n_fft 1024, R1 = 30 -> n_fft 960), used to
// transform TWO real channels at once (z = x_a + i x_b).  Replaces torch.stft as driven by
// torchaudio.transforms.MelSpectrogram in the reference (dataset.py:38-50 ->
// torchaudio/functional/functional.py:123-134): frame t = reflect-padded x[t*hop - N/2 .. +N) * hann.
//
// Decomposition (n = lane + 32 j, k = k_lo + R1 k_hi):
//   pass 1 (registers): Y[lane][k_lo] = sum_j v[j] W_R1^{j k_lo}          (Dft<R1>, immediates)
//   twiddle           : Y *= W_N^{lane k_lo}                               (table tw[k_lo][lane], smem)
//   transpose (smem)  : lane <-> k_lo, 32x32 float2 tile with row pitch 33, conflict-free both ways
//   pass 2 (registers): X[k_lo + R1 k_hi] = sum_n u[n] W_32^{n k_hi}       (Dft<32>, immediates)
// After pass 2 lane k_lo holds bins k_lo + R1*k_hi in register k_hi.  The mirror bin N-k needed to split
// the two real channels lives in lane (R1 - k_lo) % R1, register 31 - k_hi (32 - k_hi for lane 0), i.e.
// one __shfl per component with a STATIC register index.
//
// All per-lane arithmetic is __host__ __device__; tests/host_sim runs the same functions lane by lane on
// the CPU (the build container has no GPU).
#pragma once
#include "dft_inreg.cuh"

namespace seld {

template <int R1>
struct WarpFft {
    static constexpr int N = R1 * 32;
    static constexpr int NB = N / 2 + 1;   // one-sided bins
    static constexpr int HALF = N / 2;

    // reflect index (torch pad_mode="reflect": edge sample not repeated); valid for len > N/2
    static SELD_HD long long reflect(long long idx, long long len) {
        if (idx < 0) idx = -idx;
        if (idx >= len) idx = 2 * (len - 1) - idx;
        return idx;
    }

    // pass 1 on the lane's R1 windowed samples + inter-pass twiddle.  tw points at tw[0][lane], row pitch 32.
    static SELD_HD void pass1(float2 (&v)[R1], const float2* tw_lane) {
        Dft<R1, false>::run(v);
        static_for<R1 - 1>([&](auto K) {
            constexpr int k = decltype(K)::value + 1;
            v[k] = cmul(v[k], tw_lane[k * 32]);
        });
    }

    // transpose through the tile T: element (row k_lo, col lane) at row*TP + col, row pitch TP = 33 float2.
    // Stores (fixed row, consecutive lanes) and loads (fixed column, row = lane: 66*lane words -> bank pair
    // 2*lane mod 32, distinct over a half-warp) are both conflict-free, and every address is lane base +
    // compile-time immediate.
    static constexpr int TP = 33;
    static constexpr int T_FLOAT2 = 32 * TP;
    static SELD_HD void t_store(const float2 (&v)[R1], float2* T, int lane) {
        float2* p = T + lane;
        static_for<R1>([&](auto K) {
            constexpr int k = decltype(K)::value;
            p[k * TP] = v[k];
        });
    }
    static SELD_HD void t_load(float2 (&u)[32], const float2* T, int lane) {
        const float2* p = T + (R1 < 32 && lane >= R1 ? 0 : lane) * TP;
        static_for<32>([&](auto C) {
            constexpr int c = decltype(C)::value;
            u[c] = p[c];
        });
        if (R1 < 32 && lane >= R1) static_for<32>([&](auto C) { u[decltype(C)::value] = make_float2(0.f, 0.f); });
    }
    static SELD_HD void pass2(float2 (&u)[32]) { Dft<32, false>::run(u); }

    // ---- inverse transform pieces (GCC-PHAT), unnormalised: x[t] = sum_k G[k] exp(+2 pi i k t / N) ----
    // same data flow as the forward transform with conjugated twiddles; lane = t_lo after the transpose.
    static SELD_HD void pass1_inv(float2 (&v)[R1], const float2* tw_lane) {
        Dft<R1, true>::run(v);
        static_for<R1 - 1>([&](auto K) {
            constexpr int k = decltype(K)::value + 1;
            v[k] = cmul_conj(v[k], tw_lane[k * 32]);
        });
    }
    // pruned second pass: only t_hi = 0 (lags t_lo) and t_hi = 31 (lags t_lo - 32) are needed
    static SELD_HD void pass2_inv_pruned(const float2 (&w)[32], float2& lag_pos, float2& lag_neg) {
        float2 a[32], b[32];
        static_for<32>([&](auto Nn) {
            constexpr int n = decltype(Nn)::value;
            a[n] = w[n];
            b[n] = mul_w<n, 32, false>(w[n]);  // exp(+2 pi i 31 n / 32) = W_32^n
        });
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1)
#pragma unroll
            for (int i = 0; i < s; ++i) {
                a[i] = cadd(a[i], a[i + s]);
                b[i] = cadd(b[i], b[i + s]);
            }
        lag_pos = a[0];
        lag_neg = b[0];
    }

    // lane that holds the mirror bin of this lane's bins
    static SELD_HD int partner_lane(int lane) { return (lane == 0 || lane >= R1) ? 0 : R1 - lane; }
    static SELD_HD int bin_of(int lane, int k_hi) { return lane + R1 * k_hi; }

    // Is either channel of the packed pair identically zero in this frame?  (OR of the raw sample bits, sign
    // ignored, then a warp vote.)  A silent channel must come out as an exactly-zero spectrum like the
    // reference's separate FFT — the split below would otherwise leave the rounding asymmetry of the other
    // channel (~1e-7 relative) in it.
#ifdef __CUDACC__
    static __device__ __forceinline__ void silent_channels(const float2 (&v)[R1], bool& a_silent, bool& b_silent) {
        unsigned ba = 0u, bb = 0u;
        static_for<R1>([&](auto J) {
            constexpr int j = decltype(J)::value;
            ba |= __float_as_uint(v[j].x);
            bb |= __float_as_uint(v[j].y);
        });
        a_silent = !__any_sync(0xffffffffu, (ba << 1) != 0u);
        b_silent = !__any_sync(0xffffffffu, (bb << 1) != 0u);
    }
#endif

    // split Z = Xa + i Xb (window pre-scaled by 1/2, so no factor here)
    static SELD_HD void unpack(float2 z, float2 p, float2& xa, float2& xb) {
        xa = make_float2(z.x + p.x, z.y - p.y);
        xb = make_float2(z.y + p.y, p.x - z.x);
    }
};

// ---- post-processing arithmetic for one bin (shared by device kernel and host sim) --------------
constexpr float kEpsIV = 1e-8f;
constexpr float kAmin = 1e-10f;      // AmplitudeToDB amin (torchaudio/functional/functional.py:390)
constexpr float kDbPerLog2 = 3.0102999566398120f;  // 10*log10(2)

SELD_HD float power(float2 x) { return x.x * x.x + x.y * x.y; }

// FOA log-mel + IV for one bin: Q = (P0, P1, I1/E, I2/E), R = (P2, P3, I3/E, 0)
template <bool IV>
SELD_HD void bin_features(float2 x0, float2 x1, float2 x2, float2 x3, float4& Q, float4& R) {
    float p0 = power(x0), p1 = power(x1), p2 = power(x2), p3 = power(x3);
    if (IV) {
        float i1 = x0.x * x1.x + x0.y * x1.y;
        float i2 = x0.x * x2.x + x0.y * x2.y;
        float i3 = x0.x * x3.x + x0.y * x3.y;
        float e = kEpsIV + p0 + (p1 + p2 + p3) * (1.0f / 3.0f);
#ifdef __CUDA_ARCH__
        float inv;  // MUFU.RCP alone (1 ulp): e >= 1e-8 is never denormal, and the IV tolerance is 1e-4 relative
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(e));
#else
        float inv = 1.0f / e;
#endif
        Q = make_float4(p0, p1, i1 * inv, i2 * inv);
        R = make_float4(p2, p3, i3 * inv, 0.f);
    } else {
        Q = make_float4(p0, p1, 0.f, 0.f);
        R = make_float4(p2, p3, 0.f, 0.f);
    }
}

SELD_HD float power_to_db(float p) {
#ifdef __CUDA_ARCH__
    float l = __log2f(fmaxf(p, kAmin));
#else
    float l = log2f(fmaxf(p, kAmin));
#endif
    return p <= kAmin ? -100.0f : kDbPerLog2 * l;
}

}  // namespace seld
