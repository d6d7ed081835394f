// K3: GCC-PHAT for the MIC format (north_star kernel 3; SURVEY.md §8(a) A8 — not in the reference, parity
// unpinned).  One warp per (clip, frame):
//   two packed-real forward FFTs (same code as the log-mel kernel) -> unit phasors X^_c[k] of the 4 channels in
//   shared memory; for each of the 3 pairs of microphone pairs {01,02}, {03,12}, {13,23}: cross-spectrum phase
//   G^ = conj(X^_m) X^_n (|.| = 1, or 1+0j when a spectrum bin is exactly 0), two of them packed into one
//   Hermitian-extended complex spectrum G_a + i G_b, ONE inverse complex FFT whose real / imaginary parts are the
//   two correlations, pruned to the 64 lags [-32, 31] (second pass computes 2 of its 32 outputs).
// Output: out[b, t, c_off + pair, lag + 32], pairs in the order 01,02,03,12,13,23.
#include "seld_common.h"
#include "warp_fft.cuh"

namespace seld {

constexpr int kGccWarps = 8;  // per-warp shared memory: two phasor stashes + transpose tile (24.9 KB)

// edge frames (reflect padding): rare, kept out of line so the interior path is 64 plain loads
template <int R1>
__device__ __noinline__ void gcc_load_edge(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                           long long len, int lane) {
    using F = WarpFft<R1>;
#pragma unroll 4
    for (int j = 0; j < R1; ++j) {
        const long long idx = F::reflect(start + lane + 32 * j, len);
        v[j] = make_float2(__ldg(xa + idx), __ldg(xb + idx));
    }
}

template <int R1>
__device__ __forceinline__ void gcc_load_raw(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                             long long len, int lane) {
    using F = WarpFft<R1>;
    if ((start >= 0) && (start + F::N <= len)) {
        const float* pa = xa + start + lane;
        const float* pb = xb + start + lane;
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = make_float2(__ldg(pa + 32 * j), __ldg(pb + 32 * j));
    } else {  // via a scratch array so that v itself never has its address taken
        float2 tmp[R1];
        gcc_load_edge<R1>(tmp, xa, xb, start, len, lane);
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = tmp[j];
    }
}

// x / |x| scaled by keep (1, or 0 for a digitally silent channel); an exactly-zero bin stays (0, 0): the clamp keeps
// rsqrt finite and 0 * finite = 0, so no select is needed
__device__ __forceinline__ float2 unit_phasor(float2 x, float keep) {
    const float p = x.x * x.x + x.y * x.y;
    const float r = rsqrtf(fmaxf(p, 1e-37f)) * keep;
    return make_float2(x.x * r, x.y * r);
}
// conj(a) * b for unit (or zero) phasors; a zero operand means R == 0 -> exp(j*angle(0)) = 1
__device__ __forceinline__ float2 phat(float2 a, float2 b) {
    float2 g = make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
    // unit phasors have |g| ~ 1; g is exactly (0, 0) iff one operand is the zero phasor -> only g.x needs patching
    g.x = (g.x == 0.f && g.y == 0.f) ? 1.f : g.x;
    return g;
}

template <int PP>
__device__ __forceinline__ void pick_pairs(float4 q, float4 s, float2& ga, float2& gb) {
    const float2 x0 = make_float2(q.x, q.y), x1 = make_float2(q.z, q.w);
    const float2 x2 = make_float2(s.x, s.y), x3 = make_float2(s.z, s.w);
    if (PP == 0) { ga = phat(x0, x1); gb = phat(x0, x2); }
    if (PP == 1) { ga = phat(x0, x3); gb = phat(x1, x2); }
    if (PP == 2) { ga = phat(x1, x3); gb = phat(x2, x3); }
}

__global__ void __launch_bounds__(kGccWarps * 32, 1) gcc_phat_kernel(PlanDev p, FeatArgs a) {
    constexpr int R1 = 32;
    using F = WarpFft<R1>;
    constexpr int N = F::N, NB = F::NB;
    constexpr int WARP_F4 = 2 * NB + (F::T_FLOAT2 + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(s_win + N);
    float4* s_w = reinterpret_cast<float4*>(s_tw + R1 * 32);
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_win[i] = p.window[i];
    for (int i = threadIdx.x; i < R1 * 32; i += blockDim.x) s_tw[i] = p.twiddle[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* Q = s_w + warp * WARP_F4;   // unit phasors of channels 0, 1 per bin
    float4* S = Q + NB;                 // unit phasors of channels 2, 3 per bin
    float2* T = reinterpret_cast<float2*>(S + NB);
    const int src = F::partner_lane(lane);
    const float inv_n = 1.0f / float(N);

    const unsigned warps_total = gridDim.x * kGccWarps;
    const unsigned n_items = (unsigned)a.n_items, T_out = (unsigned)a.T_out;
    for (unsigned item = blockIdx.x * kGccWarps + warp; item < n_items; item += warps_total) {
        const unsigned b = item / T_out, t = item - b * T_out;
        const long long len = a.lengths ? a.lengths[b] : a.n_samples;
        const bool valid = (long long)t < 1 + len / p.hop;
        const long long start = valid ? (long long)t * p.hop - F::HALF : 0;
        const float* x = reinterpret_cast<const float*>(a.audio) + (long long)b * a.clip_stride;
        float* out_row = a.out + (((long long)b * a.T_out + t) * a.C_out + a.c_off) * p.n_mels;

        // ---- forward FFTs -> unit phasors ----
#pragma unroll 1
        for (int pair = 0; pair < 2; ++pair) {
            float2 u[32];
            bool sil_a, sil_b;
            {
                float2 v[R1];
                gcc_load_raw<R1>(v, x + (2 * pair) * a.chan_stride, x + (2 * pair + 1) * a.chan_stride, start, len, lane);
                // digitally silent channel?  cheap necessary test first (every lane's first sample is +-0), the OR over
                // all raw samples only when it passes (rare, warp-uniform)
                sil_a = !__any_sync(0xffffffffu, (__float_as_uint(v[0].x) << 1) != 0u);
                sil_b = !__any_sync(0xffffffffu, (__float_as_uint(v[0].y) << 1) != 0u);
                if (sil_a || sil_b) F::silent_channels(v, sil_a, sil_b);
#pragma unroll
                for (int j = 0; j < R1; ++j) {
                    v[j] = cscale(v[j], s_win[lane + 32 * j]);
                }
                F::pass1(v, s_tw + lane);
                __syncwarp();
                F::t_store(v, T, lane);
            }
            __syncwarp();
            F::t_load(u, T, lane);
            __syncwarp();
            F::pass2(u);
            float4* dst = pair ? S : Q;
            const float keep_a = sil_a ? 0.f : 1.f, keep_b = sil_b ? 0.f : 1.f;
            static_for<16>([&](auto KH) {
                constexpr int kh = decltype(KH)::value;
                const float2 z = u[kh], m = u[31 - kh];
                float2 pz;
                pz.x = __shfl_sync(0xffffffffu, m.x, src);
                pz.y = __shfl_sync(0xffffffffu, m.y, src);
                const float2 own = u[(32 - kh) & 31];
                pz.x = lane == 0 ? own.x : pz.x;
                pz.y = lane == 0 ? own.y : pz.y;
                float2 xa, xb;
                F::unpack(z, pz, xa, xb);
                const float2 ua = unit_phasor(xa, keep_a), ub = unit_phasor(xb, keep_b);
                dst[lane + R1 * kh] = make_float4(ua.x, ua.y, ub.x, ub.y);
            });
            if (lane == 0) {
                float2 xa, xb;
                F::unpack(u[16], u[16], xa, xb);
                const float2 ua = unit_phasor(xa, keep_a), ub = unit_phasor(xb, keep_b);
                dst[NB - 1] = make_float4(ua.x, ua.y, ub.x, ub.y);
            }
        }
        __syncwarp();

        // ---- three inverse FFTs, two microphone pairs each ----
        static_for<3>([&](auto PPc) {
            constexpr int PP = decltype(PPc)::value;
            float2 u[32];
            float2 nyq = make_float2(0.f, 0.f);
            if (lane == 0) {
                float2 ga, gb;
                pick_pairs<PP>(Q[NB - 1], S[NB - 1], ga, gb);
                nyq = make_float2(ga.x, gb.x);  // irfft ignores the imaginary part of the Nyquist bin
            }
            float2 mir[16];
            static_for<16>([&](auto KH) {
                constexpr int kh = decltype(KH)::value;
                const int k = lane + R1 * kh;
                float2 ga, gb;
                pick_pairs<PP>(Q[k], S[k], ga, gb);
                float2 g = make_float2(ga.x - gb.y, ga.y + gb.x);     // G_a + i G_b
                mir[kh] = make_float2(ga.x + gb.y, gb.x - ga.y);      // conj(G_a) + i conj(G_b) = bin N-k
                if (kh == 0) {  // lane 0 holds DC there: imaginary parts dropped like irfft
                    g.x = lane == 0 ? ga.x : g.x;
                    g.y = lane == 0 ? gb.x : g.y;
                }
                u[kh] = g;
            });
            static_for<16>([&](auto KH) {
                constexpr int kh = decltype(KH)::value;
                float2 r;
                r.x = __shfl_sync(0xffffffffu, mir[kh].x, src);
                r.y = __shfl_sync(0xffffffffu, mir[kh].y, src);
                const float2 own = kh == 15 ? nyq : mir[(kh + 1) & 15];  // lane 0: register 31-kh is bin 32*(kh+1) mirrored
                r.x = lane == 0 ? own.x : r.x;
                r.y = lane == 0 ? own.y : r.y;
                u[31 - kh] = r;
            });
            F::pass1_inv(u, s_tw + lane);
            __syncwarp();
            F::t_store(u, T, lane);
            __syncwarp();
            float2 w[32];
            F::t_load(w, T, lane);
            __syncwarp();
            float2 pos, neg;
            F::pass2_inv_pruned(w, pos, neg);
            // real part = first pair of the couple, imaginary part = second; lags [-32,-1] then [0,31]
            float* oa = out_row + (2 * PP) * p.n_mels;
            float* ob = out_row + (2 * PP + 1) * p.n_mels;
            oa[lane] = valid ? neg.x * inv_n : 0.f;
            oa[32 + lane] = valid ? pos.x * inv_n : 0.f;
            ob[lane] = valid ? neg.y * inv_n : 0.f;
            ob[32 + lane] = valid ? pos.y * inv_n : 0.f;
        });
    }
}

int configure_gcc_kernels(const seld_plan* plan) {
    (void)plan;
    SELD_CUDA_TRY(cudaFuncSetAttribute(gcc_phat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
    return SELD_OK;
}

int launch_gcc(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    if (plan->dev.r1 != 32) {
        set_error("GCC-PHAT is implemented for n_fft = 1024 only");
        return SELD_ERR_UNSUPPORTED;
    }
    const int NB = plan->dev.n_bins;
    const size_t smem = sizeof(float) * plan->dev.n_fft + sizeof(float2) * 32 * 32 +
                        (size_t)kGccWarps * (2 * NB + (WarpFft<32>::T_FLOAT2 + 1) / 2) * sizeof(float4);
    FeatArgs g = a;
    g.G = 1;
    g.n_items = (long long)a.B * a.T_out;
    long long ctas = (g.n_items + kGccWarps - 1) / kGccWarps;
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    gcc_phat_kernel<<<(unsigned)ctas, kGccWarps * 32, smem, stream>>>(plan->dev, g);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

}  // namespace seld
