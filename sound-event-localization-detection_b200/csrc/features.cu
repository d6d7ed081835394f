// K1+K2 fused: framing + Hann + real FFT (two channels per complex FFT) + power + FOA intensity vectors +
// sparse mel projection + 10*log10, one warp per (clip, frame), persistent CTAs.
// Replaces reference dataset.py:27-58 (audio_to_mel_spectrogram); IV per SURVEY.md §8(a) A7.
//
// Per-warp shared memory: two float4 arrays indexed by bin,
//   Q[k] = (P0, P1, I1/E, I2/E)   (between the two FFTs: stash of the spectra X0, X1)
//   R[k] = (P2, P3, I3/E, 0)      (aliased with the 32x32 transpose tile while an FFT is in flight)
// so the mel gather reads two LDS.128 per filterbank non-zero for all 7 feature channels.
#include "seld_common.h"
#include "warp_fft.cuh"

namespace seld {

__device__ __forceinline__ float ldg_or_zero(const float* p, long long i) { return p ? __ldg(p + i) : 0.f; }

template <int R1>
__device__ __forceinline__ void load_pair(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                          long long len, const float* s_win, int lane) {
    using F = WarpFft<R1>;
    const bool interior = (start >= 0) && (start + F::N <= len);
    if (interior) {
        const float* pa = xa ? xa + start + lane : nullptr;
        const float* pb = xb ? xb + start + lane : nullptr;
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            float w = s_win[lane + 32 * j];
            float a = pa ? __ldg(pa + 32 * j) : 0.f;
            float b = pb ? __ldg(pb + 32 * j) : 0.f;
            v[j] = make_float2(a * w, b * w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            long long idx = F::reflect(start + lane + 32 * j, len);
            float w = s_win[lane + 32 * j];
            v[j] = make_float2(ldg_or_zero(xa, idx) * w, ldg_or_zero(xb, idx) * w);
        }
    }
}

// Full complex FFT of one channel pair; result in u (lane = k_lo, register = k_hi).
template <int R1>
__device__ __forceinline__ void fft_pair(float2 (&u)[32], const float* xa, const float* xb, long long start,
                                         long long len, const float* s_win, const float2* s_tw, float2* T, int lane) {
    using F = WarpFft<R1>;
    {
        float2 v[R1];
        load_pair<R1>(v, xa, xb, start, len, s_win, lane);
        F::pass1(v, s_tw + lane);
        __syncwarp();  // everyone is done reading whatever lived in T / R before
        F::t_store(v, T, lane);
    }
    __syncwarp();
    F::t_load(u, T, lane);
    __syncwarp();  // T may be overwritten (R rows / next transpose) once all lanes have loaded
    F::pass2(u);
}

template <int NCH>
__device__ __forceinline__ void flush_stats(double* stats, int CM, int chan0, int nch, int n_mels, int melA, int melB,
                                            float* st_sum, float* st_sq) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int m = s ? melB : melA;
        if (m < 0) continue;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (c < 4 && c >= nch) continue;
            const int f = (chan0 + c) * n_mels + m;
            atomicAdd(stats + f, (double)st_sum[s * NCH + c]);
            atomicAdd(stats + CM + f, (double)st_sq[s * NCH + c]);
            st_sum[s * NCH + c] = st_sq[s * NCH + c] = 0.f;
        }
    }
}

template <int R1, bool IV, bool STATS>
__global__ void __launch_bounds__(kFeatWarps * 32, 1) features_kernel(PlanDev p, FeatArgs a) {
    using F = WarpFft<R1>;
    constexpr int N = F::N, NB = F::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(s_win + N);
    int2* s_mel = reinterpret_cast<int2*>(s_tw + R1 * 32);
    int* s_midx = reinterpret_cast<int*>(s_mel + (p.la + p.lb) * 32);
    float4* s_qr = reinterpret_cast<float4*>(s_midx + 64);

    for (int i = threadIdx.x; i < N; i += blockDim.x) s_win[i] = p.window[i];
    for (int i = threadIdx.x; i < R1 * 32; i += blockDim.x) s_tw[i] = p.twiddle[i];
    for (int i = threadIdx.x; i < (p.la + p.lb) * 32; i += blockDim.x) s_mel[i] = p.mel_entries[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_midx[i] = p.mel_idx[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float4* Q = s_qr + warp * (2 * NB);
    float4* R = Q + NB;
    float2* T = reinterpret_cast<float2*>(R);
    const int src = F::partner_lane(lane);
    const unsigned full = 0xffffffffu;
    const int melA = s_midx[lane], melB = s_midx[32 + lane];
    const int n_mels = p.n_mels;
    constexpr int NCH = IV ? 7 : 4;

    float st_sum[STATS ? 2 * NCH : 1], st_sq[STATS ? 2 * NCH : 1];
    if (STATS) {
#pragma unroll
        for (int i = 0; i < 2 * NCH; ++i) st_sum[i] = st_sq[i] = 0.f;
    }
    int st_group = -1;  // stats registers belong to one channel group at a time

    // n_items < 2^31 (checked on the host): 32-bit index arithmetic
    const unsigned warps_total = gridDim.x * kFeatWarps;
    const unsigned n_items = (unsigned)a.n_items, T_out32 = (unsigned)a.T_out;
    for (unsigned item = blockIdx.x * kFeatWarps + warp; item < n_items; item += warps_total) {
        const unsigned bg = item / T_out32;
        const long long t = item - bg * T_out32;
        const int g = int(bg % (unsigned)a.G);
        const long long b = bg / (unsigned)a.G;
        const long long len = a.lengths ? a.lengths[b] : a.n_samples;
        const long long T_b = 1 + len / p.hop;
        const int c0 = 4 * g;
        const int nch = min(4, a.C - c0);
        float* out_row = a.out + ((b * a.T_out + t) * a.C_out + a.c_off + c0) * n_mels;

        if (t >= T_b) {  // padding rows of a ragged batch
            const int n_out = IV ? 7 : nch;
            for (int c = 0; c < n_out; ++c) {
                if (melA >= 0) out_row[c * n_mels + melA] = 0.f;
                if (melB >= 0) out_row[c * n_mels + melB] = 0.f;
            }
            continue;
        }

        const float* x = a.audio + b * a.clip_stride + (long long)c0 * a.chan_stride;
        const long long start = t * p.hop - F::HALF;
        float2* spec = a.spec ? a.spec + ((b * a.C + c0) * a.T_out + t) * NB : nullptr;
        const long long spec_cs = a.T_out * NB;  // channel stride of the spectrum dump

        float2 u[32];
        // ---- channel pair (c0, c0+1) ----
        fft_pair<R1>(u, x, nch > 1 ? x + a.chan_stride : nullptr, start, len, s_win, s_tw, T, lane);
#pragma unroll
        for (int kh = 0; kh <= 16; ++kh) {
            if (kh == 16 && lane != 0) break;  // Nyquist bin lives in lane 0 only
            float2 z = u[kh], m = u[31 - (kh & 15)], pz;
            if (kh < 16) {
                pz.x = __shfl_sync(full, m.x, src);
                pz.y = __shfl_sync(full, m.y, src);
                if (lane == 0) pz = u[(32 - kh) & 31];
            } else {
                pz = z;
            }
            float2 x0, x1;
            F::unpack(z, pz, x0, x1);
            const int k = F::bin_of(lane, kh);
            if (lane < R1) {
                Q[k] = make_float4(x0.x, x0.y, x1.x, x1.y);
                if (spec) {
                    spec[k] = x0;
                    if (nch > 1) spec[spec_cs + k] = x1;
                }
            }
        }
        // ---- channel pair (c0+2, c0+3) ----
        const bool have_b = nch > 2;
        if (have_b)
            fft_pair<R1>(u, x + 2 * a.chan_stride, nch > 3 ? x + 3 * a.chan_stride : nullptr, start, len, s_win,
                         s_tw, T, lane);
        else
            __syncwarp();
#pragma unroll
        for (int kh = 0; kh <= 16; ++kh) {
            if (kh == 16 && lane != 0) break;
            float2 x2 = make_float2(0.f, 0.f), x3 = x2;
            if (have_b) {
                float2 z = u[kh], m = u[31 - (kh & 15)], pz;
                if (kh < 16) {
                    pz.x = __shfl_sync(full, m.x, src);
                    pz.y = __shfl_sync(full, m.y, src);
                    if (lane == 0) pz = u[(32 - kh) & 31];
                } else {
                    pz = z;
                }
                F::unpack(z, pz, x2, x3);
            }
            const int k = F::bin_of(lane, kh);
            if (lane < R1) {
                float4 s = Q[k];
                float4 q, r;
                bin_features<IV>(make_float2(s.x, s.y), make_float2(s.z, s.w), x2, x3, q, r);
                Q[k] = q;
                R[k] = r;
                if (spec && have_b) {
                    spec[2 * spec_cs + k] = x2;
                    if (nch > 3) spec[3 * spec_cs + k] = x3;
                }
            }
        }
        __syncwarp();

        // ---- sparse mel projection: slot A then slot B of this lane ----
        float acc[2][NCH];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[s][c] = 0.f;
        const int2* e = s_mel + lane;
#pragma unroll 4
        for (int i = 0; i < p.la; ++i, e += 32) {
            const int2 en = *e;
            const float w = __int_as_float(en.y);
            const float4 q = Q[en.x], r = R[en.x];
            acc[0][0] = fmaf(w, q.x, acc[0][0]);
            acc[0][1] = fmaf(w, q.y, acc[0][1]);
            acc[0][2] = fmaf(w, r.x, acc[0][2]);
            acc[0][3] = fmaf(w, r.y, acc[0][3]);
            if (IV) {
                acc[0][4] = fmaf(w, q.z, acc[0][4]);
                acc[0][5] = fmaf(w, q.w, acc[0][5]);
                acc[0][6] = fmaf(w, r.z, acc[0][6]);
            }
        }
#pragma unroll 4
        for (int i = 0; i < p.lb; ++i, e += 32) {
            const int2 en = *e;
            const float w = __int_as_float(en.y);
            const float4 q = Q[en.x], r = R[en.x];
            acc[1][0] = fmaf(w, q.x, acc[1][0]);
            acc[1][1] = fmaf(w, q.y, acc[1][1]);
            acc[1][2] = fmaf(w, r.x, acc[1][2]);
            acc[1][3] = fmaf(w, r.y, acc[1][3]);
            if (IV) {
                acc[1][4] = fmaf(w, q.z, acc[1][4]);
                acc[1][5] = fmaf(w, q.w, acc[1][5]);
                acc[1][6] = fmaf(w, r.z, acc[1][6]);
            }
        }

        // ---- log + store: out[b, t, c_off + c0 + c, m] ----
        if (STATS && st_group != g) {  // only when C > 4: partial sums belong to one channel group at a time
            if (st_group >= 0)
                flush_stats<NCH>(a.stats, a.C_out * n_mels, a.c_off + 4 * st_group, min(4, a.C - 4 * st_group), n_mels,
                                 melA, melB, st_sum, st_sq);
            st_group = g;
        }
        const bool in_stats = STATS && (t < (a.stat_frames ? (long long)a.stat_frames[b] : T_b));
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int m = s ? melB : melA;
            if (m < 0) continue;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (c < 4 && c >= nch) continue;
                const float v = c < 4 ? power_to_db(acc[s][c]) : acc[s][c];
                out_row[c * n_mels + m] = v;
                if (STATS && in_stats) {
                    st_sum[s * NCH + c] += v;
                    st_sq[s * NCH + c] = fmaf(v, v, st_sq[s * NCH + c]);
                }
            }
        }
    }

    if (STATS && st_group >= 0)
        flush_stats<NCH>(a.stats, a.C_out * n_mels, a.c_off + 4 * st_group, min(4, a.C - 4 * st_group), n_mels, melA,
                         melB, st_sum, st_sq);
}

template <int R1, bool IV, bool STATS>
static int launch_one(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    auto kern = features_kernel<R1, IV, STATS>;
    SELD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->feat_smem));
    long long ctas = (a.n_items + kFeatWarps - 1) / kFeatWarps;
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    kern<<<(unsigned)ctas, kFeatWarps * 32, plan->feat_smem, stream>>>(plan->dev, a);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

int launch_features(const seld_plan* plan, bool iv, const FeatArgs& a, cudaStream_t stream) {
    const bool st = a.stats != nullptr;
    if (plan->dev.r1 == 32) {
        if (iv) return st ? launch_one<32, true, true>(plan, a, stream) : launch_one<32, true, false>(plan, a, stream);
        return st ? launch_one<32, false, true>(plan, a, stream) : launch_one<32, false, false>(plan, a, stream);
    }
    if (iv) return st ? launch_one<30, true, true>(plan, a, stream) : launch_one<30, true, false>(plan, a, stream);
    return st ? launch_one<30, false, true>(plan, a, stream) : launch_one<30, false, false>(plan, a, stream);
}

}  // namespace seld
