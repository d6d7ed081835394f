// K1+K2 fused: framing + Hann + real FFT (two channels per complex FFT) + power + FOA intensity vectors +
// sparse mel projection + 10*log10, one warp per (clip, frame), persistent CTAs (one per SM).
// Replaces reference dataset.py:27-58 (audio_to_mel_spectrogram); IV per SURVEY.md §8(a) A7.
//
// Per-warp shared memory (16 656 B): two float4 arrays indexed by bin,
//   Q[k] = (P0, P1, I1/E, I2/E)   (between the two FFTs: stash of the spectra X0, X1)
//   R[k] = (P2, P3, I3/E, 0)      (aliased with the 32x33 transpose tile while an FFT is in flight)
// so the mel gather reads two LDS.128 per filterbank non-zero for all 7 feature channels.
//
// Latency hiding with only 12 warps per SM (shared memory bound) is done by register prefetch: the raw
// samples of channel pair b are requested before the post-processing of pair a, and the samples of the NEXT
// frame's pair a before the mel gather of the current frame, when the FFT registers are dead.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "features_fast.cuh"
#include "seld_common.h"
#include "warp_fft.cuh"

namespace seld {

template <int R1>
__device__ __forceinline__ void load_raw(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                         long long len, int lane) {
    using F = WarpFft<R1>;
    const bool interior = (start >= 0) && (start + F::N <= len);
    if (interior) {
        const float* pa = xa + start + lane;
        if (xb) {
            const float* pb = xb + start + lane;
#pragma unroll
            for (int j = 0; j < R1; ++j) v[j] = make_float2(__ldg(pa + 32 * j), __ldg(pb + 32 * j));
        } else {
#pragma unroll
            for (int j = 0; j < R1; ++j) v[j] = make_float2(__ldg(pa + 32 * j), 0.f);
        }
    } else if (len <= F::HALF) {  // too short for reflect padding: nothing is read, the rows are written as 0
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = make_float2(0.f, 0.f);
    } else {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const long long idx = F::reflect(start + lane + 32 * j, len);
            v[j] = make_float2(__ldg(xa + idx), xb ? __ldg(xb + idx) : 0.f);
        }
    }
}

// window -> pass 1 -> transpose -> pass 2; result in u (lane = k_lo, register = k_hi).
// Block floating point as in the fast kernel (features_fast.cuh): max |windowed x| of both channels over the warp; 0 <=>
// digitally silent (bits_* = 0); when the binary exponents differ by >= 4 the quieter channel is multiplied by the exact power of
// two that brings it to its partner's level, and inv_a / inv_b (applied to the split spectra by the caller) undo it.
template <int R1>
__device__ __forceinline__ void fft_from_raw(float2 (&u)[32], float2 (&v)[R1], const float* s_win, const float2* s_tw,
                                             float2* T, int lane, unsigned& bits_a, unsigned& bits_b, float& inv_a,
                                             float& inv_b) {
    using F = WarpFft<R1>;
    float ma = 0.f, mb = 0.f;
#pragma unroll
    for (int j = 0; j < R1; ++j) {  // window first: the level that matters is that of what enters the FFT
        const float w = s_win[lane + 32 * j];
        v[j].x *= w;
        v[j].y *= w;
        ma = fmaxf(ma, fabsf(v[j].x));
        mb = fmaxf(mb, fabsf(v[j].y));
    }
    bits_a = __reduce_max_sync(0xffffffffu, __float_as_uint(ma));
    bits_b = __reduce_max_sync(0xffffffffu, __float_as_uint(mb));
    int sh = (int)(bits_a >> 23) - (int)(bits_b >> 23);
    sh = (bits_a == 0u || bits_b == 0u) ? 0 : sh;
    sh = (sh > -4 && sh < 4) ? 0 : max(-60, min(60, sh));
    const int sa = sh < 0 ? -sh : 0, sb = sh > 0 ? sh : 0;
    const float fa = __uint_as_float((unsigned)(127 + sa) << 23), fb = __uint_as_float((unsigned)(127 + sb) << 23);
    inv_a = __uint_as_float((unsigned)(127 - sa) << 23);
    inv_b = __uint_as_float((unsigned)(127 - sb) << 23);
    if (sh != 0) {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            v[j].x *= fa;
            v[j].y *= fb;
        }
    }
    F::pass1(v, s_tw + lane);
    __syncwarp();  // every lane is done reading whatever lived in T / R before
    F::t_store(v, T, lane);
    __syncwarp();
    F::t_load(u, T, lane);
    __syncwarp();  // T may be overwritten (R rows / next transpose) once all lanes have loaded
    F::pass2(u);
}

// mirror-bin exchange + channel split for register k_hi (0..15) of every lane
template <int R1, int KH>
__device__ __forceinline__ void split_pair(const float2 (&u)[32], int lane, int src, float2& xa, float2& xb) {
    const float2 z = u[KH], m = u[31 - KH];
    float2 p;
    p.x = __shfl_sync(0xffffffffu, m.x, src);
    p.y = __shfl_sync(0xffffffffu, m.y, src);
    const float2 own = u[(32 - KH) & 31];  // lane 0 holds its own mirror bins
    p.x = lane == 0 ? own.x : p.x;
    p.y = lane == 0 ? own.y : p.y;
    WarpFft<R1>::unpack(z, p, xa, xb);
}
__device__ __forceinline__ float2 fscale(float2 x, float f) { return make_float2(x.x * f, x.y * f); }

struct ItemCtx {
    const float* x;     // channel c0 of the clip
    float* out_row;     // out[b, t, c_off + c0, 0]
    float2* spec;       // spec[b, c0, t, 0] or null
    long long start, len;
    int nch;
    bool valid;         // t < T_b (padding rows of a ragged batch are written as 0)
};

template <int R1, bool IV, bool SPEC>
__global__ void __launch_bounds__(kFeatWarps * 32, 1) features_kernel(PlanDev p, FeatArgs a) {
    const int WARPS = blockDim.x >> 5;  // chosen at plan creation from the shared-memory budget (<= kFeatWarps)
    using F = WarpFft<R1>;
    constexpr int N = F::N, NB = F::NB;
    constexpr int NCH = IV ? 7 : 4;
    constexpr int WARP_F4 = NB + (F::T_FLOAT2 > 2 * NB ? (F::T_FLOAT2 + 1) / 2 : NB);  // Q rows + max(R rows, T tile)
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
    float4* Q = s_qr + warp * WARP_F4;
    float4* R = Q + NB;
    float2* T = reinterpret_cast<float2*>(R);
    const unsigned char* Qb = reinterpret_cast<const unsigned char*>(Q);
    const int src = F::partner_lane(lane);
    const int melA = s_midx[lane], melB = s_midx[32 + lane];
    const int n_mels = p.n_mels;
    const bool active = lane < R1;  // lanes >= R1 (n_fft 960) own no bins

    // ---- work distribution: item = (b*G + g)*T_out + t, items of one warp are warps_total apart ----
    const unsigned warps_total = gridDim.x * WARPS;
    const unsigned n_items = (unsigned)a.n_items, T_out = (unsigned)a.T_out;
    unsigned item = blockIdx.x * WARPS + warp;
    if (item >= n_items) return;
    unsigned bg = item / T_out, t = item - bg * T_out;
    long long len_cache_b = -1, len_cache = a.n_samples;

    auto make_ctx = [&](unsigned bg_, unsigned t_) {
        ItemCtx c;
        const unsigned g = a.G == 1 ? 0u : bg_ % (unsigned)a.G;
        const long long b = a.G == 1 ? (long long)bg_ : (long long)(bg_ / (unsigned)a.G);
        if (a.lengths && b != len_cache_b) {
            len_cache = a.lengths[b];
            len_cache_b = b;
        }
        c.len = len_cache;
        const int c0 = 4 * (int)g;
        c.nch = min(4, a.C - c0);
        const bool too_short = c.len <= F::HALF;  // reflect padding undefined (torch raises): rows written as 0 + status bit
        if (a.lengths && too_short && t_ == 0 && lane == 0) atomicOr(a.status, 1);
        c.valid = (long long)t_ < 1 + c.len / p.hop && !too_short;
        c.start = c.valid ? (long long)t_ * p.hop - F::HALF : 0;  // padding rows read frame 0 and are zeroed at the store
        c.x = reinterpret_cast<const float*>(a.audio) + b * a.clip_stride + (long long)c0 * a.chan_stride;
        c.out_row = a.out + ((b * a.T_out + t_) * a.C_out + a.c_off + c0) * n_mels;
        c.spec = SPEC ? a.spec + ((b * a.C + c0) * a.T_out + t_) * NB : nullptr;
        return c;
    };

    ItemCtx cur = make_ctx(bg, t);
    float2 v[R1];
    load_raw<R1>(v, cur.x, cur.nch > 1 ? cur.x + a.chan_stride : nullptr, cur.start, cur.len, lane);

    while (true) {
        float2 u[32];
        const long long spec_cs = a.T_out * NB;  // channel stride of the spectrum dump
        // ================= channel pair (c0, c0+1) =================
        unsigned bits_a, bits_b;
        float inv_a, inv_b;
        fft_from_raw<R1>(u, v, s_win, s_tw, T, lane, bits_a, bits_b, inv_a, inv_b);
        const bool have_b = cur.nch > 2;
        if (have_b)  // request pair b now; it lands while pair a is post-processed
            load_raw<R1>(v, cur.x + 2 * a.chan_stride, cur.nch > 3 ? cur.x + 3 * a.chan_stride : nullptr, cur.start,
                         cur.len, lane);
        static_for<16>([&](auto KH) {
            constexpr int kh = decltype(KH)::value;
            float2 x0, x1;
            split_pair<R1, kh>(u, lane, src, x0, x1);
            x0 = fscale(x0, inv_a);
            x1 = fscale(x1, inv_b);
            const int k = lane + R1 * kh;
            if (active) {
                Q[k] = make_float4(x0.x, x0.y, x1.x, x1.y);
                if (SPEC) {
                    cur.spec[k] = x0;
                    if (cur.nch > 1) cur.spec[spec_cs + k] = x1;
                }
            }
        });
        if (lane == 0) {  // Nyquist bin: its own mirror
            float2 x0, x1;
            F::unpack(u[16], u[16], x0, x1);
            x0 = fscale(x0, inv_a);
            x1 = fscale(x1, inv_b);
            Q[NB - 1] = make_float4(x0.x, x0.y, x1.x, x1.y);
            if (SPEC) {
                cur.spec[NB - 1] = x0;
                if (cur.nch > 1) cur.spec[spec_cs + NB - 1] = x1;
            }
        }
        // A silent channel must come out as an exactly-zero spectrum like the reference's separate FFT; the
        // split leaves the rounding asymmetry of the other channel (~1e-7 relative) in it.
        const bool sil_a = bits_a == 0u, sil_b = bits_b == 0u;
        if (sil_a || sil_b) {  // rare (warp-uniform)
            const float ka = sil_a ? 0.f : 1.f, kb = sil_b ? 0.f : 1.f;
            __syncwarp();
            for (int k = lane; k < NB; k += 32) {
                float4 s = Q[k];
                Q[k] = make_float4(s.x * ka, s.y * ka, s.z * kb, s.w * kb);
                if (SPEC) {
                    cur.spec[k] = make_float2(s.x * ka, s.y * ka);
                    if (cur.nch > 1) cur.spec[spec_cs + k] = make_float2(s.z * kb, s.w * kb);
                }
            }
            __syncwarp();
        }
        // ================= channel pair (c0+2, c0+3) =================
        unsigned bits_c = 0u, bits_d = 0u;
        float inv_c = 1.f, inv_d = 1.f;
        if (have_b) {
            fft_from_raw<R1>(u, v, s_win, s_tw, T, lane, bits_c, bits_d, inv_c, inv_d);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = make_float2(0.f, 0.f);
            __syncwarp();
        }
        static_for<16>([&](auto KH) {
            constexpr int kh = decltype(KH)::value;
            float2 x2, x3;
            split_pair<R1, kh>(u, lane, src, x2, x3);
            x2 = fscale(x2, inv_c);
            x3 = fscale(x3, inv_d);
            const int k = lane + R1 * kh;
            if (active) {
                const float4 s = Q[k];
                float4 q, r;
                bin_features<IV>(make_float2(s.x, s.y), make_float2(s.z, s.w), x2, x3, q, r);
                Q[k] = q;
                R[k] = r;
                if (SPEC && have_b) {
                    cur.spec[2 * spec_cs + k] = x2;
                    if (cur.nch > 3) cur.spec[3 * spec_cs + k] = x3;
                }
            }
        });
        if (lane == 0) {
            float2 x2, x3;
            F::unpack(u[16], u[16], x2, x3);
            x2 = fscale(x2, inv_c);
            x3 = fscale(x3, inv_d);
            const float4 s = Q[NB - 1];
            float4 q, r;
            bin_features<IV>(make_float2(s.x, s.y), make_float2(s.z, s.w), x2, x3, q, r);
            Q[NB - 1] = q;
            R[NB - 1] = r;
            if (SPEC && have_b) {
                cur.spec[2 * spec_cs + NB - 1] = x2;
                if (cur.nch > 3) cur.spec[3 * spec_cs + NB - 1] = x3;
            }
        }
        const bool sil_c = bits_c == 0u, sil_d = bits_d == 0u;
        if (have_b && (sil_c || sil_d)) {  // rare (warp-uniform): redo the rows with the silent channel at exactly 0
            __syncwarp();
            for (int k = lane; k < NB; k += 32) {
                const float4 q = Q[k], r = R[k];
                const float p2 = sil_c ? 0.f : r.x, p3 = sil_d ? 0.f : r.y;
                float4 qn = make_float4(q.x, q.y, 0.f, 0.f), rn = make_float4(p2, p3, 0.f, 0.f);
                if (IV) {
                    const float e_old = kEpsIV + q.x + (q.y + r.x + r.y) * (1.0f / 3.0f);
                    const float e_new = kEpsIV + q.x + (q.y + p2 + p3) * (1.0f / 3.0f);
                    const float g = e_old / e_new;  // numerators I_c = (I_c / E_old) * E_old
                    qn.z = q.z * g;
                    qn.w = sil_c ? 0.f : q.w * g;
                    rn.z = sil_d ? 0.f : r.z * g;
                }
                Q[k] = qn;
                R[k] = rn;
                if (SPEC) {
                    if (sil_c) cur.spec[2 * spec_cs + k] = make_float2(0.f, 0.f);
                    if (sil_d && cur.nch > 3) cur.spec[3 * spec_cs + k] = make_float2(0.f, 0.f);
                }
            }
        }

        // ================= next item: request its pair a while the mel gather runs =================
        t += warps_total;
        while (t >= T_out) {
            t -= T_out;
            ++bg;
        }
        item += warps_total;
        const bool more = item < n_items;
        const ItemCtx done = cur;
        if (more) {
            cur = make_ctx(bg, t);
            load_raw<R1>(v, cur.x, cur.nch > 1 ? cur.x + a.chan_stride : nullptr, cur.start, cur.len, lane);
        }
        __syncwarp();  // Q / R rows of all lanes are visible

        // ================= sparse mel projection: slot A then slot B of this lane =================
        float acc[2][NCH];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc[s][c] = 0.f;
        const int2* e = s_mel + lane;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int n_it = s ? p.lb : p.la;  // multiples of 4 (host pads)
#pragma unroll 4
            for (int i = 0; i < n_it; ++i, e += 32) {
                const int2 en = *e;  // {byte offset of the bin row, weight bits}
                const float w = __int_as_float(en.y);
                const float4 q = *reinterpret_cast<const float4*>(Qb + en.x);
                const float4 r = *reinterpret_cast<const float4*>(Qb + en.x + NB * 16);
                acc[s][0] = fmaf(w, q.x, acc[s][0]);
                acc[s][1] = fmaf(w, q.y, acc[s][1]);
                acc[s][2] = fmaf(w, r.x, acc[s][2]);
                acc[s][3] = fmaf(w, r.y, acc[s][3]);
                if (IV) {
                    acc[s][4] = fmaf(w, q.z, acc[s][4]);
                    acc[s][5] = fmaf(w, q.w, acc[s][5]);
                    acc[s][6] = fmaf(w, r.z, acc[s][6]);
                }
            }
        }

        // ================= log + store: out[b, t, c_off + c0 + c, m] =================
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int m = s ? melB : melA;
            if (m >= 0) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (c < 4 && c >= done.nch) continue;
                    const float val = c < 4 ? power_to_db(acc[s][c]) : acc[s][c];
                    done.out_row[c * n_mels + m] = done.valid ? val : 0.f;
                }
            }
        }
        if (!more) break;
    }
}

// ---- scaler partials: per-feature sum and sum of squares over the kept frames ----
// x (B, T_out, F) float32; for the columns [col0, col0 + ncols): stats[f] += sum, stats[F + f] += sum of
// squares over rows t < frames[b] (or t < 1 + len_b/hop when frames is null).
// CTA = 64 columns x 8 row lanes over a slab of kStatRows rows: every warp reads 128 contiguous bytes per
// row, 4 rows in flight per thread; float64 accumulation throughout.
constexpr int kStatRows = 256;
__global__ void __launch_bounds__(512) feature_stats_kernel(const float* __restrict__ x, long long T_out, int F,
                                                            int col0, int ncols, int B, const int* __restrict__ frames,
                                                            const long long* __restrict__ lengths, long long n_samples,
                                                            int hop, double* __restrict__ stats) {
    __shared__ unsigned char s_valid[kStatRows];
    __shared__ double s_sum[8][64], s_sq[8][64];
    const long long total_rows = (long long)B * T_out;
    const long long r0 = (long long)blockIdx.x * kStatRows;
    const int n_rows = (int)min((long long)kStatRows, total_rows - r0);
    const int tid = threadIdx.y * 64 + threadIdx.x;
    if (tid < kStatRows) {
        bool ok = false;
        if (tid < n_rows) {
            const long long r = r0 + tid, b = r / T_out, t = r - b * T_out;
            const long long lim = frames ? (long long)frames[b] : 1 + (lengths ? lengths[b] : n_samples) / hop;
            ok = t < lim;
        }
        s_valid[tid] = ok;
    }
    __syncthreads();
    const int j = blockIdx.y * 64 + threadIdx.x;
    double cs = 0.0, css = 0.0;  // float64 throughout: std = sqrt(E[x^2] - mean^2) amplifies the error of the sums
    if (j < ncols) {
        const float* px = x + r0 * F + col0 + j;
#pragma unroll 4
        for (int r = threadIdx.y; r < n_rows; r += 8) {
            const double v = s_valid[r] ? (double)__ldg(px + (long long)r * F) : 0.0;
            cs += v;
            css = fma(v, v, css);
        }
    }
    s_sum[threadIdx.y][threadIdx.x] = cs;
    s_sq[threadIdx.y][threadIdx.x] = css;
    __syncthreads();
    if (threadIdx.y == 0 && j < ncols) {
        double s = 0.0, ss = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += s_sum[i][threadIdx.x];
            ss += s_sq[i][threadIdx.x];
        }
        atomicAdd(stats + col0 + j, s);
        atomicAdd(stats + F + col0 + j, ss);
    }
}

// 128-bit variant: a thread owns 4 adjacent columns and every 8th row of a 128-row slab, 8 loads of 16 bytes in
// flight per thread (the scalar kernel above keeps too few bytes in flight to fill HBM).
constexpr int kStatRows4 = 128;
__global__ void __launch_bounds__(1024) feature_stats_kernel_v4(const float* __restrict__ x, long long T_out, int F,
                                                                int col0, int ncols4, int B, const int* __restrict__ frames,
                                                                const long long* __restrict__ lengths, long long n_samples,
                                                                int hop, double* __restrict__ stats) {
    extern __shared__ double s_part[];  // [8][ncols4][8]
    __shared__ unsigned char s_valid[2][kStatRows4];
    const long long total_rows = (long long)B * T_out;
    const long long n_slabs = (total_rows + kStatRows4 - 1) / kStatRows4;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int j = threadIdx.x;  // column quad
    const int F4 = F / 4;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int buf = 0;
    // persistent CTAs: the block reduction and the atomics happen once per CTA, not once per slab
    for (long long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x, buf ^= 1) {
        const long long r0 = slab * kStatRows4;
        const int n_rows = (int)min((long long)kStatRows4, total_rows - r0);
        if (tid < kStatRows4) {
            bool ok = false;
            if (tid < n_rows) {
                const long long r = r0 + tid, b = r / T_out, t = r - b * T_out;
                const long long lim = frames ? (long long)frames[b] : 1 + (lengths ? lengths[b] : n_samples) / hop;
                ok = t < lim;
            }
            s_valid[buf][tid] = ok;
        }
        __syncthreads();  // (double-buffered flags: one barrier per slab is enough)
        if (j < ncols4) {
            const float4* px = reinterpret_cast<const float4*>(x + r0 * F + col0) + j;
#pragma unroll 8
            for (int r = threadIdx.y; r < n_rows; r += 8) {
                float4 v = __ldg(px + (long long)r * F4);
                if (!s_valid[buf][r]) v = make_float4(0.f, 0.f, 0.f, 0.f);
                const double a0 = v.x, a1 = v.y, a2 = v.z, a3 = v.w;
                acc[0] += a0; acc[1] += a1; acc[2] += a2; acc[3] += a3;
                acc[4] = fma(a0, a0, acc[4]); acc[5] = fma(a1, a1, acc[5]);
                acc[6] = fma(a2, a2, acc[6]); acc[7] = fma(a3, a3, acc[7]);
            }
        }
    }
    if (j < ncols4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_part[(threadIdx.y * ncols4 + j) * 8 + i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.y == 0 && j < ncols4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double t = 0.0;
#pragma unroll
            for (int y = 0; y < 8; ++y) t += s_part[(y * ncols4 + j) * 8 + i];
            atomicAdd(stats + (i < 4 ? 0 : F) + col0 + 4 * j + (i & 3), t);
        }
    }
}

template <int R1, bool IV, bool SPEC>
static int launch_w(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    const int warps = plan->generic_warps;  // from the shared-memory budget of this plan's tables (seld_plan_create)
    const size_t smem = plan->table_bytes + (size_t)warps * plan->warp_smem;
    long long ctas = (a.n_items + warps - 1) / warps;
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    features_kernel<R1, IV, SPEC><<<(unsigned)ctas, warps * 32, smem, stream>>>(plan->dev, a);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

template <int R1, bool IV, bool SPEC>
static int configure_w(const seld_plan* plan) {
    (void)plan;  // the attribute is a per-function maximum: opt in to the whole 227 KB once, every plan fits below it
    SELD_CUDA_TRY(cudaFuncSetAttribute(features_kernel<R1, IV, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
    return SELD_OK;
}

int launch_feature_stats(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    const int F = a.C_out * plan->dev.n_mels;
    const long long rows = (long long)a.B * a.T_out;
    const int ncols = a.n_out * plan->dev.n_mels;
    if (rows < 1 || ncols < 1) return SELD_OK;
    const int col0 = a.c_off * plan->dev.n_mels;
    if (F % 4 == 0 && col0 % 4 == 0 && ncols % 4 == 0 && ncols / 4 <= 128 &&
        (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) {
        const int ncols4 = ncols / 4;
        const long long slabs = (rows + kStatRows4 - 1) / kStatRows4;
        dim3 grid4((unsigned)std::min<long long>(slabs, 2ll * plan->num_sms)), block4((unsigned)((ncols4 + 31) / 32 * 32), 8);
        const size_t smem = sizeof(double) * 8 * ncols4 * 8;
        feature_stats_kernel_v4<<<grid4, block4, smem, stream>>>(a.out, a.T_out, F, col0, ncols4, a.B, a.stat_frames,
                                                                  a.lengths, a.n_samples, plan->dev.hop, a.stats);
        SELD_CUDA_TRY(cudaGetLastError());
        return SELD_OK;
    }
    dim3 grid((unsigned)((rows + kStatRows - 1) / kStatRows), (unsigned)((ncols + 63) / 64)), block(64, 8);
    feature_stats_kernel<<<grid, block, 0, stream>>>(a.out, a.T_out, F, a.c_off * plan->dev.n_mels, ncols, a.B,
                                                     a.stat_frames, a.lengths, a.n_samples, plan->dev.hop, a.stats);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

// Does the caller's filterbank equal the baked one bit for bit?  (host, at plan creation)
template <int NFFT>
static bool fb_matches(const float* fb, int n_mels) {
    using MB = MelBaked<NFFT>;
    if (n_mels != MB::N_MELS) return false;
    for (int k = 0; k < MB::NB; ++k)
        for (int m = 0; m < n_mels; ++m) {
            float want = 0.f;
            if (m == MB::m0[k]) want = MB::w0[k];
            if (m == MB::m1[k]) want = MB::w1[k];
            if (fb[(size_t)k * n_mels + m] != want) return false;
        }
    return true;
}
bool fast_filterbank_matches(int n_fft, const float* fb, int n_mels) {
    return n_fft == 1024 ? fb_matches<1024>(fb, n_mels) : n_fft == 960 ? fb_matches<960>(fb, n_mels) : false;
}

// One-time kernel configuration for a plan's device (cudaFuncSetAttribute is per device): called by seld_plan_create
// so that the hot calls do no host work besides the launch.
int configure_feature_kernels(const seld_plan* plan) {
    int rc = SELD_OK;
    auto acc = [&](int r) { if (rc == SELD_OK) rc = r; };
    {
        cudaError_t e = cudaFuncSetAttribute(feature_stats_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 128 * 8 * (int)sizeof(double));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(feature_stats_kernel_v4)");
    }
    if (plan->dev.r1 == 32) {
        acc(configure_w<32, true, true>(plan)); acc(configure_w<32, true, false>(plan));
        acc(configure_w<32, false, true>(plan)); acc(configure_w<32, false, false>(plan));
        if (plan->v3_ok) { acc(fast_configure_r32_f32()); acc(fast_configure_r32_i16()); }
    } else {
        acc(configure_w<30, true, true>(plan)); acc(configure_w<30, true, false>(plan));
        acc(configure_w<30, false, true>(plan)); acc(configure_w<30, false, false>(plan));
        if (plan->v3_ok) { acc(fast_configure_r30_f32()); acc(fast_configure_r30_i16()); }
    }
    return rc;
}

// Redo list of the stream a call runs on.  Calls on one stream are ordered, so a slot is never shared by two calls in
// flight; the first kRedoSlots streams that use a plan get a slot each, later ones get none (block floating on every
// frame then, ~8 % slower).  Host-only bookkeeping: a mutex and a 16-entry table, no CUDA call.
static_assert(kRedoCap == (1 << 16), "abi.cu sizes the redo lists with this capacity");
unsigned* plan_redo_slot(seld_plan* plan, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(plan->slot_mutex));
    for (int i = 0; i < plan->n_slots; ++i)
        if (plan->slot_stream[i] == (void*)stream) return plan->d_redo + (size_t)i * (4 + kRedoCap);
    if (plan->n_slots >= kRedoSlots) return nullptr;
    plan->slot_stream[plan->n_slots] = (void*)stream;
    return plan->d_redo + (size_t)(plan->n_slots++) * (4 + kRedoCap);
}

// fast path: the reference's own configurations (4 channels, baked 64-mel HTK bank), 16-byte aligned rows
bool fast_path_ok(const seld_plan* plan, const FeatArgs& a) {
    return plan->v3_ok && !plan->force_generic && a.C == 4 && a.spec == nullptr &&
           (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

int launch_features(const seld_plan* plan, bool iv, const FeatArgs& a, cudaStream_t stream) {
    const bool sp = a.spec != nullptr;
    const bool ext = a.mean || a.inv_std || a.out_ctf || a.out_bf16;
    int rc;
    if (fast_path_ok(plan, a)) {
        if (ext && a.stats) {  // the partials are those of the raw float32 features
            set_error("seld_features_ex: d_stats cannot be combined with normalisation / layout / bf16 options");
            return SELD_ERR_UNSUPPORTED;
        }
        const int epi = ext ? 2 : 0;
        const int warps = a.in_i16 ? 12 : plan->fast_warps;
        auto run = [&](bool bf, const FeatArgs& fa) {
            if (plan->dev.r1 == 32)
                return fa.in_i16 ? fast_launch_r32_i16(plan, iv, epi, 12, bf, fa, stream)
                                 : fast_launch_r32_f32(plan, iv, epi, warps, bf, fa, stream);
            return fa.in_i16 ? fast_launch_r30_i16(plan, iv, epi, 12, bf, fa, stream)
                             : fast_launch_r30_f32(plan, iv, epi, warps, bf, fa, stream);
        };
        FeatArgs fa = a;
        fa.redo = plan_redo_slot(const_cast<seld_plan*>(plan), stream);
        if (plan->force_bf || !fa.redo) {  // A/B switch, or more streams than slots: block floating on every frame
            fa.redo_mode = 0;
            rc = run(true, fa);
        } else {
            fa.redo_mode = 0;
            rc = run(false, fa);           // lean kernel: all frames, flags the ones whose channel pairs differ too much
            if (rc != SELD_OK) return rc;
            fa.redo_mode = 1;
            rc = run(true, fa);            // block-floating kernel: the same (persistent) grid over the flagged frames
        }
        if (rc != SELD_OK) return rc;
        return a.stats ? launch_feature_stats(plan, a, stream) : SELD_OK;
    }
    if (a.in_i16 || ext) {
        set_error("seld_features_ex: int16 input and the output options need the fast path (4 channels, 64 HTK mels, "
                  "no spectrum dump, 16-byte aligned output)");
        return SELD_ERR_UNSUPPORTED;
    }
    if (plan->dev.r1 == 32) {
        if (iv) rc = sp ? launch_w<32, true, true>(plan, a, stream) : launch_w<32, true, false>(plan, a, stream);
        else rc = sp ? launch_w<32, false, true>(plan, a, stream) : launch_w<32, false, false>(plan, a, stream);
    } else {
        if (iv) rc = sp ? launch_w<30, true, true>(plan, a, stream) : launch_w<30, true, false>(plan, a, stream);
        else rc = sp ? launch_w<30, false, true>(plan, a, stream) : launch_w<30, false, false>(plan, a, stream);
    }
    if (rc != SELD_OK) return rc;
    if (a.stats) return launch_feature_stats(plan, a, stream);
    return SELD_OK;
}

}  // namespace seld
