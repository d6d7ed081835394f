// K1+K2+K3 fused for the MIC format (north_star kernel 3; SURVEY.md §8(a) A8 — GCC-PHAT is not in the reference, parity
// unpinned): ONE launch writes the 4 log-mel channels (reference dataset.py:27-58) and the 6 GCC-PHAT channels
//   out[b, t, c_off + 0..3, mel]          10 log10(max(mel power, 1e-10))
//   out[b, t, c_off + 4 + pair, lag + 32]  irfft(exp(j angle(conj(X_m) X_n)))[lag], lag in [-32, 31], pairs 01 02 03 12 13 23
// from five transforms per frame instead of the seven of round 1 (the log-mel launch and the GCC launch each ran the
// two forward FFTs): two packed forward FFTs shared by both feature sets, three packed inverse FFTs.
//
// One warp per frame, groups of four warps for the mel phase, exactly like the FOA kernel (features_fast.cuh):
//   pair a : window, block-floating level equalisation, packed FFT, split -> X0, X1; scaled powers -> planes P0, P1;
//            unit phasors (U0, U1) of the lane's 16 bins parked in TENSOR MEMORY (64 words per lane, two STTM.x32; the
//            same lane reads them back: no synchronisation; lane 0's Nyquist bin goes to four plane pad words)
//   pair b : the same -> P2, P3 planes; unit phasors (U2, U3) -> 64 more tensor-memory words per lane
//   3 sweeps: cross-spectrum phases of two microphone pairs {01,02}, {03,12}, {13,23} from (U0, U1) and (U2, U3) [four
//            LDTM.x32 per sweep], packed as G_a + i G_b with the Hermitian mirror bins fetched by one shuffle per
//            component, ONE inverse complex FFT whose real / imaginary parts are the two correlations: in-lane DFT-32 over
//            k_hi, twiddle, transpose, second pass pruned to the two outputs that hold the lags [0, 31] and [-32, -1];
//            128-byte coalesced stores straight from registers
//   mel    : lane = (frame, channel) over the four power planes with the baked filterbank; rows staged and copied out
//            by the owning warp after the group's next barrier.
// The two real channels of a packed FFT share a rounding-noise floor, and GCC-PHAT keeps only PHASES — a quiet microphone
// next to a loud one would lose its phase first — so every frame equalises each pair with an exact power of two (max
// |windowed sample| per channel, FMNMX3 + CREDUX); phases do not see the factor, the mel rows are un-scaled by it.
// n_fft 1024 (R1 = 32) and 960 (R1 = 30, the reference's default config.py:85): for 960 the inverse runs over 30 lanes
// with its own twiddle table W_960^(k_lo t_lo).
// Round 2 first kept (U0, U1) in shared memory (8.3 KB per warp) and (U2, U3) in 68 registers per thread: 25 KB and 255
// registers per warp allowed 8 warps per SM (9.4 ms per 256 x 60 s).  With both in tensor memory a warp needs 17 KB and 168
// registers: 12 warps, 8.1 ms; the window / twiddle rows follow them there for n_fft 1024 (7.9 ms; 960 would need 544 of
// the 512 columns).  -DSELD_MIC_TMEM=0 builds the shared-memory form.
// The packed two-instruction complex multiply of dft_inreg.cuh (register pairs + uniform-register constants) measured
// SLOWER here than the four scalar instructions — 3-7 % at 8 warps x 255 registers, 20 % at 12 warps x 168 (tools/mic_ab.py,
// -DSELD_MIC_PACKED_CMUL=1) — so this kernel keeps the scalar form.
#ifndef SELD_MIC_PACKED_CMUL
#define SELD_SCALAR_CMUL
#endif
#include "features_fast.cuh"

namespace seld {

// 1: the unit phasors a lane parks for the inverse sweeps — (U0, U1) of pair a, (U2, U3) of pair b: 2 x 64 words per lane,
//    written and read by the same lane — live in tensor memory (tmem_store.cuh, x32 accesses) instead of 8.3 KB of shared
//    memory and 68 registers per warp; that is what lets the kernel run 12 warps x 168 registers instead of 8 x 255.
#ifndef SELD_MIC_TMEM
#define SELD_MIC_TMEM 1
#endif
constexpr bool kMicTm = SELD_MIC_TMEM != 0;
// 1: (n_fft 1024 only: 960 needs a second twiddle row and 544 > 512 columns) window and twiddle rows in tensor memory too
#ifndef SELD_MIC_TMEM_TABLES
#define SELD_MIC_TMEM_TABLES 1
#endif
constexpr int kMicWarps = kMicTm ? 12 : 8;  // per-warp shared memory: 17 KB (4 power planes + tile) | 25 KB (+ phasor park)
constexpr int kMicTmCols = 512, kMicTmPitch = 128;  // per warp of a lane quarter: Q [16 bins x 4] at +0, (U2, U3) [16 x 4] at +64

template <int R1>
struct MicLayout {
    using F = WarpFft<R1>;
    using FL = FastLayout<R1>;
    static constexpr int NB = F::NB;
    static constexpr int PITCH = FL::PITCH;                  // power planes: same pitch / bank pattern as the FOA kernel
    static constexpr int Q_WORDS = kMicTm ? 0 : 4 * ((NB + 3) & ~3);  // float4 (U0.re, U0.im, U1.re, U1.im) per bin
    static constexpr int P_OFF = Q_WORDS;                    // 4 power planes P0..P3
    static constexpr int TILE_OFF = P_OFF + 4 * PITCH;
    static constexpr int TP = 34;                            // tile row pitch in float2 (17 x 16 B)
    static constexpr int TILE_WORDS = 2 * 32 * TP;           // forward: R1 rows; inverse: 32 rows (t_lo) x R1 columns
    static constexpr int OUT_OFF = TILE_OFF, OUT_PITCH = 65; // staged log-mel rows overlay the (dead) tile
    static constexpr int REGION = (TILE_OFF + TILE_WORDS + 3) & ~3;
    static __device__ __forceinline__ int out_skew(int f) { return (8 * f - f * REGION) & 31; }
    static_assert(PITCH == NB + 3 && OUT_OFF + 31 + 4 * OUT_PITCH <= REGION && Q_WORDS % 4 == 0 && PITCH % 4 == 0, "layout");
};

// x / |x| (0 for an exactly-zero bin: the clamp keeps rsqrt finite and 0 * finite = 0); keep = 0 for a silent channel
__device__ __forceinline__ float2 mic_unit(float2 x, float p, float keep) {
    const float r = rsqrtf(fmaxf(p, 1e-37f)) * keep;
    return make_float2(x.x * r, x.y * r);
}
// conj(a) * b for unit phasors.  A digitally silent channel has zero phasors, so its pairs give G == 0 here; the
// definition R == 0 -> exp(j*angle(0)) = 1 (a delta at lag 0) is restored at the store, where it costs one add per pair
// instead of a test per bin (by linearity: irfft(1) = delta).
__device__ __forceinline__ float2 mic_phat(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.x, b.y, -(a.y * b.x)));
}
template <int S>
__device__ __forceinline__ void mic_pick(float4 q, float4 s, float2& ga, float2& gb) {
    const float2 x0 = make_float2(q.x, q.y), x1 = make_float2(q.z, q.w);
    const float2 x2 = make_float2(s.x, s.y), x3 = make_float2(s.z, s.w);
    if (S == 0) { ga = mic_phat(x0, x1); gb = mic_phat(x0, x2); }
    if (S == 1) { ga = mic_phat(x0, x3); gb = mic_phat(x1, x2); }
    if (S == 2) { ga = mic_phat(x1, x3); gb = mic_phat(x2, x3); }
}

template <int R1>
__global__ void __launch_bounds__(kMicWarps * 32, 1) features_mic_kernel(PlanDev p, FeatArgs a) {
    using F = WarpFft<R1>;
    using L = MicLayout<R1>;
    constexpr int N = F::N, NB = F::NB, WARPS = kMicWarps, G = WARPS / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);                        // [32][36]: window[lane + 32 j] / 2
    float2* s_tw = reinterpret_cast<float2*>(s_win + 32 * 36);               // [32][34]: W_N^(lane k), k < R1
    float2* s_twi = R1 == 32 ? s_tw : s_tw + 32 * 34;                        // [32][34]: W_N^(lane t), t < 32 (lane < R1)
    float* s_regions = reinterpret_cast<float*>(s_twi + 32 * 34);

    for (int i = threadIdx.x; i < 32 * 36; i += blockDim.x) {
        const int l = i / 36, j = i - l * 36;
        s_win[i] = j < R1 ? p.window[l + 32 * j] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * 34; i += blockDim.x) {
        const int l = i / 34, k = i - l * 34;
        s_tw[i] = k < R1 ? p.twiddle[k * 32 + l] : make_float2(0.f, 0.f);
        if (R1 != 32) s_twi[i] = (l < R1 && k < 32) ? p.twiddle[l * 32 + k] : make_float2(1.f, 0.f);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = warp >> 2, wi = warp & 3;
    constexpr bool TT = kMicTm && SELD_MIC_TMEM_TABLES != 0 && R1 == 32;
    uint32_t tm_base = 0, tm_q = 0, tm_tab = 0;  // tm_q: this warp's 128 columns (Q at +0, (U2, U3) at +64)
    if constexpr (kMicTm) {
        __shared__ uint32_t s_tm_slot;
        if (warp == 0) tmem::alloc<kMicTmCols>(&s_tm_slot);
        tmem::fence_before_sync();
        __syncthreads();
        tmem::fence_after_sync();
        tm_base = s_tm_slot;
        tm_q = tmem::lane_base(tm_base, warp) + kMicTmPitch * group;
        if constexpr (TT) {
            tm_tab = tmem::lane_base(tm_base, warp) + 384;
            if (warp < 4) {
                float r[16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = p.window[lane + 32 * (16 * c + i)];
                    tmem::st16(tm_tab + 16 * c, r);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 t = p.twiddle[(8 * c + i) * 32 + lane];
                        r[2 * i] = t.x;
                        r[2 * i + 1] = t.y;
                    }
                    tmem::st16(tm_tab + 32 + 16 * c, r);
                }
                tmem::wait_st();
            }
            tmem::fence_before_sync();
            __syncthreads();
            tmem::fence_after_sync();
        }
    }
    float* region = s_regions + warp * L::REGION;
    float* gregion = s_regions + (group * 4) * L::REGION;
    float4* Q = reinterpret_cast<float4*>(region);
    float* P = region + L::P_OFF;
    float2* T = reinterpret_cast<float2*>(region + L::TILE_OFF);
    const int src = F::partner_lane(lane);
    const bool active = R1 == 32 || lane < R1;
    const int bar_id = 1 + group;
    const float inv_n = 1.0f / float(N);

    const unsigned n_items = (unsigned)a.n_items, T_out = (unsigned)a.T_out;
    const unsigned n_gitems = (n_items + 3u) >> 2;
    unsigned gidx = blockIdx.x * G + group;
    const unsigned gstride = gridDim.x * G;
    if (kMicTm ? false : gidx >= n_gitems) return;  // whole group leaves together (tensor-memory build: everybody meets at the end)
    if (gidx < n_gitems) {

    const float* audio = reinterpret_cast<const float*>(a.audio);
    const int n_samples = (int)a.n_samples, hop = p.hop;
    const int frames_all = 1 + n_samples / hop;
    auto len_of = [&](unsigned b) -> int { return a.lengths ? (int)min(a.lengths[b], (long long)0x7fffffff) : n_samples; };
    auto make_ctx = [&](unsigned slot) {
        FastCtx c;
        const bool exists = slot < n_items;
        c.item = exists ? slot : 0u;
        c.b = c.item / T_out;
        c.t = c.item - c.b * T_out;
        int frames = frames_all;
        bool too_short = false;
        if (a.lengths) {
            const int len = len_of(c.b);
            too_short = len <= F::HALF;
            if (too_short && exists && c.t == 0 && lane == 0) atomicOr(a.status, 1);
            frames = 1 + len / hop;
        }
        c.flags = (exists ? 1 : 0) | ((exists && (int)c.t < frames && !too_short) ? 2 : 0);
        return c;
    };
    auto request = [&](float2 (&v)[R1], const FastCtx& c, int ch) {
        const float* xa = audio + (long long)c.b * a.clip_stride + (long long)ch * a.chan_stride;
        const long long start = (c.flags & 2) ? (long long)c.t * hop - F::HALF : 0ll;
        fast_load_raw<R1, float>(v, xa, xa + a.chan_stride, start, len_of(c.b), lane);
    };
    auto pad = [&](float* planes, int plane, int w) -> float& { return planes[plane * L::PITCH + NB + w]; };

    int prev_flags = 0;
    long long prev_off = 0;
    auto copy_out = [&]() {  // the previous frame's 4 log-mel rows
        if (prev_flags & 1) {
            const float* stage = region + L::OUT_OFF + L::out_skew(wi) + lane;
            const bool ok = prev_flags & 2;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float x0 = power_to_db(stage[c * L::OUT_PITCH]), x1 = power_to_db(stage[c * L::OUT_PITCH + 32]);
                a.out[prev_off + c * 64 + lane] = ok ? x0 : 0.f;
                a.out[prev_off + c * 64 + 32 + lane] = ok ? x1 : 0.f;
            }
        }
    };

    FastCtx cur = make_ctx(4 * gidx + wi);
    float2 v[R1];
    request(v, cur, 0);

    while (true) {
        const unsigned gnext = gidx + gstride;
        const bool more = gnext < n_gitems;
        const long long row_off = (((long long)cur.b * T_out + cur.t) * a.C_out + a.c_off) * 64;
        const bool row_ok = cur.flags & 2, row_exists = cur.flags & 1;
        float4 U23[kMicTm ? 1 : 17];  // unit phasors (U2, U3) of this lane's bins lane + R1 kh (kh < 16) and of the Nyquist bin (lane 0)
        unsigned silent = 0u;  // bit c: channel c is digitally silent in this frame

#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            float wv[32];
            if constexpr (TT) {
                tmem::ld32_issue(tm_tab, wv);
            } else {
                const float4* wrow = reinterpret_cast<const float4*>(s_win + lane * 36);
#pragma unroll
                for (int i = 0; i < (R1 + 3) / 4; ++i) {
                    const float4 w4 = wrow[i];
                    wv[4 * i] = w4.x, wv[4 * i + 1] = w4.y, wv[4 * i + 2] = w4.z, wv[4 * i + 3] = w4.w;
                }
            }
            if (pr == 0) {
                group_barrier(bar_id);
                copy_out();
            }
            if constexpr (TT) tmem::ld32_wait(wv);
#pragma unroll
            for (int j = 0; j < R1; ++j) v[j] = cscale(v[j], wv[j]);
            // block floating point (see features_fast.cuh, BF kernel)
            float ma[4] = {0.f, 0.f, 0.f, 0.f}, mb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < R1; ++j) {
                ma[j & 3] = fmaxf(ma[j & 3], fabsf(v[j].x));
                mb[j & 3] = fmaxf(mb[j & 3], fabsf(v[j].y));
            }
            const unsigned ua = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(fmaxf(ma[0], ma[1]), fmaxf(ma[2], ma[3]))));
            const unsigned ub = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3]))));
            const bool sil_a = ua == 0u, sil_b = ub == 0u;
            int sh = (int)(ua >> 23) - (int)(ub >> 23);
            sh = (sil_a || sil_b) ? 0 : sh;
            sh = (sh > -kBfMinShift && sh < kBfMinShift) ? 0 : max(-kBfMaxShift, min(kBfMaxShift, sh));
            float inv_a = 1.f, inv_b = 1.f;
            if (sh != 0) {
                const int s = sh > 0 ? sh : -sh;
                const float f = __uint_as_float((unsigned)(127 + s) << 23), fi = __uint_as_float((unsigned)(127 - s) << 23);
                if (sh > 0) {
                    inv_b = fi;
#pragma unroll
                    for (int j = 0; j < R1; ++j) v[j].y *= f;
                } else {
                    inv_a = fi;
#pragma unroll
                    for (int j = 0; j < R1; ++j) v[j].x *= f;
                }
            }
            const float keep_a = sil_a ? 0.f : 1.f, keep_b = sil_b ? 0.f : 1.f;
            silent |= (sil_a ? 1u : 0u) << (2 * pr) | (sil_b ? 2u : 0u) << (2 * pr);
            float2 u[32];
            if constexpr (TT) {
                float tw0[32], tw1[32];
                tmem::ld32_issue(tm_tab + 32, tw0);
                tmem::ld32_issue(tm_tab + 64, tw1);
                Dft<R1, false>::run(v);
                tmem::ld32_wait(tw0);
                tmem::ld32_wait(tw1);
                static_for<R1>([&](auto Kc) {
                    constexpr int k = decltype(Kc)::value;
                    if constexpr (k >= 1 && k < 16) v[k] = cmul(v[k], make_float2(tw0[2 * k], tw0[2 * k + 1]));
                    if constexpr (k >= 16) v[k] = cmul(v[k], make_float2(tw1[2 * (k - 16)], tw1[2 * (k - 16) + 1]));
                });
            } else {
                Dft<R1, false>::run(v);
                const float4* trow = reinterpret_cast<const float4*>(s_tw + lane * 34);
                static_for<(R1 + 1) / 2>([&](auto Kq) {
                    constexpr int k0 = 2 * decltype(Kq)::value;
                    const float4 t4 = trow[k0 / 2];
                    if constexpr (k0 >= 1) v[k0] = cmul(v[k0], make_float2(t4.x, t4.y));
                    if constexpr (k0 + 1 < R1) v[k0 + 1] = cmul(v[k0 + 1], make_float2(t4.z, t4.w));
                });
            }
            __syncwarp();
            static_for<R1>([&](auto K) {
                constexpr int k = decltype(K)::value;
                T[k * L::TP + lane] = v[k];
            });
            __syncwarp();
            {
                const float4* urow = reinterpret_cast<const float4*>(T + (active ? lane : 0) * L::TP);
                static_for<16>([&](auto Nq) {
                    constexpr int n = 2 * decltype(Nq)::value;
                    const float4 t4 = urow[n / 2];
                    u[n] = make_float2(t4.x, t4.y);
                    u[n + 1] = make_float2(t4.z, t4.w);
                });
                if (!active) static_for<32>([&](auto Nn) { u[decltype(Nn)::value] = make_float2(0.f, 0.f); });
            }
            __syncwarp();
            if (pr == 0) request(v, cur, 2);  // pair b of this frame; the next frame's pair a is requested in the last sweep
            F::pass2(u);

            auto split = [&](auto KH, float2& xa, float2& xb) {  // X_a = (s.x, d.y), X_b = (s.y, -d.x)
                constexpr int kh = decltype(KH)::value;
                const float2 z = u[kh], m = u[31 - kh];
                float2 q;
                q.x = __shfl_sync(0xffffffffu, m.x, src);
                q.y = __shfl_sync(0xffffffffu, m.y, src);
                const float2 own = u[(32 - kh) & 31];
                q.x = lane == 0 ? own.x : q.x;
                q.y = lane == 0 ? own.y : q.y;
                const float2 sS = cadd(z, q), dD = csub(z, q);
                xa = make_float2(sS.x, dD.y);
                xb = make_float2(sS.y, -dD.x);
            };
            // per bin: scaled powers -> planes (2 pr, 2 pr + 1); unit phasors -> park (pair a) / registers (pair b)
            float stg[32];  // tensor-memory build: unit phasors of eight bins, stored with one STTM.x32
            auto emit = [&](auto Slot, int k, float2 xa, float2 xb) {
                constexpr int slot = decltype(Slot)::value;
                const float pa = fmaf(xa.x, xa.x, xa.y * xa.y), pb = fmaf(xb.x, xb.x, xb.y * xb.y);
                if (active) {  // (tensor-memory build: lanes >= R1 come here with zeros for the warp-wide store below)
                    P[(2 * pr) * L::PITCH + k] = pa;
                    P[(2 * pr + 1) * L::PITCH + k] = pb;
                }
                const float2 ua2 = mic_unit(xa, pa, keep_a), ub2 = mic_unit(xb, pb, keep_b);
                const float4 uu = make_float4(ua2.x, ua2.y, ub2.x, ub2.y);
                if constexpr (kMicTm) {
                    if constexpr (slot < 16) {
                        stg[4 * (slot & 7)] = uu.x, stg[4 * (slot & 7) + 1] = uu.y, stg[4 * (slot & 7) + 2] = uu.z, stg[4 * (slot & 7) + 3] = uu.w;
                    } else {  // the Nyquist bin (lane 0 only): two pad words of the pair's two planes, same thread reads them back
                        pad(P, 2 * pr, 1) = uu.x, pad(P, 2 * pr, 2) = uu.y, pad(P, 2 * pr + 1, 1) = uu.z, pad(P, 2 * pr + 1, 2) = uu.w;
                    }
                } else {
                    if (pr == 0) Q[k] = uu;
                    else U23[slot] = uu;
                }
            };
            static_for<16>([&](auto KH) {
                constexpr int kh = decltype(KH)::value;
                float2 xa, xb;
                split(KH, xa, xb);
                if constexpr (kMicTm) {
                    // (lanes >= R1 of the 960 configuration transform zeros: their phasors are exact zeros)
                    emit(KH, active ? lane + R1 * kh : 0, active ? xa : make_float2(0.f, 0.f), active ? xb : make_float2(0.f, 0.f));
                    if constexpr ((kh & 7) == 7) tmem::st32(tm_q + 64 * pr + 32 * (kh >> 3), stg);
                } else {
                    if (active) emit(KH, lane + R1 * kh, xa, xb);
                    else if (pr == 1) U23[kh] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            });
            {
                const float2 z = cadd(u[16], u[16]);  // Nyquist bin: its own mirror -> X_a = (2 re, 0), X_b = (2 im, 0)
                if (lane == 0) emit(std::integral_constant<int, 16>{}, NB - 1, make_float2(z.x, 0.f), make_float2(z.y, 0.f));
                else if (!kMicTm && pr == 1) U23[kMicTm ? 0 : 16] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (lane == 0) {  // factors of the two mel rows: block-floating un-scaling; 0 for a silent channel, whose packed
                              // spectrum is the partner's rounding noise (-> exactly -100 dB, like the reference's own FFT)
                pad(P, 2 * pr, 0) = sil_a ? 0.f : inv_a * inv_a;
                pad(P, 2 * pr + 1, 0) = sil_b ? 0.f : inv_b * inv_b;
            }
        }

        // ---- three inverse transforms, two microphone pairs each ----
        // (a run-time loop: only the pair selection is specialised per sweep, the transform itself exists once — the
        //  fully unrolled kernel was 168 KB of SASS and stalled on instruction fetch as soon as the audio streamed
        //  through L2 evicted its code)
        float* out_row = a.out + row_off + 4 * 64;
        if constexpr (kMicTm) tmem::wait_st();
#pragma unroll 1
        for (int S = 0; S < 3; ++S) {
            float2 w[32];
            auto build = [&](auto Sc) {
                constexpr int SS = decltype(Sc)::value;
                if constexpr (kMicTm) {
                    // phasors come back from tensor memory eight bins at a time (2 x LDTM.x32); the Hermitian mirror of bin
                    // (lane, kh) is register 31 - kh of lane R1 - lane — for lane 0 its own register 32 - kh, patched one step later
                    static_for<2>([&](auto Hc) {
                        constexpr int h = decltype(Hc)::value;
                        float q[32], uq[32];
                        tmem::ld32(tm_q + 32 * h, q);
                        tmem::ld32(tm_q + 64 + 32 * h, uq);
                        static_for<8>([&](auto Ic) {
                            constexpr int i = decltype(Ic)::value, kh = 8 * h + i;
                            float2 ga, gb;
                            mic_pick<SS>(make_float4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]),
                                         make_float4(uq[4 * i], uq[4 * i + 1], uq[4 * i + 2], uq[4 * i + 3]), ga, gb);
                            float2 g = make_float2(ga.x - gb.y, ga.y + gb.x);            // G_a + i G_b
                            const float2 mir = make_float2(ga.x + gb.y, gb.x - ga.y);   // conj(G_a) + i conj(G_b) = bin N-k
                            if (kh == 0) {  // lane 0 holds DC there: imaginary parts dropped like irfft
                                g.x = lane == 0 ? ga.x : g.x;
                                g.y = lane == 0 ? gb.x : g.y;
                            }
                            w[kh] = g;
                            w[31 - kh] = make_float2(__shfl_sync(0xffffffffu, mir.x, src), __shfl_sync(0xffffffffu, mir.y, src));
                            if constexpr (kh >= 1) {
                                w[32 - kh].x = lane == 0 ? mir.x : w[32 - kh].x;
                                w[32 - kh].y = lane == 0 ? mir.y : w[32 - kh].y;
                            }
                        });
                    });
                    if (lane == 0) {  // register 16 of lane 0: the Nyquist bin, real parts only (irfft ignores its imaginary part)
                        float2 ga, gb;
                        mic_pick<SS>(make_float4(pad(P, 0, 1), pad(P, 0, 2), pad(P, 1, 1), pad(P, 1, 2)),
                                     make_float4(pad(P, 2, 1), pad(P, 2, 2), pad(P, 3, 1), pad(P, 3, 2)), ga, gb);
                        w[16] = make_float2(ga.x, gb.x);
                    }
                    return;
                }
                float2 nyq = make_float2(0.f, 0.f);
                if (lane == 0) {
                    float2 ga, gb;
                    mic_pick<SS>(Q[NB - 1], U23[kMicTm ? 0 : 16], ga, gb);
                    nyq = make_float2(ga.x, gb.x);  // irfft ignores the imaginary part of the Nyquist bin
                }
                float2 mir[16];
                static_for<16>([&](auto KH) {
                    constexpr int kh = decltype(KH)::value;
                    const float4 q = Q[active ? lane + R1 * kh : 0];
                    float2 ga, gb;
                    mic_pick<SS>(q, U23[kMicTm ? 0 : kh], ga, gb);
                    float2 g = make_float2(ga.x - gb.y, ga.y + gb.x);     // G_a + i G_b
                    mir[kh] = make_float2(ga.x + gb.y, gb.x - ga.y);      // conj(G_a) + i conj(G_b) = bin N-k
                    if (kh == 0) {  // lane 0 holds DC there: imaginary parts dropped like irfft
                        g.x = lane == 0 ? ga.x : g.x;
                        g.y = lane == 0 ? gb.x : g.y;
                    }
                    w[kh] = g;
                });
                static_for<16>([&](auto KH) {
                    constexpr int kh = decltype(KH)::value;
                    float2 r;
                    r.x = __shfl_sync(0xffffffffu, mir[kh].x, src);
                    r.y = __shfl_sync(0xffffffffu, mir[kh].y, src);
                    const float2 own = kh == 15 ? nyq : mir[(kh + 1) & 15];  // lane 0: register 31-kh is bin R1*(kh+1) mirrored
                    r.x = lane == 0 ? own.x : r.x;
                    r.y = lane == 0 ? own.y : r.y;
                    w[31 - kh] = r;
                });
            };
            if (S == 0) build(std::integral_constant<int, 0>{});
            else if (S == 1) build(std::integral_constant<int, 1>{});
            else build(std::integral_constant<int, 2>{});
            // first pass (registers): inverse DFT-32 over k_hi, then the twiddle conj(W_N^(k_lo t_lo))
            if constexpr (TT) {
                float tw0[32], tw1[32];
                tmem::ld32_issue(tm_tab + 32, tw0);
                tmem::ld32_issue(tm_tab + 64, tw1);
                Dft<32, true>::run(w);
                tmem::ld32_wait(tw0);
                tmem::ld32_wait(tw1);
                static_for<32>([&](auto Kc) {
                    constexpr int k = decltype(Kc)::value;
                    if constexpr (k >= 1 && k < 16) w[k] = cmul_conj(w[k], make_float2(tw0[2 * k], tw0[2 * k + 1]));
                    if constexpr (k >= 16) w[k] = cmul_conj(w[k], make_float2(tw1[2 * (k - 16)], tw1[2 * (k - 16) + 1]));
                });
            } else {
                Dft<32, true>::run(w);
                const float4* trow = reinterpret_cast<const float4*>(s_twi + lane * 34);
                static_for<16>([&](auto Kq) {
                    constexpr int k0 = 2 * decltype(Kq)::value;
                    const float4 t4 = trow[k0 / 2];
                    if constexpr (k0 >= 1) w[k0] = cmul_conj(w[k0], make_float2(t4.x, t4.y));
                    w[k0 + 1] = cmul_conj(w[k0 + 1], make_float2(t4.z, t4.w));
                });
            }
            __syncwarp();
            if (active) static_for<32>([&](auto K) {  // row t_lo belongs to reader lane t_lo; column = this lane (k_lo)
                constexpr int k = decltype(K)::value;
                T[k * L::TP + lane] = w[k];
            });
            __syncwarp();
            if (S == 2 && more) request(v, make_ctx(4 * gnext + wi), 0);  // the group's next frame, pair a
            // second pass, pruned: t_hi = 0 -> lags t_lo; t_hi = R1 - 1 -> lags t_lo - 32
            float2 pos = make_float2(0.f, 0.f), neg = make_float2(0.f, 0.f);
            {
                const float4* urow = reinterpret_cast<const float4*>(T + lane * L::TP);
                float2 pos2 = make_float2(0.f, 0.f), neg2 = make_float2(0.f, 0.f);
                static_for<R1 / 2>([&](auto Nq) {
                    constexpr int n = 2 * decltype(Nq)::value;
                    const float4 t4 = urow[n / 2];
                    const float2 x0 = make_float2(t4.x, t4.y), x1 = make_float2(t4.z, t4.w);
                    pos = cadd(pos, x0);
                    pos2 = cadd(pos2, x1);
                    neg = cadd(neg, mul_w<n, R1, false>(x0));       // exp(+2 pi i (R1-1) n / R1) = W_R1^n
                    neg2 = cadd(neg2, mul_w<n + 1, R1, false>(x1));
                });
                pos = cadd(pos, pos2);
                neg = cadd(neg, neg2);
            }
            __syncwarp();
            if (row_exists) {  // real part = first pair of the couple, imaginary part = second; lags [-32,-1] then [0,31]
                // pairs of the sweeps: {01, 02}, {03, 12}, {13, 23}; a pair with a silent channel is a delta at lag 0
                const unsigned ma_ = S == 0 ? 0x3u : S == 1 ? 0x9u : 0xau, mb_ = S == 0 ? 0x5u : S == 1 ? 0x6u : 0xcu;
                const float da = (lane == 0 && (silent & ma_)) ? 1.f : 0.f, db = (lane == 0 && (silent & mb_)) ? 1.f : 0.f;
                float* oa = out_row + (2 * S) * 64;
                float* ob = oa + 64;
                oa[lane] = row_ok ? neg.x * inv_n : 0.f;
                oa[32 + lane] = row_ok ? fmaf(pos.x, inv_n, da) : 0.f;
                ob[lane] = row_ok ? neg.y * inv_n : 0.f;
                ob[32 + lane] = row_ok ? fmaf(pos.y, inv_n, db) : 0.f;
            }
        }

        prev_flags = cur.flags;
        prev_off = row_off;
        group_barrier(bar_id);  // the four frames of the group have their power planes

        // ---- mel phase: lane = (frame, channel), 4 power planes; this warp owns filter chunk wi ----
        {
            const int f = lane >> 3, c = lane & 7;
            if (c < 4) {
                float* rf = gregion + f * L::REGION;
                const float* vp = rf + L::P_OFF + c * L::PITCH;
                float* orow = rf + L::OUT_OFF + L::out_skew(f) + c * L::OUT_PITCH;
                const float fac = vp[NB];
                switch (wi) {
                    case 0: fast_mel_chunk<N, 0, true>(vp, orow, fac); break;
                    case 1: fast_mel_chunk<N, 1, true>(vp, orow, fac); break;
                    case 2: fast_mel_chunk<N, 2, true>(vp, orow, fac); break;
                    default: fast_mel_chunk<N, 3, true>(vp, orow, fac); break;
                }
            }
        }
        if (!more) break;
        gidx = gnext;
        cur = make_ctx(4 * gidx + wi);
    }
    group_barrier(bar_id);
    copy_out();
    }  // gidx < n_gitems
    if constexpr (kMicTm) {
        tmem::fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem::dealloc<kMicTmCols>(tm_base);
    }
}

template <int R1>
static constexpr size_t mic_smem_bytes() {
    return sizeof(float) * 32 * 36 + sizeof(float2) * 32 * 34 * (R1 == 32 ? 1 : 2) +
           sizeof(float) * (size_t)kMicWarps * MicLayout<R1>::REGION;
}

int configure_gcc_kernels(const seld_plan* plan) {
    static_assert(mic_smem_bytes<32>() <= (size_t)kMaxSmemOptin && mic_smem_bytes<30>() <= (size_t)kMaxSmemOptin, "smem budget");
    if (!plan->v3_ok) return SELD_OK;  // the fused MIC kernel uses the baked 64-mel filterbank
    if (plan->dev.r1 == 32)
        SELD_CUDA_TRY(cudaFuncSetAttribute(features_mic_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mic_smem_bytes<32>()));
    else
        SELD_CUDA_TRY(cudaFuncSetAttribute(features_mic_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mic_smem_bytes<30>()));
    return SELD_OK;
}

// 4 log-mel + 6 GCC-PHAT channels of a 4-microphone batch in one launch (a.c_off = first of the 10 output channels)
int launch_gcc(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    if (!plan->v3_ok) {
        set_error("seld_features: the MIC (GCC-PHAT) mode needs the reference's 64-mel HTK filterbank");
        return SELD_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(a.out) & 15) != 0) {
        set_error("seld_features: the MIC (GCC-PHAT) mode needs a 16-byte aligned output");
        return SELD_ERR_BAD_ARG;
    }
    FeatArgs g = a;
    g.G = 1;
    g.n_items = (long long)a.B * a.T_out;
    const long long n_gitems = (g.n_items + 3) / 4;
    long long ctas = (n_gitems + kMicWarps / 4 - 1) / (kMicWarps / 4);
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    if (plan->dev.r1 == 32)
        features_mic_kernel<32><<<(unsigned)ctas, kMicWarps * 32, mic_smem_bytes<32>(), stream>>>(plan->dev, g);
    else
        features_mic_kernel<30><<<(unsigned)ctas, kMicWarps * 32, mic_smem_bytes<30>(), stream>>>(plan->dev, g);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

}  // namespace seld
