// v3 of the fused feature kernel (K1+K2): framing + Hann + real FFT (two channels per complex FFT) + power +
// FOA intensity vectors + mel projection + 10*log10 for the reference's own configurations
// (4 channels, 64 HTK mels, n_fft 1024 or 960 — reference dataset.py:27-58 with config.py:85-87).
// Anything else (other channel counts, other filterbanks, spectrum dumps) runs on the generic kernel in
// features.cu.
//
// What bounds this path on B200 is not HBM but the SM's shared-memory datapath (128 B/clk) and the FP32
// pipe; v3 is organised around shared-memory wavefronts per frame:
//   * warps work in groups of four.  Phase A: every warp transforms one frame (all four channels, two
//     packed complex FFTs) exactly like v2, but leaves its per-bin features as seven PLANES
//     V[c][k] (P0 P1 P2 P3 I1/E I2/E I3/E) in its own shared-memory region.  What the second channel pair
//     needs from the first (X0, |X1|^2, Re(conj(X0) X1)) is parked in planes 0..3 (same lane, same bin: in
//     place), the transpose tile lives behind them where planes 4..6 and the staged output row go later.
//   * Phase B (after a 128-thread named barrier): the four warps project the group's four frames onto the
//     mel filters.  Lane = (frame, channel), so all lanes walk the SAME bins: the filterbank is baked into
//     the instruction stream (mel_baked.h: weights are FFMA immediates, filter boundaries are straight-line
//     code), every plane word is read exactly once (no gather tables, no padding, no bank conflicts) and
//     each warp owns a quarter of the filters (balanced by instruction count).  The raw mel energies are
//     staged behind the planes; after the group's next barrier the owning warp applies 10 log10 and writes
//     the frame's 1792-byte row with coalesced 128-byte stores.
//   * arithmetic is packed where the data come in pairs (add/sub/mul/fma.rn.f32x2 -> FADD2/FMUL2/FFMA2 on
//     sm_100a): butterflies, window, channel split, power pairs.
#include <cstdlib>

#include "mel_baked.h"
#include "seld_common.h"
#include "warp_fft.cuh"

namespace seld {

template <int R1, bool REGSTASH>
struct V3Layout {
    using F = WarpFft<R1>;
    static constexpr int NB = F::NB;                         // 513 / 481
    // plane pitch = 4 (mod 8) words: in the mel phase lane (f, c) reads the float4 at word f*REGION + c*PITCH + k;
    // the 8 lanes of a quarter-warp (one f, c = 0..7) then hit 8 different 16-byte bank groups.
    static constexpr int PITCH = NB + ((4 - NB % 8) + 8) % 8;  // 516 / 484
    static constexpr int TILE_OFF = 4 * PITCH;               // the transpose tile sits behind the four parking planes
    static constexpr int TP = 34;                            // tile row pitch in float2: 272 B = 17 x 16 B, see below
    static constexpr int TILE_WORDS = 2 * R1 * TP;           // R1 rows (one per reader lane) of 34 float2
    // constant tables, one row per LANE so that a lane fetches its values with 128-bit loads; row pitches of
    // 9 x 16 B and 17 x 16 B put the 8 lanes of a quarter-warp on 8 different 16-byte bank groups
    static constexpr int WIN_PITCH = 36;                     // floats:  window[lane + 32 j] / 2 at [lane][j]
    static constexpr int TW_PITCH = 34;                      // float2s: W_N^(lane k) at [lane][k]
    static constexpr int OUT_END = 7 * PITCH + 31 + 7 * 65;
    static constexpr int REGION = ((TILE_OFF + TILE_WORDS > OUT_END ? TILE_OFF + TILE_WORDS : OUT_END) + 3) & ~3;
    // behind the seven planes: the frame's finished output row (7 x 64, channel pitch 65), staged by the mel phase
    // and copied out as full 128-byte lines by the owning warp.  Frame slot f starts at OUT_OFF + out_skew(f)
    // so that lane (f, c) lands on bank 8 f + c + m.
    static constexpr int OUT_OFF = 7 * PITCH, OUT_PITCH = 65;
    static __device__ __forceinline__ int out_skew(int f) { return (8 * f - f * REGION) & 31; }
    static_assert(PITCH % 8 == 4 && REGION % 4 == 0 && REGION >= OUT_OFF + 31 + 7 * OUT_PITCH, "region layout");
};


// (s.x^2 + d.y^2, s.y^2 + d.x^2): the powers of the two real channels of a packed pair, two packed instructions
// (ptxas folds the swap of d into an operand swizzle)
__device__ __forceinline__ float2 pow_pair(float2 s, float2 d) {
    float2 r;
    asm("{.reg .b64 rs, rd, t; mov.b64 rs, {%2, %3}; mov.b64 rd, {%5, %4}; mul.rn.f32x2 t, rs, rs; fma.rn.f32x2 t, rd, rd, t; "
        "mov.b64 {%0, %1}, t;}"
        : "=f"(r.x), "=f"(r.y) : "f"(s.x), "f"(s.y), "f"(d.x), "f"(d.y));
    return r;
}

__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// edge frames (reflect padding, torch.stft center=True): rare, kept out of line
template <int R1>
__device__ __noinline__ void v3_load_edge(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                          long long len, int lane) {
    using F = WarpFft<R1>;
#pragma unroll 4
    for (int j = 0; j < R1; ++j) {
        const long long idx = F::reflect(start + lane + 32 * j, len);
        v[j] = make_float2(__ldg(xa + idx), __ldg(xb + idx));
    }
}

template <int R1>
__device__ __forceinline__ void v3_load_raw(float2 (&v)[R1], const float* xa, const float* xb, long long start,
                                            long long len, int lane) {
    using F = WarpFft<R1>;
    if ((start >= 0) && (start + F::N <= len)) {
        const float* pa = xa + start + lane;
        const float* pb = xb + start + lane;
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = make_float2(__ldg(pa + 32 * j), __ldg(pb + 32 * j));
    } else {  // via a scratch array so that v itself never has its address taken (it must stay in registers)
        float2 tmp[R1];
        v3_load_edge<R1>(tmp, xa, xb, start, len, lane);
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = tmp[j];
    }
}

// ask L2 for the 128-byte lines of a channel pair's frame (lane l: line l of each channel); no registers held
template <int R1>
__device__ __forceinline__ void v3_prefetch_l2(const float* xa, const float* xb, long long start, long long len, int lane) {
    long long o = start + 32 * lane;
    o = o < 0 ? 0 : o;
    if (lane < R1 && o < len) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xa + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + o));
    }
}

// ---- mel projection of one filter chunk for lane (frame, channel) -------------------------------
template <int NFFT, int CH>
__device__ __forceinline__ void v3_mel_chunk(const float* __restrict__ vp, float* __restrict__ orow) {
    using MB = MelBaked<NFFT>;
    constexpr int M_LO = MB::chunk_m[CH], M_HI = MB::chunk_m[CH + 1];
    constexpr int K0 = MB::chunk_k0[CH], K1 = MB::chunk_k1[CH];
    // filters of even / odd index (filters two apart never overlap) x even / odd bins (two FMA chains per filter)
    float acc00 = 0.f, acc01 = 0.f, acc10 = 0.f, acc11 = 0.f;
    auto contribute = [&](auto Mc, auto Kc, auto Sel, float v) {
        constexpr int m = decltype(Mc)::value, k = decltype(Kc)::value;
        constexpr float w = decltype(Sel)::value ? MB::w1[k] : MB::w0[k];
        if constexpr (m >= M_LO && m < M_HI) {
            float& acc = (m & 1) ? ((k & 1) ? acc11 : acc10) : ((k & 1) ? acc01 : acc00);
            if constexpr (k == MB::first[m] || k == MB::first[m] + 1) acc = w * v;
            else acc = fmaf(w, v, acc);
            if constexpr (k == MB::last[m]) {
                float sum;
                if constexpr (MB::last[m] == MB::first[m]) sum = acc;
                else sum = ((m & 1) ? acc10 : acc00) + ((m & 1) ? acc11 : acc01);
                orow[m] = sum;  // raw mel energy / mel-binned IV; the copy-out applies 10 log10 and row validity
            }
        }
    };
    constexpr int G0 = K0 & ~3, NG = ((K1 + 3) & ~3) - G0;
    static_for<NG / 4>([&](auto Gi) {
        constexpr int g = G0 + 4 * decltype(Gi)::value;
        const float4 v4 = *reinterpret_cast<const float4*>(vp + g);
        static_for<4>([&](auto I) {
            constexpr int k = g + decltype(I)::value;
            if constexpr (k >= K0 && k < K1) {
                constexpr int ma = MB::m0[k], mb = MB::m1[k];
                constexpr bool use_a = ma >= M_LO && ma < M_HI, use_b = mb >= M_LO && mb < M_HI;
                const float v = decltype(I)::value == 0 ? v4.x : decltype(I)::value == 1 ? v4.y : decltype(I)::value == 2 ? v4.z : v4.w;
                if constexpr (use_a)
                    contribute(std::integral_constant<int, (ma < 0 ? 0 : ma)>{}, std::integral_constant<int, k>{},
                               std::integral_constant<int, 0>{}, v);
                if constexpr (use_b)
                    contribute(std::integral_constant<int, (mb < 0 ? 0 : mb)>{}, std::integral_constant<int, k>{},
                               std::integral_constant<int, 1>{}, v);
            }
        });
    });
}

struct V3Ctx {        // one frame: clip b, frame t
    unsigned b, t;
    long long len;    // samples of the clip
    int valid, exists;
};

template <int R1, bool IV, int WARPS, bool REGSTASH>
__global__ void __launch_bounds__(WARPS * 32, 1) features_v3_kernel(PlanDev p, FeatArgs a) {
    using F = WarpFft<R1>;
    using L = V3Layout<R1, REGSTASH>;
    constexpr int N = F::N, NB = F::NB;
    constexpr int NCH = IV ? 7 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(s_win + 32 * L::WIN_PITCH);
    float* s_regions = reinterpret_cast<float*>(s_tw + 32 * L::TW_PITCH);

    for (int i = threadIdx.x; i < 32 * L::WIN_PITCH; i += blockDim.x) {
        const int l = i / L::WIN_PITCH, j = i - l * L::WIN_PITCH;
        s_win[i] = j < R1 ? p.window[l + 32 * j] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * L::TW_PITCH; i += blockDim.x) {
        const int l = i / L::TW_PITCH, k = i - l * L::TW_PITCH;
        s_tw[i] = k < R1 ? p.twiddle[k * 32 + l] : make_float2(0.f, 0.f);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int group = warp >> 2, wi = warp & 3;
    float* region = s_regions + warp * L::REGION;
    float* gregion = s_regions + (group * 4) * L::REGION;
    float2* T = reinterpret_cast<float2*>(region + L::TILE_OFF);
    const int src = F::partner_lane(lane);
    const bool active = R1 == 32 || lane < R1;
    const int bar_id = 1 + group;
    constexpr int G = WARPS / 4;

    const long long n_items = a.n_items, T_out = a.T_out;
    const long long n_gitems = (n_items + 3) >> 2;
    long long gidx = (long long)blockIdx.x * G + group;
    const long long gstride = (long long)gridDim.x * G;
    if (gidx >= n_gitems) return;  // whole group leaves together

    const unsigned T_out_u = (unsigned)T_out;
    const long long frames_all = 1 + a.n_samples / p.hop;
    auto make_ctx = [&](long long item) {
        V3Ctx c;
        c.exists = item < n_items;
        const unsigned it = c.exists ? (unsigned)item : 0u;
        c.b = it / T_out_u;
        c.t = it - c.b * T_out_u;
        c.len = a.lengths ? a.lengths[c.b] : a.n_samples;
        const long long frames = a.lengths ? 1 + c.len / p.hop : frames_all;
        c.valid = c.exists && (long long)c.t < frames;
        return c;
    };
    auto chan0 = [&](const V3Ctx& c) { return a.audio + (long long)c.b * a.clip_stride; };
    auto frame_start = [&](const V3Ctx& c) { return c.valid ? (long long)c.t * p.hop - F::HALF : 0ll; };

    // copy this warp's staged output row (7 x 64 floats) to global memory, 128 bytes per store
    // (10 log10 of the power channels happens here; rows beyond the clip's last frame are written as 0)
    float* prev_row = nullptr;
    bool prev_valid = false;
    const float* stage = region + L::OUT_OFF + L::out_skew(wi) + lane;
    auto copy_out = [&]() {
        if (prev_row) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float x0 = stage[c * L::OUT_PITCH], x1 = stage[c * L::OUT_PITCH + 32];
                if (c < 4) {
                    x0 = power_to_db(x0);
                    x1 = power_to_db(x1);
                }
                prev_row[c * 64 + lane] = prev_valid ? x0 : 0.f;
                prev_row[c * 64 + 32 + lane] = prev_valid ? x1 : 0.f;
            }
        }
    };

    V3Ctx cur = make_ctx(4 * gidx + wi);
    float2 v[R1];  // raw samples of the channel pair about to be transformed, requested one phase ahead
    v3_load_raw<R1>(v, chan0(cur), chan0(cur) + a.chan_stride, frame_start(cur), cur.len, lane);

    while (true) {
        const long long gnext = gidx + gstride;
        const bool more = gnext < n_gitems;
        V3Ctx nxt = cur;
        if (more) nxt = make_ctx(4 * gnext + wi);

        float4 st[REGSTASH ? 17 : 1];  // REGSTASH: (X0.re, X0.im, P1, I1) of the lane's 17 bins while pair b is transformed
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            // ---- window + pass 1 + twiddle (registers only) ----
            // this lane's window row (a constant table: fetched before the barrier so that the loads are in flight)
            float4 wreg[(R1 + 3) / 4];
            {
                const float4* wrow = reinterpret_cast<const float4*>(s_win + lane * L::WIN_PITCH);
#pragma unroll
                for (int i = 0; i < (R1 + 3) / 4; ++i) wreg[i] = wrow[i];
            }
            // the other warps of the group have finished reading this region (previous mel phase); the previous
            // frame's row is complete (all four filter chunks staged): its copy-out overlaps the first pass below
            if (pr == 0) {
                group_barrier(bar_id);
                copy_out();
            }
            // Digitally silent channel?  Cheap necessary test first (every lane's first sample is +-0); the OR over all
            // raw samples only runs when it passes (rare, warp-uniform).
            bool sil_a = !__any_sync(0xffffffffu, (__float_as_uint(v[0].x) << 1) != 0u);
            bool sil_b = !__any_sync(0xffffffffu, (__float_as_uint(v[0].y) << 1) != 0u);
            if (sil_a || sil_b) {
                unsigned bits_a = 0u, bits_b = 0u;
#pragma unroll
                for (int j = 0; j < R1; ++j) {
                    bits_a |= __float_as_uint(v[j].x);
                    bits_b |= __float_as_uint(v[j].y);
                }
                sil_a = !__any_sync(0xffffffffu, (bits_a << 1) != 0u);
                sil_b = !__any_sync(0xffffffffu, (bits_b << 1) != 0u);
            }
            {
                static_for<(R1 + 3) / 4>([&](auto Jq) {
                    constexpr int j0 = 4 * decltype(Jq)::value;
                    const float4 w4 = wreg[j0 / 4];
                    static_for<4>([&](auto Ji) {
                        constexpr int j = j0 + decltype(Ji)::value;
                        if constexpr (j < R1) {
                            const float w = decltype(Ji)::value == 0 ? w4.x : decltype(Ji)::value == 1 ? w4.y : decltype(Ji)::value == 2 ? w4.z : w4.w;
                            v[j] = cscale(v[j], w);
                        }
                    });
                });
            }
            Dft<R1, false>::run(v);
            {
                const float4* trow = reinterpret_cast<const float4*>(s_tw + lane * L::TW_PITCH);
                static_for<(R1 + 1) / 2>([&](auto Kq) {
                    constexpr int k0 = 2 * decltype(Kq)::value;
                    const float4 t4 = trow[k0 / 2];
                    if constexpr (k0 >= 1) v[k0] = cmul(v[k0], make_float2(t4.x, t4.y));
                    if constexpr (k0 + 1 < R1) v[k0 + 1] = cmul(v[k0 + 1], make_float2(t4.z, t4.w));
                });
            }
            __syncwarp();
            static_for<R1>([&](auto K) {  // row k belongs to reader lane k; column = this lane
                constexpr int k = decltype(K)::value;
                T[k * L::TP + lane] = v[k];
            });
            __syncwarp();
            float2 u[32];
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
            // v is dead: request the NEXT channel pair — pair b of this frame, or pair a of the group's next frame.
            // The loads are issued here so that they interleave with the arithmetic of pass 2 and land before
            // the next window pass (one code site for both cases).
            if (pr == 0 || more) {
                const V3Ctx& nx = pr == 0 ? cur : nxt;
                const float* xn = chan0(nx) + (pr == 0 ? 2 : 0) * a.chan_stride;
                v3_load_raw<R1>(v, xn, xn + a.chan_stride, frame_start(nx), nx.len, lane);
            }
            F::pass2(u);

            // Channel split in packed form: with z = Z[k], p = conj-mirror partner Z[N-k] as shuffled (p.x, p.y),
            // s = z + p and d = z - p (one FADD2 each) give X_a = (s.x, d.y), X_b = (s.y, -d.x)  [window pre-scaled by 1/2].
            auto split = [&](auto KH, float2& sS, float2& dD) {
                constexpr int kh = decltype(KH)::value;
                const float2 z = u[kh], m = u[31 - kh];
                float2 q;
                q.x = __shfl_sync(0xffffffffu, m.x, src);
                q.y = __shfl_sync(0xffffffffu, m.y, src);
                const float2 own = u[(32 - kh) & 31];  // lane 0 holds its own mirror bins
                q.x = lane == 0 ? own.x : q.x;
                q.y = lane == 0 ? own.y : q.y;
                sS = cadd(z, q);
                dD = csub(z, q);
            };
            if (pr == 0) {
                // ---- park X0 = (s.x, d.y), P1 = |X1|^2 and I1 = Re(conj(X0) X1) in planes 0..3 (same lane reads them back) ----
                auto park = [&](auto Slot, int k, float2 sS, float2 dD) {
                    const float p1 = fmaf(sS.y, sS.y, dD.x * dD.x), i1 = fmaf(sS.x, sS.y, -(dD.x * dD.y));
                    if constexpr (REGSTASH) {
                        st[decltype(Slot)::value] = make_float4(sS.x, dD.y, p1, i1);
                    } else {
                        region[k] = sS.x;
                        region[L::PITCH + k] = dD.y;
                        region[2 * L::PITCH + k] = p1;
                        region[3 * L::PITCH + k] = i1;
                    }
                };
                static_for<16>([&](auto KH) {
                    float2 sS, dD;
                    split(KH, sS, dD);
                    if (REGSTASH || active) park(KH, lane + R1 * decltype(KH)::value, sS, dD);
                });
                if (REGSTASH || lane == 0) park(std::integral_constant<int, REGSTASH ? 16 : 0>{}, NB - 1, cadd(u[16], u[16]), make_float2(0.f, 0.f));
                if constexpr (REGSTASH) {
                    if (sil_a || sil_b) {  // rare: a digitally silent channel must give an exactly-zero spectrum
                        const float ka = sil_a ? 0.f : 1.f, kb = sil_b ? 0.f : 1.f;
#pragma unroll
                        for (int i = 0; i < 17; ++i) st[i] = make_float4(st[i].x * ka, st[i].y * ka, st[i].z * kb, st[i].w * ka * kb);
                    }
                } else if (sil_a || sil_b) {  // rare: a digitally silent channel must give an exactly-zero spectrum
                    const float ka = sil_a ? 0.f : 1.f, kb = sil_b ? 0.f : 1.f;
                    __syncwarp();
                    for (int k = lane; k < NB; k += 32) {
                        region[k] *= ka;
                        region[L::PITCH + k] *= ka;
                        region[2 * L::PITCH + k] *= kb;
                        region[3 * L::PITCH + k] *= ka * kb;
                    }
                    __syncwarp();
                }
            } else {
                // ---- per-bin features -> seven planes (planes 4..6 overlay the dead transpose tile) ----
                auto finish = [&](auto Slot, int k, float2 sS, float2 dD) {
                    float x0r, x0i, p1, i1;
                    if constexpr (REGSTASH) {
                        const float4 t = st[decltype(Slot)::value];
                        x0r = t.x, x0i = t.y, p1 = t.z, i1 = t.w;
                    } else {
                        x0r = region[k], x0i = region[L::PITCH + k];
                        p1 = region[2 * L::PITCH + k], i1 = region[3 * L::PITCH + k];
                    }
                    const float2 p23 = pow_pair(sS, dD);  // (|X2|^2, |X3|^2)
                    const float p0 = fmaf(x0r, x0r, x0i * x0i);
                    region[k] = p0;
                    region[L::PITCH + k] = p1;
                    region[2 * L::PITCH + k] = p23.x;
                    region[3 * L::PITCH + k] = p23.y;
                    if (IV) {
                        const float i2 = fmaf(x0r, sS.x, x0i * dD.y);     // Re(conj(X0) X2), X2 = (s.x, d.y)
                        const float i3 = fmaf(x0r, sS.y, -(x0i * dD.x));  // Re(conj(X0) X3), X3 = (s.y, -d.x)
                        const float e = kEpsIV + p0 + (p1 + p23.x + p23.y) * (1.0f / 3.0f);
                        float inv;  // MUFU.RCP alone (1 ulp): e >= 1e-8 is never denormal; the IV tolerance is 1e-4 relative
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(e));
                        region[4 * L::PITCH + k] = i1 * inv;
                        region[5 * L::PITCH + k] = i2 * inv;
                        region[6 * L::PITCH + k] = i3 * inv;
                    }
                };
                static_for<16>([&](auto KH) {
                    float2 sS, dD;
                    split(KH, sS, dD);
                    if (active) finish(KH, lane + R1 * decltype(KH)::value, sS, dD);
                });
                if (lane == 0) finish(std::integral_constant<int, REGSTASH ? 16 : 0>{}, NB - 1, cadd(u[16], u[16]), make_float2(0.f, 0.f));
                if (sil_a || sil_b) {  // rare: redo the planes with the silent channel (2 = a, 3 = b) at exactly 0
                    __syncwarp();
                    for (int k = lane; k < NB; k += 32) {
                        const float p0 = region[k], p1 = region[L::PITCH + k], p2o = region[2 * L::PITCH + k], p3o = region[3 * L::PITCH + k];
                        const float p2 = sil_a ? 0.f : p2o, p3 = sil_b ? 0.f : p3o;
                        region[2 * L::PITCH + k] = p2;
                        region[3 * L::PITCH + k] = p3;
                        if (IV) {
                            const float e_old = kEpsIV + p0 + (p1 + p2o + p3o) * (1.0f / 3.0f);
                            const float e_new = kEpsIV + p0 + (p1 + p2 + p3) * (1.0f / 3.0f);
                            const float g = e_old / e_new;
                            region[4 * L::PITCH + k] *= g;
                            region[5 * L::PITCH + k] = sil_a ? 0.f : region[5 * L::PITCH + k] * g;
                            region[6 * L::PITCH + k] = sil_b ? 0.f : region[6 * L::PITCH + k] * g;
                        }
                    }
                }
            }
        }
        prev_valid = cur.valid;
        prev_row = cur.exists ? a.out + (((long long)cur.b * T_out + cur.t) * a.C_out + a.c_off) * 64 + 0 : nullptr;
        group_barrier(bar_id);  // the four frames of the group are in their planes

        // ---- mel phase: lane = (frame, channel); this warp owns filter chunk wi ----
        {
            const int f = lane >> 3, c = lane & 7;
            if (c < NCH) {  // (slots of frames that do not exist hold finite dummy planes; their rows are never copied out)
                const float* vp = gregion + f * L::REGION + c * L::PITCH;
                float* orow = gregion + f * L::REGION + L::OUT_OFF + L::out_skew(f) + c * L::OUT_PITCH;
                switch (wi) {
                    case 0: v3_mel_chunk<N, 0>(vp, orow); break;
                    case 1: v3_mel_chunk<N, 1>(vp, orow); break;
                    case 2: v3_mel_chunk<N, 2>(vp, orow); break;
                    default: v3_mel_chunk<N, 3>(vp, orow); break;
                }
            }
        }
        if (!more) break;
        gidx = gnext;
        cur = nxt;
    }
    group_barrier(bar_id);
    copy_out();
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
bool v3_filterbank_matches(int n_fft, const float* fb, int n_mels) {
    return n_fft == 1024 ? fb_matches<1024>(fb, n_mels) : n_fft == 960 ? fb_matches<960>(fb, n_mels) : false;
}

template <int R1, bool IV, int WARPS, bool REGSTASH>
static int launch_v3_cfg(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    using L = V3Layout<R1, REGSTASH>;
    auto kern = features_v3_kernel<R1, IV, WARPS, REGSTASH>;
    const size_t smem = sizeof(float) * (32 * L::WIN_PITCH) + sizeof(float2) * (32 * L::TW_PITCH) +
                        sizeof(float) * (size_t)WARPS * L::REGION;
    SELD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_gitems = (a.n_items + 3) / 4;
    long long ctas = (n_gitems + WARPS / 4 - 1) / (WARPS / 4);
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    kern<<<(unsigned)ctas, WARPS * 32, smem, stream>>>(plan->dev, a);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

// Two resource configurations of the same kernel (SELD_V3_CFG=12s|8r selects one for A/B runs):
//   12s: 12 warps x 168 registers, first pair's spectra parked in shared memory
//   8r :  8 warps x 255 registers, parked in registers (136 fewer shared-memory wavefronts per frame, more L1)
template <int R1, bool IV>
static int launch_v3_one(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    const char* e = getenv("SELD_V3_CFG");  // measured equal at n_fft 1024 (4.45 vs 4.47 ms), 12s faster at 960
    if (e && e[0] == '8') return launch_v3_cfg<R1, IV, 8, true>(plan, a, stream);
    return launch_v3_cfg<R1, IV, 12, false>(plan, a, stream);
}

int launch_features_v3(const seld_plan* plan, bool iv, const FeatArgs& a, cudaStream_t stream) {
    if (plan->dev.r1 == 32) return iv ? launch_v3_one<32, true>(plan, a, stream) : launch_v3_one<32, false>(plan, a, stream);
    return iv ? launch_v3_one<30, true>(plan, a, stream) : launch_v3_one<30, false>(plan, a, stream);
}

}  // namespace seld
