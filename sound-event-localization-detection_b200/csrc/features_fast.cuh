// Fast path of the fused feature kernel (K1+K2): framing + Hann + real FFT (two channels per complex FFT) + power +
// FOA intensity vectors + mel projection + 10*log10 (+ scaler partials | + normalise / model-stem layout / bf16)
// for the reference's own configurations (4 channels, 64 HTK mels, n_fft 1024 or 960 — reference
// dataset.py:27-58 with config.py:85-87).  Anything else (other channel counts, other filterbanks, spectrum
// dumps) runs on the generic kernel in features.cu.
//
// What bounds this path on B200 is not HBM but the SM's shared-memory datapath (128 B/clk) and the FP32 pipe; the
// kernel is organised around shared-memory wavefronts per frame:
//   * warps work in groups of four.  Phase A: every warp transforms one frame (all four channels, two packed
//     complex FFTs) and leaves its per-bin features as seven PLANES V[c][k] (P0 P1 P2 P3 I1/E I2/E I3/E) in its own
//     shared-memory region.  What the second channel pair needs from the first (X0, |X1|^2, Re(conj(X0) X1)) is
//     parked in planes 0..3 (same lane, same bin: in place), the transpose tile lives behind them where planes
//     4..6 and the staged output row go later.
//   * Phase B (after a 128-thread named barrier): the four warps project the group's four frames onto the mel
//     filters.  Lane = (frame, channel), so all lanes walk the SAME bins: the filterbank is baked into the
//     instruction stream (mel_baked.h: weights are FFMA immediates, filter boundaries are straight-line code),
//     every plane word is read exactly once and each warp owns a quarter of the filters.  The raw mel energies
//     are staged behind the planes; after the group's next barrier the owning warp applies 10 log10 and writes
//     the frame's row with coalesced 128-byte stores.
//   * Two real channels share one complex FFT, hence one rounding-noise floor (about -125 dB below the louder
//     channel); the reference (dataset.py:46-50) transforms every channel on its own.  The path therefore has two
//     kernels built from this one template:
//       LEAN (BF = false): no level logic in the hot loop.  The mel phase sums each channel's mel energies (= the
//       frame's total power: HTK triangle weights add up to 1 per bin), and at copy-out the owning warp compares
//       the two channels of each pair; a frame whose pair differs by more than 2^kRedoLog2 in power (36 dB) is
//       appended to a redo list in global memory.  Up to that difference the leaked noise changes the quiet
//       channel's log-mel by < 3e-4 dB (tests/test_kernel_paths_gpu.py).
//       BLOCK-FLOATING (BF = true): run right after, over the listed frames only (a persistent grid that reads the
//       list length on the device and returns at once when it is 0).  Per frame and pair it takes max |windowed
//       sample| of both channels (FMNMX3 + CREDUX) and multiplies the quieter one by the exact power of two that
//       brings it to its partner's level before they share an FFT; the inverse factors are applied to the
//       finished mel rows (powers: 2^-2s, intensities: product of the two channel factors), exact as well.
//     Equal-level audio costs one near-empty extra launch; a clip with a near-dead channel costs its frames twice.
//     tools/featbench measures both kernels over a whole batch (BF over everything: +8 %).
//   * Digitally silent channels must give exactly -100 dB like the reference's separate FFTs; a packed FFT leaves the
//     other channel's rounding noise, so a silent frame of a channel is detected (LEAN: the first windowed sample
//     of every lane, then an OR over all samples only if that test passes; BF: max == 0) and a (rare, warp-uniform)
//     fix-up zeroes that channel's planes.
//   * arithmetic is packed where the data come in pairs (add/sub/mul/fma.rn.f32x2 -> FADD2/FMUL2/FFMA2 on
//     sm_100a): butterflies, complex multiplies, window, channel split, power pairs.
//   * the lane's window row and pass-1 twiddle row (96 words every frame re-reads) live in TENSOR MEMORY
//     (tmem_store.cuh): three LDTM.x32 per channel pair, issued ahead of the in-register DFT, instead of 24 LDS.128 — a
//     fifth of the kernel's shared-memory wavefronts.  The block-floating variant keeps shared-memory tables (it has no
//     registers to spare for the addresses and only runs over the redo list).
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "mel_baked.h"
#include "seld_common.h"
#include "tmem_store.cuh"
#include "warp_fft.cuh"

// 1: the per-lane constant rows (window, pass-1 twiddles) live in tensor memory instead of shared memory (tmem_store.cuh)
#ifndef SELD_TMEM
#define SELD_TMEM 1
#endif

namespace seld {

constexpr int kBfMinShift = 3;    // BF kernel: exponent difference below which a pair is transformed as is
constexpr int kBfMaxShift = 60;   // largest applied shift: 2^-120 for the powers stays a normal float
constexpr int kRedoLog2 = 12;     // LEAN kernel: pair power ratio (log2) above which a frame goes to the redo list
constexpr int kRedoCap = 1 << 16; // redo list capacity (frames); beyond it the BF kernel redoes the whole call
constexpr bool kTm = SELD_TMEM != 0;
constexpr int kTmWinCol = 0, kTmTwCol = 32, kTmCols = 128;  // TMEM columns: window row [32], twiddle row [32 x (re, im)]

template <int R1>
struct FastLayout {
    using F = WarpFft<R1>;
    static constexpr int NB = F::NB;                         // 513 / 481
    // plane pitch = 4 (mod 8) words: in the mel phase lane (f, c) reads the float4 at word f*REGION + c*PITCH + k;
    // the 8 lanes of a quarter-warp (one f, c = 0..7) then hit 8 different 16-byte bank groups.
    static constexpr int PITCH = NB + ((4 - NB % 8) + 8) % 8;  // 516 / 484
    static constexpr int TILE_OFF = 4 * PITCH;               // the transpose tile sits behind the four parking planes
    static constexpr int TP = 34;                            // tile row pitch in float2: 272 B = 17 x 16 B
    static constexpr int TILE_WORDS = 2 * R1 * TP;           // R1 rows (one per reader lane) of 34 float2
    // constant tables, one row per LANE so that a lane fetches its values with 128-bit loads; row pitches of
    // 9 x 16 B and 17 x 16 B put the 8 lanes of a quarter-warp on 8 different 16-byte bank groups
    static constexpr int WIN_PITCH = 36;                     // floats:  window[lane + 32 j] / 2 at [lane][j]
    static constexpr int TW_PITCH = 34;                      // float2s: W_N^(lane k) at [lane][k]
    // behind the seven planes: the frame's finished output row (7 x 64, channel pitch 65), staged by the mel phase
    // and copied out as full 128-byte lines by the owning warp.  Frame slot f starts at OUT_OFF + out_skew(f)
    // so that lane (f, c) lands on bank 8 f + c + m.
    static constexpr int OUT_OFF = 7 * PITCH, OUT_PITCH = 65;
    static constexpr int OUT_END = OUT_OFF + 31 + 7 * OUT_PITCH;
    // behind the staged row: 16 words, total mel energy of power channel c accumulated by filter chunk w at [4 c + w]
    static constexpr int ESUM_OFF = (OUT_END + 3) & ~3, ESUM_WORDS = 16;
    static constexpr int REGION_MIN = (TILE_OFF + TILE_WORDS > ESUM_OFF + ESUM_WORDS ? TILE_OFF + TILE_WORDS : ESUM_OFF + ESUM_WORDS);
    static constexpr int REGION = (REGION_MIN + 3) & ~3;
    static __device__ __forceinline__ int out_skew(int f) { return (8 * f - f * REGION) & 31; }
    static_assert(PITCH % 8 == 4 && REGION % 4 == 0, "region layout");
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

// t = m0 * (m1a, m1b);  result = f0 * (f1a, f1b) + t: two packed instructions.  ptxas folds a swap and a negated half of
// the pair (f1a, f1b) into the FFMA2 operand (.LO_HI / .NP), so sums of two products like Re(conj(a) b) pairs cost 2
// instructions instead of 4.
__device__ __forceinline__ float2 mul_fma_pair(float m0, float m1a, float m1b, float f0, float f1a, float f1b) {
    float2 r;
    asm("{.reg .b64 a, b, c, e, t; mov.b64 a, {%2, %2}; mov.b64 b, {%3, %4}; mov.b64 c, {%5, %5}; mov.b64 e, {%6, %7};"
        " mul.rn.f32x2 t, a, b; fma.rn.f32x2 t, e, c, t; mov.b64 {%0, %1}, t;}"
        : "=f"(r.x), "=f"(r.y) : "f"(m0), "f"(m1a), "f"(m1b), "f"(f0), "f"(f1a), "f"(f1b));
    return r;
}

__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

__device__ __forceinline__ float sample_of(const float* p) { return __ldg(p); }
// int16 -> float without the conversion unit (I2F runs at a quarter of the FP32 rate and cost the PCM kernel 13 %): the
// integer 0x4B000000 + 32768 + x is the float 2^23 + 32768 + x, exactly, for every int16 x; one integer add, one float
// add.  (x 1/32768 lives in the window table.)
__device__ __forceinline__ float sample_of(const short* p) {
    return __int_as_float((int)__ldg(p) + 0x4B008000) - 8421376.0f;
}

// edge frames (reflect padding, torch.stft center=True) and clips too short to be padded: rare, kept out of line
template <int R1, class In>
__device__ __noinline__ void fast_load_edge(float2 (&v)[R1], const In* xa, const In* xb, long long start, long long len,
                                            int lane) {
    using F = WarpFft<R1>;
    if (len <= F::HALF) {  // reflect padding is undefined (torch raises): the rows are written as 0, nothing is read
#pragma unroll 4
        for (int j = 0; j < R1; ++j) v[j] = make_float2(0.f, 0.f);
        return;
    }
#pragma unroll 4
    for (int j = 0; j < R1; ++j) {
        const long long idx = F::reflect(start + lane + 32 * j, len);
        v[j] = make_float2(sample_of(xa + idx), sample_of(xb + idx));
    }
}

template <int R1, class In>
__device__ __forceinline__ void fast_load_raw(float2 (&v)[R1], const In* xa, const In* xb, long long start, long long len,
                                              int lane) {
    using F = WarpFft<R1>;
    if ((start >= 0) && (start + F::N <= len)) {
        const In* pa = xa + start + lane;
        const In* pb = xb + start + lane;
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = make_float2(sample_of(pa + 32 * j), sample_of(pb + 32 * j));
    } else {  // via a scratch array so that v itself never has its address taken (it must stay in registers)
        float2 tmp[R1];
        fast_load_edge<R1, In>(tmp, xa, xb, start, len, lane);
#pragma unroll
        for (int j = 0; j < R1; ++j) v[j] = tmp[j];
    }
}

// ---- mel projection of one filter chunk for lane (frame, channel) -------------------------------
// BF: fac = the lane's un-scaling factor of this (frame, channel) (block floating point); LEAN: the sum of the chunk's
// mel energies is returned (the redo test of the copy-out)
template <int NFFT, int CH, bool BF>
__device__ __forceinline__ float fast_mel_chunk(const float* __restrict__ vp, float* __restrict__ orow, float fac) {
    using MB = MelBaked<NFFT>;
    constexpr int M_LO = MB::chunk_m[CH], M_HI = MB::chunk_m[CH + 1];
    constexpr int K0 = MB::chunk_k0[CH], K1 = MB::chunk_k1[CH];
    // filters of even / odd index (filters two apart never overlap) x even / odd bins (two FMA chains per filter)
    float acc00 = 0.f, acc01 = 0.f, acc10 = 0.f, acc11 = 0.f, esum = 0.f;
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
                if constexpr (BF) sum *= fac;
                else esum += sum;
                orow[m] = sum;  // mel energy / mel-binned IV; the copy-out applies 10 log10 and row validity
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
    return esum;
}

struct FastCtx {      // one frame: clip b, frame t; flags: 1 exists, 2 valid (a frame of the clip)
    unsigned b, t, item;
    int flags;
};

// BF: false = LEAN kernel over all frames of the call (appends to the redo list); true = block-floating kernel, over the
//     redo list (a.redo_mode = 1; over everything when the list overflowed) or over all frames (a.redo_mode = 0).
// EPI: 0 = float32 (B, T, C, 64) rows; 2 = run-time options (normalise with mean / inv_std, (B, C, T, 64), bfloat16).
// STRIP > 0 (tools/featbench only): stages removed from the END of the pipeline to measure what the rest costs —
// 1 no row copy-out, 2 + no mel phase, 3 + no per-bin features / planes, 4 + no channel split, 5 + no FFT (loads only).
template <int R1, bool IV, bool IN16, int EPI, int WARPS, bool BF, int STRIP>
__global__ void __launch_bounds__(WARPS * 32, 1) features_fast_kernel(PlanDev p, FeatArgs a) {
    using F = WarpFft<R1>;
    using L = FastLayout<R1>;
    using In = std::conditional_t<IN16, short, float>;
    constexpr int N = F::N, NB = F::NB;
    constexpr int NCH = IV ? 7 : 4, NF = NCH * 64;
    constexpr int G = WARPS / 4;
    // the block-floating kernel keeps its tables in shared memory: it has no registers to spare for the tensor-memory
    // addresses (88 bytes of spills in the frame loop made it 12 % slower), and it only runs over the redo list
    constexpr bool TM = kTm && !BF;
    // tensor-memory loads of the float32 kernel are issued ahead of the code that hides them and completed later (-1.5 %
    // against issue + wait in one place); with the int16 / run-time-option epilogues that form measured 0.3-1.8 % slower
    constexpr bool kTmAsync = EPI == 0 && !IN16;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    // work items (frames) are 32-bit: the ABI rejects B * T_out >= 2^31
    unsigned n_items = (unsigned)a.n_items;
    const unsigned T_out = (unsigned)a.T_out;
    const unsigned* list = nullptr;  // BF kernel in list mode: the frames the lean kernel flagged
    if (BF && a.redo_mode) {
        const unsigned cnt = a.redo[0];
        if (cnt == 0) return;  // the common case: nothing to redo, leave before touching shared memory
        if (cnt <= (unsigned)kRedoCap) {
            list = a.redo + 4;
            n_items = cnt;
        }  // else: the list overflowed, redo every frame of the call
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* s_win = reinterpret_cast<float*>(smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(s_win + 32 * L::WIN_PITCH);
    float* s_regions = TM ? reinterpret_cast<float*>(smem_raw) : reinterpret_cast<float*>(s_tw + 32 * L::TW_PITCH);
    float* s_norm = s_regions + WARPS * L::REGION;                 // EPI == 2: [2][NF] mean, 1/std
    uint32_t tm_base = 0, tm_lane = 0;

    if constexpr (TM) {
        // constant rows of a lane -> tensor memory: thread i of a warp owns TMEM lane 32 (warp % 4) + i, so warps 0..3
        // fill the four lane quarters and every warp reads the quarter of its own position
        __shared__ uint32_t s_tm_slot;
        if (warp == 0) tmem::alloc<kTmCols>(&s_tm_slot);
        tmem::fence_before_sync();
        __syncthreads();
        tmem::fence_after_sync();
        tm_base = s_tm_slot;
        tm_lane = tmem::lane_base(tm_base, warp);
        if (warp < 4) {
            float r[16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = 16 * c + i;
                    r[i] = j < R1 ? p.window[lane + 32 * j] * (IN16 ? (1.0f / 32768.0f) : 1.0f) : 0.f;
                }
                tmem::st16(tm_lane + kTmWinCol + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = 8 * c + i;
                    const float2 t = k < R1 ? p.twiddle[k * 32 + lane] : make_float2(0.f, 0.f);
                    r[2 * i] = t.x;
                    r[2 * i + 1] = t.y;
                }
                tmem::st16(tm_lane + kTmTwCol + 16 * c, r);
            }
            tmem::wait_st();
        }
        tmem::fence_before_sync();
    } else {
        for (int i = threadIdx.x; i < 32 * L::WIN_PITCH; i += blockDim.x) {
            const int l = i / L::WIN_PITCH, j = i - l * L::WIN_PITCH;
            s_win[i] = j < R1 ? p.window[l + 32 * j] * (IN16 ? (1.0f / 32768.0f) : 1.0f) : 0.f;
        }
        for (int i = threadIdx.x; i < 32 * L::TW_PITCH; i += blockDim.x) {
            const int l = i / L::TW_PITCH, k = i - l * L::TW_PITCH;
            s_tw[i] = k < R1 ? p.twiddle[k * 32 + l] : make_float2(0.f, 0.f);
        }
    }
    if constexpr (EPI == 2)
        for (int i = threadIdx.x; i < NF; i += blockDim.x) {
            s_norm[i] = a.mean ? a.mean[a.c_off * 64 + i] : 0.f;
            s_norm[NF + i] = a.inv_std ? a.inv_std[a.c_off * 64 + i] : 1.f;
        }
    __syncthreads();
    if constexpr (TM) tmem::fence_after_sync();

    const int group = warp >> 2, wi = warp & 3;
    float* region = s_regions + warp * L::REGION;
    float* gregion = s_regions + (group * 4) * L::REGION;
    float2* T = reinterpret_cast<float2*>(region + L::TILE_OFF);
    const int src = F::partner_lane(lane);
    const bool active = R1 == 32 || lane < R1;
    const int bar_id = 1 + group;

    const unsigned n_gitems = (n_items + 3u) >> 2;
    unsigned gidx = blockIdx.x * G + group;
    const unsigned gstride = gridDim.x * G;

    const In* audio = reinterpret_cast<const In*>(a.audio);
    // clip lengths fit 31 bits (the ABI rejects longer clips): 32-bit arithmetic, no 64-bit division in the loop
    const int n_samples = (int)a.n_samples, hop = p.hop;
    const int frames_all = 1 + n_samples / hop;
    auto len_of = [&](unsigned b) -> int { return a.lengths ? (int)min(a.lengths[b], (long long)0x7fffffff) : n_samples; };
    // validity of frame (b, t) of work slot `slot`
    auto classify = [&](FastCtx& c, unsigned slot) {
        const bool exists = slot < n_items;
        int frames = frames_all;
        bool too_short = false;
        if (a.lengths) {  // ragged batch (uniform branch)
            const int len = len_of(c.b);
            too_short = len <= F::HALF;
            if ((!BF || !a.redo_mode) && too_short && exists && c.t == 0 && lane == 0) atomicOr(a.status, 1);
            frames = 1 + len / hop;
        }
        const bool valid = exists && (int)c.t < frames && !too_short;
        c.flags = (exists ? 1 : 0) | (valid ? 2 : 0);
    };
    auto make_ctx = [&](unsigned slot) {  // slot = position in the work list (or the frame index itself)
        FastCtx c;
        c.item = slot < n_items ? (list ? list[slot] : slot) : 0u;  // slots past the end transform frame 0 and are never copied out
        c.b = c.item / T_out;
        c.t = c.item - c.b * T_out;
        classify(c, slot);
        return c;
    };
    // lean kernel: the group's next frame is 4 * gstride frames further; (b, t) advance without a division
    const unsigned step_b = (4u * gstride) / T_out, step_t = 4u * gstride - step_b * T_out;
    auto next_ctx = [&](const FastCtx& cur, unsigned slot) {
        if (BF) return make_ctx(slot);
        FastCtx c;
        c.item = slot;
        c.b = cur.b + step_b;
        c.t = cur.t + step_t;
        if (c.t >= T_out) {
            c.t -= T_out;
            ++c.b;
        }
        if (slot >= n_items) c.item = c.b = c.t = 0u;
        classify(c, slot);
        return c;
    };
    // request the raw samples of channels (ch, ch + 1) of a frame
    auto request = [&](float2 (&v)[R1], const FastCtx& c, int ch) {
        const In* xa = audio + (long long)c.b * a.clip_stride + (long long)ch * a.chan_stride;
        const long long start = (c.flags & 2) ? (long long)c.t * hop - F::HALF : 0ll;
        fast_load_raw<R1, In>(v, xa, xa + a.chan_stride, start, len_of(c.b), lane);
    };

    // copy this warp's staged output row (7 x 64 values) to global memory, 128 bytes per float32 store
    // (10 log10 of the power channels happens here; rows beyond the clip's last frame are written as 0)
    int prev_flags = 0;
    unsigned prev_item = 0;
    long long prev_off = 0;  // element offset of out[b, t, c_off, 0] (or out[b, c_off, t, 0])
    auto copy_out = [&]() {
        if constexpr (STRIP >= 1) {
            if (!a.sink) return;
        }
        if (prev_flags & 1) {
            const float* stage = region + L::OUT_OFF + L::out_skew(wi) + lane;
            const long long cs = (EPI == 2 && a.out_ctf) ? (long long)T_out * 64 : 64;
            const bool ok = prev_flags & 2;
            bool redo = false;
            if constexpr (!BF) {
                // Redo test: total mel energy of the four power channels (16 chunk partials behind the staged row).
                // A pair whose powers differ by more than 2^kRedoLog2 goes to the block-floating kernel; an exactly
                // silent channel (energy 0) leaks nothing and needs nothing.  (Straight-line code: the shuffles
                // interleave with the row copy below; the list append comes after it.)
                float e = lane < 16 ? region[L::ESUM_OFF + lane] : 0.f;
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);       // lanes 4c..4c+3: channel c
                const float ep = __shfl_xor_sync(0xffffffffu, e, 4);  // the partner channel of the pair
                const float hi = fmaxf(e, ep), lo = fminf(e, ep);
                redo = lane < 16 && lo > 0.f && !(hi <= lo * (float)(1 << kRedoLog2));  // (NaN / inf: redo)
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float x0 = stage[c * L::OUT_PITCH], x1 = stage[c * L::OUT_PITCH + 32];
                if (c < 4) {
                    x0 = power_to_db(x0);
                    x1 = power_to_db(x1);
                }
                if constexpr (EPI == 2) {
                    x0 = (x0 - s_norm[c * 64 + lane]) * s_norm[NF + c * 64 + lane];
                    x1 = (x1 - s_norm[c * 64 + 32 + lane]) * s_norm[NF + c * 64 + 32 + lane];
                }
                x0 = ok ? x0 : 0.f;
                x1 = ok ? x1 : 0.f;
                const long long o = prev_off + c * cs + lane;
                if (EPI == 2 && a.out_bf16) {
                    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.out);
                    ob[o] = __float2bfloat16_rn(x0);
                    ob[o + 32] = __float2bfloat16_rn(x1);
                } else {
                    a.out[o] = x0;
                    a.out[o + 32] = x1;
                }
            }
            if constexpr (!BF) {
                if (__any_sync(0xffffffffu, redo) && ok && lane == 0) {
                    const unsigned slot = atomicAdd(a.redo, 1u);
                    if (slot < (unsigned)kRedoCap) a.redo[4 + slot] = prev_item;
                }
            }
        }
    };

    // BF kernel: small per-frame values travel in the three padding words behind every plane (PITCH = NB + 3; nothing
    // else touches them): word NB of plane c = un-scaling factor of output channel c (for the mel phase); word NB + 1
    // of planes 0, 1 = block-floating factors of pair a, kept for pair b's epilogue.
    static_assert(L::PITCH == NB + 3, "plane padding");
    auto pad = [&](float* reg, int plane, int w) -> float& { return reg[plane * L::PITCH + NB + w]; };

    if (gidx < n_gitems) {  // (groups without work skip to the end: the BF kernel's last CTA resets the list)
    FastCtx cur = make_ctx(4 * gidx + wi);
    float2 v[R1];  // raw samples of the channel pair about to be transformed, requested one phase ahead
    request(v, cur, 0);

    while (true) {
        const unsigned gnext = gidx + gstride;
        const bool more = gnext < n_gitems;

#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            // ---- window + pass 1 + twiddle (registers only) ----
            // this lane's window row (a constant table: fetched before the barrier so that the loads are in flight)
            float wv[32];
            if constexpr (TM) {
                if constexpr (kTmAsync) tmem::ld32_issue(tm_lane + kTmWinCol, wv);
            } else {
                const float4* wrow = reinterpret_cast<const float4*>(s_win + lane * L::WIN_PITCH);
#pragma unroll
                for (int i = 0; i < (R1 + 3) / 4; ++i) {
                    const float4 w4 = wrow[i];
                    wv[4 * i] = w4.x, wv[4 * i + 1] = w4.y, wv[4 * i + 2] = w4.z, wv[4 * i + 3] = w4.w;
                }
            }
            // the other warps of the group have finished reading this region (previous mel phase); the previous
            // frame's row is complete (all four filter chunks staged): its copy-out overlaps the first pass below
            if (pr == 0) {
                if constexpr (STRIP < 3) group_barrier(bar_id);
                copy_out();
            }
            bool sil_a, sil_b;
            if constexpr (!BF) {
                // Digitally silent channel?  Cheap necessary test first (every lane's first sample is +-0); the OR over
                // all raw samples only runs when it passes (rare, warp-uniform).
                sil_a = !__any_sync(0xffffffffu, (__float_as_uint(v[0].x) << 1) != 0u);
                sil_b = !__any_sync(0xffffffffu, (__float_as_uint(v[0].y) << 1) != 0u);
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
            }
            // window (pre-scaled by 1/2; by 1/65536 for int16 input)
            if constexpr (TM) {
                if constexpr (kTmAsync) tmem::ld32_wait(wv);
                else tmem::ld32(tm_lane + kTmWinCol, wv);
            }
#pragma unroll
            for (int j = 0; j < R1; ++j) v[j] = cscale(v[j], wv[j]);
            float inv_a = 1.f, inv_b = 1.f;  // BF: 1 / factor applied to channel a, b
            if constexpr (BF) {
                // Level of the two channels in this frame: max |windowed sample| over the warp (what enters the FFT).
                // 0 <=> silent frame of that channel; exponents apart <=> the quieter channel is scaled up by an exact
                // power of two before it shares an FFT with the louder one.
                float ma[4] = {0.f, 0.f, 0.f, 0.f}, mb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < R1; ++j) {
                    ma[j & 3] = fmaxf(ma[j & 3], fabsf(v[j].x));
                    mb[j & 3] = fmaxf(mb[j & 3], fabsf(v[j].y));
                }
                const unsigned ua = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(fmaxf(ma[0], ma[1]), fmaxf(ma[2], ma[3]))));
                const unsigned ub = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3]))));
                sil_a = ua == 0u;
                sil_b = ub == 0u;
                int sh = (int)(ua >> 23) - (int)(ub >> 23);  // > 0: channel a is the louder one
                sh = (sil_a || sil_b) ? 0 : sh;
                sh = (sh > -kBfMinShift && sh < kBfMinShift) ? 0 : max(-kBfMaxShift, min(kBfMaxShift, sh));
                if (sh != 0) {  // warp-uniform
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
            }
            float2 u[32];
            if constexpr (STRIP < 5) {
                if constexpr (TM) {
                    float tw0[32], tw1[32];  // issued ahead of the in-register DFT, which hides the access
                    if constexpr (kTmAsync) {
                        tmem::ld32_issue(tm_lane + kTmTwCol, tw0);
                        tmem::ld32_issue(tm_lane + kTmTwCol + 32, tw1);
                        Dft<R1, false>::run(v);
                        tmem::ld32_wait(tw0);
                        tmem::ld32_wait(tw1);
                    } else {
                        Dft<R1, false>::run(v);
                        tmem::ld32(tm_lane + kTmTwCol, tw0);
                        tmem::ld32(tm_lane + kTmTwCol + 32, tw1);
                    }
                    static_for<R1>([&](auto Kc) {
                        constexpr int k = decltype(Kc)::value;
                        if constexpr (k >= 1 && k < 16) v[k] = cmul(v[k], make_float2(tw0[2 * k], tw0[2 * k + 1]));
                        if constexpr (k >= 16) v[k] = cmul(v[k], make_float2(tw1[2 * (k - 16)], tw1[2 * (k - 16) + 1]));
                    });
                } else {
                    Dft<R1, false>::run(v);
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
            } else {
                if (a.sink) static_for<R1>([&](auto K) { a.sink[decltype(K)::value * 32 + lane] = v[decltype(K)::value].x + v[decltype(K)::value].y + inv_a; });
            }
            // v is dead: request the NEXT channel pair — pair b of this frame, or pair a of the group's next frame.
            // The loads are issued here so that they interleave with the arithmetic of pass 2 and land before
            // the next window pass (one code site for both cases).
            if (pr == 0 || more) request(v, pr == 0 ? cur : next_ctx(cur, 4 * gnext + wi), pr == 0 ? 2 : 0);
            if constexpr (STRIP >= 5) continue;
            F::pass2(u);
            if constexpr (STRIP >= 4) {
                if (a.sink) static_for<32>([&](auto K) { a.sink[decltype(K)::value * 32 + lane] = u[decltype(K)::value].x + u[decltype(K)::value].y; });
                continue;
            }

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
            if constexpr (STRIP >= 3) {
                static_for<16>([&](auto KH) {
                    float2 sS, dD;
                    split(KH, sS, dD);
                    if (a.sink) a.sink[decltype(KH)::value * 32 + lane] = sS.x + sS.y + dD.x + dD.y;
                });
                continue;
            }
            if (pr == 0) {
                // ---- park X0 = (s.x, d.y), P1 = |X1|^2 and I1 = Re(conj(X0) X1) in planes 0..3 (same lane reads them back) ----
                auto park = [&](int k, float2 sS, float2 dD) {
                    // (I1, P1) = s.y * (s.x, s.y) + d.x * (-d.y, d.x)   [X0 = (s.x, d.y), X1 = (s.y, -d.x)]
                    const float2 ip = mul_fma_pair(sS.y, sS.x, sS.y, dD.x, -dD.y, dD.x);
                    region[k] = sS.x;
                    region[L::PITCH + k] = dD.y;
                    region[2 * L::PITCH + k] = ip.y;
                    region[3 * L::PITCH + k] = ip.x;
                };
                static_for<16>([&](auto KH) {
                    float2 sS, dD;
                    split(KH, sS, dD);
                    if (active) park(lane + R1 * decltype(KH)::value, sS, dD);
                });
                if (lane == 0) {
                    park(NB - 1, cadd(u[16], u[16]), make_float2(0.f, 0.f));
                    if constexpr (BF) {
                        pad(region, 0, 1) = inv_a;
                        pad(region, 1, 1) = inv_b;
                    }
                }
                if (sil_a || sil_b) {  // rare: a digitally silent channel must give an exactly-zero spectrum
                    const float ka = sil_a ? 0.f : 1.f, kb = sil_b ? 0.f : 1.f;
                    __syncwarp();
                    for (int k = lane; k < NB; k += 32) {
                        region[k] *= ka;
                        region[L::PITCH + k] *= ka;
                        region[2 * L::PITCH + k] *= kb;
                        region[3 * L::PITCH + k] *= ka * kb;
                    }
                }
                if constexpr (BF) __syncwarp();  // (the pad words written by lane 0 are read by every lane in the next pass)
            } else {
                // ---- per-bin features -> seven planes (planes 4..6 overlay the dead transpose tile) ----
                // BF: planes hold the SCALED powers |X'_c|^2 = |X_c|^2 / inv_c^2 and the intensities of the scaled
                // spectra over the TRUE energy E; the mel phase multiplies the finished rows by the pad-word factors.
                float inv0 = 1.f, inv1 = 1.f;
                if constexpr (BF) {
                    inv0 = pad(region, 0, 1);
                    inv1 = pad(region, 1, 1);
                }
                const float k0 = inv0 * inv0, k1 = inv1 * inv1 * (1.0f / 3.0f), k2 = inv_a * inv_a * (1.0f / 3.0f),
                            k3 = inv_b * inv_b * (1.0f / 3.0f);
                auto finish = [&](int k, float2 sS, float2 dD) {
                    const float x0r = region[k], x0i = region[L::PITCH + k];
                    const float p1 = region[2 * L::PITCH + k], i1 = region[3 * L::PITCH + k];
                    const float2 p23 = pow_pair(sS, dD);  // (|X2|^2, |X3|^2)
                    const float p0 = fmaf(x0r, x0r, x0i * x0i);
                    region[k] = p0;
                    region[L::PITCH + k] = p1;
                    region[2 * L::PITCH + k] = p23.x;
                    region[3 * L::PITCH + k] = p23.y;
                    if (IV) {
                        // (I2, I3) = (Re(conj(X0) X2), Re(conj(X0) X3)) = x0r * (s.x, s.y) + x0i * (d.y, -d.x)
                        const float2 i23 = mul_fma_pair(x0r, sS.x, sS.y, x0i, dD.y, -dD.x);
                        float e;
                        if constexpr (BF) e = fmaf(p23.y, k3, fmaf(p23.x, k2, fmaf(p1, k1, fmaf(p0, k0, kEpsIV))));
                        else e = kEpsIV + p0 + (p1 + p23.x + p23.y) * (1.0f / 3.0f);
                        float inv;  // MUFU.RCP alone (1 ulp): e >= 1e-8 is never denormal; the IV tolerance is 1e-4 relative
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(e));
                        const float2 i23n = cscale(i23, inv);
                        region[4 * L::PITCH + k] = i1 * inv;
                        region[5 * L::PITCH + k] = i23n.x;
                        region[6 * L::PITCH + k] = i23n.y;
                    }
                };
                static_for<16>([&](auto KH) {
                    float2 sS, dD;
                    split(KH, sS, dD);
                    if (active) finish(lane + R1 * decltype(KH)::value, sS, dD);
                });
                if (lane == 0) finish(NB - 1, cadd(u[16], u[16]), make_float2(0.f, 0.f));
                if (sil_a || sil_b) {  // rare: redo the planes with the silent channel (2 = a, 3 = b) at exactly 0
                    __syncwarp();
                    for (int k = lane; k < NB; k += 32) {
                        const float p0 = region[k], p1 = region[L::PITCH + k], p2o = region[2 * L::PITCH + k], p3o = region[3 * L::PITCH + k];
                        const float p2 = sil_a ? 0.f : p2o, p3 = sil_b ? 0.f : p3o;
                        region[2 * L::PITCH + k] = p2;
                        region[3 * L::PITCH + k] = p3;
                        if (IV) {
                            const float base = fmaf(p1, k1, fmaf(p0, k0, kEpsIV));
                            const float e_old = fmaf(p3o, k3, fmaf(p2o, k2, base));
                            const float e_new = fmaf(p3, k3, fmaf(p2, k2, base));
                            const float g = e_old / e_new;
                            region[4 * L::PITCH + k] *= g;
                            region[5 * L::PITCH + k] = sil_a ? 0.f : region[5 * L::PITCH + k] * g;
                            region[6 * L::PITCH + k] = sil_b ? 0.f : region[6 * L::PITCH + k] * g;
                        }
                    }
                }
                // BF: un-scaling factors of the frame's output channels (the tile is dead: every lane has loaded its row)
                if (BF && lane == 0) {
                    pad(region, 0, 0) = inv0 * inv0;
                    pad(region, 1, 0) = inv1 * inv1;
                    pad(region, 2, 0) = inv_a * inv_a;
                    pad(region, 3, 0) = inv_b * inv_b;
                    if (IV) {
                        pad(region, 4, 0) = inv0 * inv1;
                        pad(region, 5, 0) = inv0 * inv_a;
                        pad(region, 6, 0) = inv0 * inv_b;
                    }
                }
            }
        }
        prev_flags = cur.flags;
        prev_item = cur.item;
        if (EPI == 2 && a.out_ctf) prev_off = (((long long)cur.b * a.C_out + a.c_off) * T_out + cur.t) * 64;
        else prev_off = (((long long)cur.b * T_out + cur.t) * a.C_out + a.c_off) * 64;
        if constexpr (STRIP >= 3) {
            if (!more) break;
            gidx = gnext;
            cur = next_ctx(cur, 4 * gidx + wi);
            continue;
        }
        group_barrier(bar_id);  // the four frames of the group are in their planes, the previous rows are copied out

        // ---- mel phase: lane = (frame, channel); this warp owns filter chunk wi ----
        if constexpr (STRIP < 2) {
            const int f = lane >> 3, c = lane & 7;
            if (c < NCH) {  // (slots of frames that do not exist hold finite dummy planes; their rows are never copied out)
                float* rf = gregion + f * L::REGION;
                const float* vp = rf + c * L::PITCH;
                float* orow = rf + L::OUT_OFF + L::out_skew(f) + c * L::OUT_PITCH;
                const float fac = BF ? vp[NB] : 1.0f;
                float esum;
                switch (wi) {
                    case 0: esum = fast_mel_chunk<N, 0, BF>(vp, orow, fac); break;
                    case 1: esum = fast_mel_chunk<N, 1, BF>(vp, orow, fac); break;
                    case 2: esum = fast_mel_chunk<N, 2, BF>(vp, orow, fac); break;
                    default: esum = fast_mel_chunk<N, 3, BF>(vp, orow, fac); break;
                }
                if (!BF && c < 4) rf[L::ESUM_OFF + 4 * c + wi] = esum;
            }
        }
        if (!more) break;
        gidx = gnext;
        cur = next_ctx(cur, 4 * gidx + wi);
    }
    if constexpr (STRIP < 3) {
        group_barrier(bar_id);
        copy_out();
    }
    }  // gidx < n_gitems
    if (BF && a.redo_mode) {  // last CTA out clears the list for the next call on this stream slot
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.redo + 1, 1u) == gridDim.x - 1) {
                a.redo[0] = 0u;
                a.redo[1] = 0u;
            }
        }
    }
    if constexpr (TM) {
        tmem::fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem::dealloc<kTmCols>(tm_base);
    }
}

template <int R1, int EPI, int WARPS, bool IV, bool BF>
constexpr size_t fast_smem_bytes() {
    using L = FastLayout<R1>;
    return (kTm && !BF ? 0 : sizeof(float) * (32 * L::WIN_PITCH) + sizeof(float2) * (32 * L::TW_PITCH)) +
           sizeof(float) * ((size_t)WARPS * L::REGION + (EPI == 2 ? 2 * (IV ? 7 : 4) * 64 : 0));
}

// one-time per device (plan creation): opt in to the large dynamic shared memory carve-out
template <int R1, bool IV, bool IN16, int EPI, int WARPS, bool BF, int STRIP = 0>
static int fast_configure() {
    auto kern = features_fast_kernel<R1, IV, IN16, EPI, WARPS, BF, STRIP>;
    SELD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)fast_smem_bytes<R1, EPI, WARPS, IV, BF>()));
    return SELD_OK;
}

// a.redo_mode (BF kernel only): persistent grid over the redo list, one CTA per SM
template <int R1, bool IV, bool IN16, int EPI, int WARPS, bool BF, int STRIP = 0>
static int fast_launch(const seld_plan* plan, const FeatArgs& a, cudaStream_t stream) {
    static_assert(fast_smem_bytes<R1, EPI, WARPS, IV, BF>() <= (size_t)kMaxSmemOptin, "shared memory budget (227 KB)");
    const long long n_gitems = (a.n_items + 3) / 4;
    long long ctas = (n_gitems + WARPS / 4 - 1) / (WARPS / 4);
    if (ctas > plan->num_sms) ctas = plan->num_sms;
    if (ctas < 1) return SELD_OK;
    features_fast_kernel<R1, IV, IN16, EPI, WARPS, BF, STRIP>
        <<<(unsigned)ctas, WARPS * 32, fast_smem_bytes<R1, EPI, WARPS, IV, BF>(), stream>>>(plan->dev, a);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

// per-(R1, input type) translation units (features_fast_*.cu) export these two:
//   configure: cudaFuncSetAttribute for every variant of the unit (called at plan creation, on the plan's device)
//   launch   : run-time dispatch on (iv, epi, warps, bf)
#define SELD_FAST_UNIT_DECL(NAME)           \
    int fast_configure_##NAME();            \
    int fast_launch_##NAME(const seld_plan* plan, bool iv, int epi, int warps, bool bf, const FeatArgs& a, cudaStream_t stream);

SELD_FAST_UNIT_DECL(r32_f32)
SELD_FAST_UNIT_DECL(r30_f32)
SELD_FAST_UNIT_DECL(r32_i16)
SELD_FAST_UNIT_DECL(r30_i16)

#define SELD_FAST_DISPATCH(R1, IN16, EPI, W, BF) \
    (iv ? fast_launch<R1, true, IN16, EPI, W, BF>(plan, a, s) : fast_launch<R1, false, IN16, EPI, W, BF>(plan, a, s))
#define SELD_FAST_UNIT_DEFINE(NAME, R1, IN16, WITH_W8)                                                              \
    int fast_configure_##NAME() {                                                                                    \
        int rc = SELD_OK;                                                                                            \
        auto acc = [&](int r) { if (rc == SELD_OK) rc = r; };                                                        \
        acc(fast_configure<R1, false, IN16, 0, 12, false>()); acc(fast_configure<R1, true, IN16, 0, 12, false>());   \
        acc(fast_configure<R1, false, IN16, 0, 12, true>()); acc(fast_configure<R1, true, IN16, 0, 12, true>());     \
        acc(fast_configure<R1, false, IN16, 2, 12, false>()); acc(fast_configure<R1, true, IN16, 2, 12, false>());   \
        acc(fast_configure<R1, false, IN16, 2, 12, true>()); acc(fast_configure<R1, true, IN16, 2, 12, true>());     \
        if (WITH_W8) {                                                                                               \
            acc(fast_configure<R1, false, IN16, 0, 8, false>()); acc(fast_configure<R1, true, IN16, 0, 8, false>()); \
            acc(fast_configure<R1, false, IN16, 0, 8, true>()); acc(fast_configure<R1, true, IN16, 0, 8, true>());   \
        }                                                                                                            \
        return rc;                                                                                                   \
    }                                                                                                                \
    int fast_launch_##NAME(const seld_plan* plan, bool iv, int epi, int warps, bool bf, const FeatArgs& a, cudaStream_t s) { \
        if (WITH_W8 && warps == 8 && epi == 0)                                                                       \
            return bf ? SELD_FAST_DISPATCH(R1, IN16, 0, (WITH_W8 ? 8 : 12), true) : SELD_FAST_DISPATCH(R1, IN16, 0, (WITH_W8 ? 8 : 12), false); \
        if (epi == 0) return bf ? SELD_FAST_DISPATCH(R1, IN16, 0, 12, true) : SELD_FAST_DISPATCH(R1, IN16, 0, 12, false); \
        return bf ? SELD_FAST_DISPATCH(R1, IN16, 2, 12, true) : SELD_FAST_DISPATCH(R1, IN16, 2, 12, false);          \
    }

}  // namespace seld
