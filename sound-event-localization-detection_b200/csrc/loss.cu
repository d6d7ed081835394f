// N4: the class loss of reference loss.py:27-54 (SMRSELDLoss.class_mse_loss / class_ce_loss) computed
// from COMPACT targets.  The dense (B, T, 648, 14) float32 target tensor (145 MB per batch, > 99.7 % background) is
// never built: a batch's events are painted into a uint16 class-set mask per (window, frame, cell) — 5 MB — and the
// loss kernels read the logits once.
//   mask == 0        <=> no event touches the cell: target = one-hot background (dataset.py:114-117)
//   mask bit c       <=> some event of class c covers the cell: target[c] = 1 (multi-hot when classes overlap; the
//                        background bit is set only by an event of class M-1, like the reference)
// softmax-MSE : mean over all (b, t, g, m) of (softmax(z)_m - y_m)^2                       (loss.py:43-54)
// CE          : nn.CrossEntropyLoss(weight)(z, argmax_m y) — argmax of a multi-hot row is its LOWEST set class
//               (torch.argmax returns the first maximum)                                   (loss.py:27-41)
// Forward kernels add float64 partial sums into d_sums; backward kernels write d loss / d logits scaled by a device
// scalar (the upstream gradient times the mean's 1/N), so nothing synchronises with the host.
#include "seld_common.h"

namespace seld {

constexpr int kLossSliceRows = 5;
__device__ __forceinline__ bool loss_cell_in_region(int gi, int gj, int I, int J, double c_az, double c_el, double two_s_az,
                                                    double two_s_el) {  // == cell_in_region of labels.cu
    const double el_min = fmax(__dsub_rn(c_el, two_s_el), -90.0);
    const double el_max = fmin(__dadd_rn(c_el, two_s_el), 90.0);
    const double cell_el = __dadd_rn(-90.0, __dmul_rn((double)gi + 0.5, 180.0 / (double)I));
    const double cell_az = __dadd_rn(-180.0, __dmul_rn((double)gj + 0.5, 360.0 / (double)J));
    double diff = __dsub_rn(cell_az, c_az);
    for (int it = 0; it < 64 && diff > 180.0; ++it) diff = __dsub_rn(diff, 360.0);
    for (int it = 0; it < 64 && diff < -180.0; ++it) diff = __dadd_rn(diff, 360.0);
    return (fabs(diff) <= two_s_az) && (el_min <= cell_el) && (cell_el <= el_max);
}

// class-set masks of one batch: CTA (s, w) owns rows [s * kLossSliceRows, ...) of window w, clears them and ORs in the
// events of the window that touch them (same event selection as loader_batch_kernel)
__global__ void __launch_bounds__(256) batch_class_mask_kernel(const int* __restrict__ order, int first,
                                                               const int* __restrict__ win_start, const int* __restrict__ win_lo,
                                                               const int* __restrict__ win_hi, int win_len,
                                                               const int4* __restrict__ events, const double2* __restrict__ centres,
                                                               int I, int J, int M, double two_s_az, double two_s_el,
                                                               unsigned short* __restrict__ mask) {
    const int w = blockIdx.y, s = blockIdx.x;
    const int widx = order ? order[first + w] : first + w;
    const long long start = win_start[widx];
    const int f0 = s * kLossSliceRows, f1 = min(win_len, f0 + kLossSliceRows);
    if (f0 >= f1) return;
    const int cells = I * J;
    unsigned short* m = mask + ((long long)w * win_len + f0) * cells;
    const int n = (f1 - f0) * cells;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m[i] = 0;
    const int lo = win_lo[widx], hi = win_hi[widx];
    const long long g0 = start + f0, g1 = start + f1;
    __shared__ int s_hits[256];
    __shared__ int s_n;
    for (int base = lo; base < hi; base += blockDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int e = base + threadIdx.x;
        if (e < hi) {
            const int4 ev = events[e];
            if (min((long long)ev.y, g1) > max((long long)ev.x, g0) && ev.z >= 0 && ev.z < M && ev.w < cells)
                s_hits[atomicAdd(&s_n, 1)] = e;
        }
        __syncthreads();
        const int nh = s_n;
        for (int h = 0; h < nh; ++h) {
            const int e2 = s_hits[h];
            const int4 ev = events[e2];
            const long long r0 = max((long long)ev.x, g0), r1 = min((long long)ev.y, g1);
            const unsigned bit = 1u << ev.z;
            // 16-bit cells are OR-ed through their aligned 32-bit word (atomicOr has no 16-bit form); the slice starts on
            // an even cell index because I * J is even (checked by the launcher)
            unsigned* m32 = reinterpret_cast<unsigned*>(m);
            auto or_at = [&](long long idx) { atomicOr(m32 + (idx >> 1), (idx & 1) ? bit << 16 : bit); };
            if (ev.w >= 0) {
                for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x) or_at((r - g0) * cells + ev.w);
            } else if (centres) {  // (all cells are tested here: the mask path is 3 % of a batch, see labels.cu for the boxed form)
                const double2 c = centres[e2];
                for (int cell = threadIdx.x; cell < cells; cell += blockDim.x) {
                    if (!loss_cell_in_region(cell / J, cell % J, I, J, c.x, c.y, two_s_az, two_s_el)) continue;
                    for (long long r = r0; r < r1; ++r) or_at((r - g0) * cells + cell);
                }
            }
        }
    }
}

// one thread per (b, t, g) cell: M <= 16 logits in registers
template <int M>
__device__ __forceinline__ void softmax_of(const float* __restrict__ z, float (&p)[M], float& lse) {
    float mx = z[0];
#pragma unroll
    for (int i = 1; i < M; ++i) mx = fmaxf(mx, z[i]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        p[i] = expf(z[i] - mx);
        sum += p[i];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] *= inv;
    lse = mx + logf(sum);
}

__device__ __forceinline__ void block_add(double v, double* dst) {
    __shared__ double s_part[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? s_part[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(dst, v);
    }
    __syncthreads();
}

// MODE 0: softmax-MSE, 1: cross entropy.  grad != null: also writes d loss / d logits * (*gscale)
// sums[0] += sum of squared errors (MSE) | sum of w_t * nll (CE);  sums[1] += sum of w_t (CE)
template <int M, int MODE>
__global__ void __launch_bounds__(256) class_loss_kernel(const float* __restrict__ logits, const unsigned short* __restrict__ mask,
                                                         long long n_cells, const float* __restrict__ weight,
                                                         double* __restrict__ sums, float* __restrict__ grad,
                                                         const float* __restrict__ gscale) {
    double acc = 0.0, wacc = 0.0;
    const float gs = grad ? *gscale : 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
        float z[M], p[M], lse;
#pragma unroll
        for (int k = 0; k < M; ++k) z[k] = logits[i * M + k];
        softmax_of<M>(z, p, lse);
        unsigned y = mask[i];
        if (y == 0u) y = 1u << (M - 1);  // untouched cell: one-hot background
        if (MODE == 0) {
            float se = 0.f, dot = 0.f;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const float d = p[k] - ((y >> k) & 1u ? 1.f : 0.f);
                se = fmaf(d, d, se);
                dot = fmaf(d, p[k], dot);
            }
            acc += (double)se;
            if (grad) {  // d/dz_k sum_m (p_m - y_m)^2 = 2 p_k ((p_k - y_k) - sum_m (p_m - y_m) p_m)
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const float d = p[k] - ((y >> k) & 1u ? 1.f : 0.f);
                    grad[i * M + k] = gs * 2.f * p[k] * (d - dot);
                }
            }
        } else {
            const int t = __ffs(y) - 1;  // argmax of a {0,1} row = its first 1
            const float w = weight ? weight[t] : 1.f;
            float zt = z[0];
#pragma unroll
            for (int k = 1; k < M; ++k) zt = k == t ? z[k] : zt;
            acc += (double)(w * (lse - zt));
            wacc += (double)w;
            if (grad) {  // d/dz_k w (lse - z_t) = w (p_k - [k == t]); the caller's gscale carries 1 / sum of w
#pragma unroll
                for (int k = 0; k < M; ++k) grad[i * M + k] = gs * w * (p[k] - (k == t ? 1.f : 0.f));
            }
        }
    }
    if (sums) {
        block_add(acc, sums);
        if (MODE == 1) block_add(wacc, sums + 1);
    }
}

// Tiled form of the same kernel (used when logits / grad are 16-byte aligned, i.e. always from torch): a CTA stages 256
// cells = 3 584 floats through shared memory with coalesced 16-byte loads — one thread per cell reading its 14 logits
// straight from global memory touches every 32-byte sector of the tile 8 times — computes from shared memory (a cell's
// 14 floats as seven 8-byte loads: 56-byte pitch, conflict-free per half-warp) and, for the backward pass, writes the
// gradient tile back in place and out with coalesced 16-byte stores.
template <int M, int MODE, bool BWD>
__global__ void __launch_bounds__(256, 4) class_loss_tiled_kernel(const float* __restrict__ logits,
                                                               const unsigned short* __restrict__ mask, long long n_cells,
                                                               const float* __restrict__ weight, double* __restrict__ sums,
                                                               float* __restrict__ grad, const float* __restrict__ gscale) {
    static_assert(M % 2 == 0, "cells are read as float2");
    constexpr int TILE = 256, TW = TILE * M;
    __shared__ __align__(16) float s_z[TW];
    double acc = 0.0, wacc = 0.0;
    const float gs = BWD ? *gscale : 0.f;
    const long long n_tiles = (n_cells + TILE - 1) / TILE;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long c0 = tile * TILE;
        const int nc = (int)min((long long)TILE, n_cells - c0);
        const int nf = nc * M, nv = nf >> 2;
        const float* src = logits + c0 * M;  // tile * 14 336 bytes: 16-byte aligned when logits is
        for (int i = threadIdx.x; i < nv; i += TILE) reinterpret_cast<float4*>(s_z)[i] = __ldcs(reinterpret_cast<const float4*>(src) + i);
        for (int i = (nv << 2) + threadIdx.x; i < nf; i += TILE) s_z[i] = src[i];
        __syncthreads();
        if ((int)threadIdx.x < nc) {
            float z[M], p[M], lse;
            float2* zp = reinterpret_cast<float2*>(s_z + threadIdx.x * M);
#pragma unroll
            for (int k = 0; k < M / 2; ++k) {
                const float2 t = zp[k];
                z[2 * k] = t.x;
                z[2 * k + 1] = t.y;
            }
            softmax_of<M>(z, p, lse);
            unsigned y = mask[c0 + threadIdx.x];
            if (y == 0u) y = 1u << (M - 1);  // untouched cell: one-hot background
            float g[M];
            if (MODE == 0) {
                float se = 0.f, dot = 0.f;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const float d = p[k] - ((y >> k) & 1u ? 1.f : 0.f);
                    se = fmaf(d, d, se);
                    dot = fmaf(d, p[k], dot);
                }
                acc += (double)se;
                if (BWD) {
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        const float d = p[k] - ((y >> k) & 1u ? 1.f : 0.f);
                        g[k] = gs * 2.f * p[k] * (d - dot);
                    }
                }
            } else {
                const int t = __ffs(y) - 1;
                const float w = weight ? weight[t] : 1.f;
                float zt = z[0];
#pragma unroll
                for (int k = 1; k < M; ++k) zt = k == t ? z[k] : zt;
                acc += (double)(w * (lse - zt));
                wacc += (double)w;
                if (BWD) {
#pragma unroll
                    for (int k = 0; k < M; ++k) g[k] = gs * w * (p[k] - (k == t ? 1.f : 0.f));
                }
            }
            if (BWD) {
#pragma unroll
                for (int k = 0; k < M / 2; ++k) zp[k] = make_float2(g[2 * k], g[2 * k + 1]);
            }
        }
        if (BWD) {
            __syncthreads();
            float* dst = grad + c0 * M;
            for (int i = threadIdx.x; i < nv; i += TILE) __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(s_z)[i]);
            for (int i = (nv << 2) + threadIdx.x; i < nf; i += TILE) dst[i] = s_z[i];
        }
        __syncthreads();  // the next tile overwrites s_z
    }
    if (sums) {
        block_add(acc, sums);
        if (MODE == 1) block_add(wacc, sums + 1);
    }
}

// The two dormant terms of the reference's loss (loss.py:56-146; commented out of SMRSELDLoss.forward, loss.py:158-165) from
// the same compact targets.  One CTA per frame (b, t); "event cell" = some class below the background class covers it
// (mask & low bits != 0), which is both argmax(y_true) != M-1 (loss.py:66-71: the first maximum of a multi-hot row is its
// lowest class) and sum(y_true[:-1]) > 0.01 (loss.py:113-118).
//   AIUR (loss.py:56-88): IoU of {cells whose argmax prediction is not background} and {event cells} per frame
//        (1 when both are empty), summed into sums[2].  argmax has no gradient.
//   converging localisation (loss.py:90-146): y' = 1 on background cells, -N_bac / (N_non + 1e-10) on event cells; y_at =
//        y' + mean over the 8 circular neighbours of (neighbour - y'), in the reference's summation order; the frame adds
//        sum over cells of (1 - p_background as the sum of the 13 event probabilities) * y_at to sums[0] and 1 to sums[1]
//        when it has events.  grad != null: d/d logits of that sum times (*gscale): y_at * p_k * ([k < M-1] - p_nonbg).
template <int M>
__global__ void __launch_bounds__(256) aux_loss_kernel(const float* __restrict__ logits, const unsigned short* __restrict__ mask,
                                                       int I, int J, double* __restrict__ sums, float* __restrict__ grad,
                                                       const float* __restrict__ gscale) {
    extern __shared__ float s_f[];  // [cells] y', [cells] y_at
    __shared__ int s_cnt[3];        // event cells | predicted event cells | both
    const int cells = I * J;
    float* s_y = s_f;
    float* s_at = s_f + cells;
    const long long f = blockIdx.x;
    const unsigned short* mk = mask + f * cells;
    const float* zf = logits + f * cells * M;
    constexpr unsigned kEventBits = (1u << (M - 1)) - 1u;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int n_local = 0;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) n_local += (mk[c] & kEventBits) != 0u;
    for (int o = 16; o > 0; o >>= 1) n_local += __shfl_down_sync(0xffffffffu, n_local, o);
    if ((threadIdx.x & 31) == 0 && n_local) atomicAdd(&s_cnt[0], n_local);
    __syncthreads();
    const int n_non_i = s_cnt[0];
    const float n_non = (float)n_non_i, n_bac = (float)(cells - n_non_i);
    const float ratio = -(n_bac / (n_non + 1e-10f));
    const bool has = n_non_i > 0;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) s_y[c] = (mk[c] & kEventBits) != 0u ? ratio : 1.f;
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const int i = c / J, j = c - i * J;
        const float y = s_y[c];
        float d = 0.f;
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                if (di == 0 && dj == 0) continue;
                const int ni = (i + di + I) % I, nj = (j + dj + J) % J;
                d = __fadd_rn(d, __fsub_rn(s_y[ni * J + nj], y));
            }
        s_at[c] = __fadd_rn(y, __fdiv_rn(d, 8.0f));
    }
    __syncthreads();
    const float gs = grad ? *gscale : 0.f;
    double acc = 0.0;
    int pc = 0, inter = 0;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        float z[M], p[M], lse;
        const float2* zp = reinterpret_cast<const float2*>(zf + (long long)c * M);
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
            const float2 t = zp[k];
            z[2 * k] = t.x;
            z[2 * k + 1] = t.y;
        }
        softmax_of<M>(z, p, lse);
        float nonbg = 0.f, best = p[0];
        int arg = 0;
#pragma unroll
        for (int k = 0; k < M - 1; ++k) nonbg += p[k];
#pragma unroll
        for (int k = 1; k < M; ++k)
            if (p[k] > best) {
                best = p[k];
                arg = k;
            }
        const bool ev = (mk[c] & kEventBits) != 0u, pe = arg != M - 1;
        pc += pe;
        inter += pe && ev;
        const float yat = s_at[c];
        if (has) acc += (double)(nonbg * yat);
        if (grad) {
            const float gy = has ? gs * yat : 0.f;
            float* gp = grad + (f * cells + c) * M;
#pragma unroll
            for (int k = 0; k < M; ++k)  // d/dz_k of sum_{m < M-1} p_m = p_k ([k < M-1] - p_nonbg); 1 - p_nonbg taken as p_background
                gp[k] = gy * p[k] * (k < M - 1 ? p[M - 1] : -nonbg);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        pc += __shfl_down_sync(0xffffffffu, pc, o);
        inter += __shfl_down_sync(0xffffffffu, inter, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (pc) atomicAdd(&s_cnt[1], pc);
        if (inter) atomicAdd(&s_cnt[2], inter);
    }
    if (sums) {
        block_add(acc, sums);  // (contains the __syncthreads that orders the counters above)
        if (threadIdx.x == 0) {
            const float uni = (float)(s_cnt[1] + n_non_i - s_cnt[2]);
            const float iou = uni > 0.f ? (float)s_cnt[2] / (uni + 1e-8f) : 1.f;
            atomicAdd(sums + 2, (double)iou);
            if (has) atomicAdd(sums + 1, 1.0);
        }
    }
}

int launch_aux_losses(const float* logits, const unsigned short* mask, long long n_frames, int I, int J, int M, double* sums,
                      float* grad, const float* gscale, cudaStream_t st) {
    if (n_frames == 0) return SELD_OK;
    if (M != 14) {
        set_error("seld_aux_losses: compiled for the reference's 14 classes (config.py NUM_CLASSES)");
        return SELD_ERR_UNSUPPORTED;
    }
    const long long cells = (long long)I * J;
    if (cells < 1 || cells > 4096 || n_frames > 0x7fffffffll || (reinterpret_cast<uintptr_t>(logits) & 7) != 0) {
        set_error("seld_aux_losses: needs 1 <= I * J <= 4096 cells, < 2^31 frames and 8-byte aligned logits");
        return SELD_ERR_UNSUPPORTED;
    }
    aux_loss_kernel<14><<<(unsigned)n_frames, 256, 2 * cells * sizeof(float), st>>>(logits, mask, I, J, sums, grad, gscale);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

int launch_batch_class_mask(const int* order, int first, int n_win, const int* win_start, const int* win_lo, const int* win_hi,
                            int win_len, const int* events, const double* centres, int I, int J, int M, double sigma_az,
                            double sigma_el, unsigned short* mask, cudaStream_t st) {
    if (n_win == 0 || win_len == 0) return SELD_OK;
    if (M > 16 || (I * J) % 2 != 0 || (reinterpret_cast<uintptr_t>(mask) & 3) != 0) {
        set_error("seld_batch_class_mask: needs n_classes <= 16, an even number of grid cells and a 4-byte aligned mask");
        return SELD_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)((win_len + kLossSliceRows - 1) / kLossSliceRows), (unsigned)n_win);
    batch_class_mask_kernel<<<grid, 256, 0, st>>>(order, first, win_start, win_lo, win_hi, win_len,
                                                  reinterpret_cast<const int4*>(events), reinterpret_cast<const double2*>(centres), I,
                                                  J, M, 2 * sigma_az, 2 * sigma_el, mask);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

int launch_class_loss(int mode, const float* logits, const unsigned short* mask, long long n_cells, int M, const float* weight,
                      double* sums, float* grad, const float* gscale, cudaStream_t st) {
    if (n_cells == 0) return SELD_OK;
    if (M != 14) {
        set_error("seld_class_loss: compiled for the reference's 14 classes (config.py NUM_CLASSES)");
        return SELD_ERR_UNSUPPORTED;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (n_cells + 255) / 256;
    const bool tiled = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
    if (tiled) {
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;  // 8 CTAs of 256 threads per SM, grid-stride over tiles
#define SELD_LOSS_LAUNCH(MODE, BWD) \
    class_loss_tiled_kernel<14, MODE, BWD><<<(unsigned)blocks, 256, 0, st>>>(logits, mask, n_cells, weight, sums, grad, gscale)
        if (mode == 0) { if (grad) SELD_LOSS_LAUNCH(0, true); else SELD_LOSS_LAUNCH(0, false); }
        else { if (grad) SELD_LOSS_LAUNCH(1, true); else SELD_LOSS_LAUNCH(1, false); }
#undef SELD_LOSS_LAUNCH
        SELD_CUDA_TRY(cudaGetLastError());
        return SELD_OK;
    }
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (mode == 0) class_loss_kernel<14, 0><<<(unsigned)blocks, 256, 0, st>>>(logits, mask, n_cells, weight, sums, grad, gscale);
    else class_loss_kernel<14, 1><<<(unsigned)blocks, 256, 0, st>>>(logits, mask, n_cells, weight, sums, grad, gscale);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

}  // namespace seld
