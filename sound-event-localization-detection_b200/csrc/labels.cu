// K4: dense SELD grid label encoder (fill + paint), K6: window/batch gather, scaler apply.
//
// Reference semantics reproduced (bit-exact):
//   dataset.py:84            labels = zeros(T, I*J, M)
//   dataset.py:100-111       labels[t, cell, class] = 1.0 for the 5 frames of each CSV row
//   dataset.py:114-117       labels[t, cell, M-1] = 1.0 for every (t, cell) that no row touched
//   smrl_seld_gaussian.py:474-518  region variant: every cell whose centre is inside the +-2 sigma rectangle
//   dataset.py:267-317       windows [50k, 50k+250) with zero / background padding of the tail
// "fill" writes the no-event state everywhere at HBM write bandwidth; "paint" then touches only the few
// (row, cell) pairs that events cover: pass 0 clears the background class, pass 1 sets the event class (a
// class M-1 event therefore leaves the background at 1, like the reference).
#include "seld_common.h"

namespace seld {

// ---- fill: period-M pattern (0,...,0,1), 128-bit stores --------------------------------------------
__global__ void labels_fill_vec4(float4* __restrict__ out, long long n_vec, int M) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    int r = int((4 * i) % M);                         // class index of the first lane element
    const int step = int((4 * stride) % M);
    for (; i < n_vec; i += stride) {
        int r1 = r + 1; if (r1 >= M) r1 -= M;
        int r2 = r1 + 1; if (r2 >= M) r2 -= M;
        int r3 = r2 + 1; if (r3 >= M) r3 -= M;
        float4 v = make_float4(r == M - 1 ? 1.f : 0.f, r1 == M - 1 ? 1.f : 0.f, r2 == M - 1 ? 1.f : 0.f,
                               r3 == M - 1 ? 1.f : 0.f);
        __stcs(out + i, v);                            // streaming store: written once, not re-read here
        r += step; if (r >= M) r -= M;
    }
}

__global__ void labels_fill_scalar(float* __restrict__ out, long long n, int M) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (i % M == M - 1) ? 1.f : 0.f;
}

// ---- paint ---------------------------------------------------------------------------------------
// Region membership, float64, same operations in the same order as smrl_seld_gaussian.py:474-514.
// __d*_rn intrinsics forbid FMA contraction so every step rounds like CPython's float arithmetic.
__device__ __forceinline__ bool cell_in_region(int gi, int gj, int I, int J, double c_az, double c_el, double two_s_az,
                                               double two_s_el) {
    const double el_min = fmax(__dsub_rn(c_el, two_s_el), -90.0);
    const double el_max = fmin(__dadd_rn(c_el, two_s_el), 90.0);
    const double cell_el = __dadd_rn(-90.0, __dmul_rn((double)gi + 0.5, 180.0 / (double)I));
    const double cell_az = __dadd_rn(-180.0, __dmul_rn((double)gj + 0.5, 360.0 / (double)J));
    double diff = __dsub_rn(cell_az, c_az);
    for (int it = 0; it < 64 && diff > 180.0; ++it) diff = __dsub_rn(diff, 360.0);
    for (int it = 0; it < 64 && diff < -180.0; ++it) diff = __dadd_rn(diff, 360.0);
    return (fabs(diff) <= two_s_az) && (el_min <= cell_el) && (cell_el <= el_max);
}

// Candidate cells of a region event: a box of grid cells that contains every cell the exact test can accept (its
// half-width is 2 sigma plus one cell of margin in each direction; azimuth wraps, elevation clamps).  The exact float64
// test still decides each candidate, so the result is that of testing all I*J cells (smrl_seld_gaussian.py:485-518) —
// 9-25 evaluations instead of 648.
struct RegionBox { int i0, ni, j0, nj; };
__device__ __forceinline__ RegionBox region_box(int I, int J, double c_az, double c_el, double two_s_az, double two_s_el) {
    const double ch = 180.0 / (double)I, cw = 360.0 / (double)J;
    RegionBox b;
    int ri = (int)ceil(two_s_el / ch) + 1, rj = (int)ceil(two_s_az / cw) + 1;
    int ic = (int)floor((c_el + 90.0) / ch), jc = (int)floor((c_az + 180.0) / cw);
    if (!(two_s_el >= 0.0) || !(two_s_az >= 0.0) || !(fabs(c_el) < 1e6) || !(fabs(c_az) < 1e6) || ri > I || rj > J) {
        b.i0 = 0; b.ni = I; b.j0 = 0; b.nj = J;  // degenerate inputs: test every cell, like the reference
        return b;
    }
    int i0 = ic - ri, i1 = ic + ri;
    i0 = i0 < 0 ? 0 : i0;
    i1 = i1 > I - 1 ? I - 1 : i1;
    b.i0 = i0;
    b.ni = i1 >= i0 ? i1 - i0 + 1 : 0;
    if (2 * rj + 1 >= J) { b.j0 = 0; b.nj = J; }
    else { b.j0 = ((jc - rj) % J + J) % J; b.nj = 2 * rj + 1; }
    return b;
}

// one CTA per event; pass 0: background <- 0, pass 1: class <- 1
__global__ void labels_paint_kernel(float* __restrict__ out, long long rows, int I, int J, int M,
                                    const int4* __restrict__ events, const double2* __restrict__ centres,
                                    double two_s_az, double two_s_el, int pass) {
    const int4 ev = events[blockIdx.x];  // {row0, row1, cls, cell}
    // a malformed table must not write outside the label tensor: rows are clamped, events with a class or cell out
    // of range (or a region event without centres) are skipped
    const long long row0 = ev.x < 0 ? 0 : ev.x, row1 = ev.y > rows ? rows : ev.y;
    if (row1 <= row0) return;
    const int cells = I * J;
    if (ev.z < 0 || ev.z >= M || ev.w >= cells || (ev.w < 0 && centres == nullptr)) return;
    const int col = pass == 0 ? M - 1 : ev.z;
    const float val = pass == 0 ? 0.f : 1.f;
    if (ev.w >= 0) {
        for (long long r = row0 + threadIdx.x; r < row1; r += blockDim.x)
            out[(r * cells + ev.w) * M + col] = val;
        return;
    }
    const double2 c = centres[blockIdx.x];
    const RegionBox bx = region_box(I, J, c.x, c.y, two_s_az, two_s_el);
    for (int q = threadIdx.x; q < bx.ni * bx.nj; q += blockDim.x) {
        const int gi = bx.i0 + q / bx.nj, gj = (bx.j0 + q % bx.nj) % J;
        if (!cell_in_region(gi, gj, I, J, c.x, c.y, two_s_az, two_s_el)) continue;
        const int cell = gi * J + gj;
        for (long long r = row0; r < row1; ++r) out[(r * cells + cell) * M + col] = val;
    }
}

// ---- window gather -------------------------------------------------------------------------------
template <typename V>
__global__ void window_gather_kernel(const V* __restrict__ src, long long rows, long long row_len,
                                     const long long* __restrict__ starts, int win_len,
                                     const V* __restrict__ pad_row, V* __restrict__ out, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long j = i % row_len;
        const long long wf = i / row_len;
        const long long f = wf % win_len;
        const long long w = wf / win_len;
        const long long r = starts[w] + f;
        out[i] = (r >= 0 && r < rows) ? src[r * row_len + j] : pad_row[j];
    }
}

template <typename V>
__global__ void scaler_apply_kernel(V* __restrict__ x, long long total, int n_feat_v, const V* __restrict__ mean,
                                    const V* __restrict__ inv_std);

template <>
__global__ void scaler_apply_kernel<float4>(float4* __restrict__ x, long long total, int n_feat_v,
                                            const float4* __restrict__ mean, const float4* __restrict__ inv_std) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int f = int(i % n_feat_v);
        float4 v = x[i];
        const float4 m = __ldg(mean + f), s = __ldg(inv_std + f);
        v.x = (v.x - m.x) * s.x; v.y = (v.y - m.y) * s.y; v.z = (v.z - m.z) * s.z; v.w = (v.w - m.w) * s.w;
        x[i] = v;
    }
}

template <>
__global__ void scaler_apply_kernel<float>(float* __restrict__ x, long long total, int n_feat_v,
                                           const float* __restrict__ mean, const float* __restrict__ inv_std) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int f = int(i % n_feat_v);
        x[i] = (x[i] - __ldg(mean + f)) * __ldg(inv_std + f);
    }
}

static int grid_for(long long work, int block, int device_sms) {
    long long g = (work + block - 1) / block;
    long long cap = (long long)device_sms * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static int current_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_labels_fill(float* out, long long rows, int cells, int M, cudaStream_t st) {
    const long long n = rows * cells * M;
    if (n == 0) return SELD_OK;
    const int sms = current_sms();
    if (aligned16(out) && n % 4 == 0) {
        labels_fill_vec4<<<grid_for(n / 4, 256, sms), 256, 0, st>>>(reinterpret_cast<float4*>(out), n / 4, M);
    } else {
        labels_fill_scalar<<<grid_for(n, 256, sms), 256, 0, st>>>(out, n, M);
    }
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

int launch_labels_paint(float* out, long long rows, int I, int J, int M, const int* events, const double* centres,
                        int n_events, double sigma_az, double sigma_el, cudaStream_t st) {
    if (n_events == 0) return SELD_OK;
    const double two_az = 2 * sigma_az, two_el = 2 * sigma_el;  // the reference's `2 * sigma_*`
    for (int pass = 0; pass < 2; ++pass) {
        labels_paint_kernel<<<n_events, 128, 0, st>>>(out, rows, I, J, M, reinterpret_cast<const int4*>(events),
                                                       reinterpret_cast<const double2*>(centres), two_az, two_el,
                                                       pass);
        SELD_CUDA_TRY(cudaGetLastError());
    }
    return SELD_OK;
}

int launch_window_gather(const float* src, long long rows, long long row_len, const long long* starts, int n_win,
                         int win_len, const float* pad_row, float* out, cudaStream_t st) {
    const long long total = (long long)n_win * win_len * row_len;
    if (total == 0) return SELD_OK;
    const int sms = current_sms();
    if (row_len % 4 == 0 && aligned16(src) && aligned16(out) && aligned16(pad_row)) {
        window_gather_kernel<float4><<<grid_for(total / 4, 256, sms), 256, 0, st>>>(
            reinterpret_cast<const float4*>(src), rows, row_len / 4, starts, win_len,
            reinterpret_cast<const float4*>(pad_row), reinterpret_cast<float4*>(out), total / 4);
    } else {
        window_gather_kernel<float><<<grid_for(total, 256, sms), 256, 0, st>>>(src, rows, row_len, starts, win_len,
                                                                                 pad_row, out, total);
    }
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

int launch_scaler_apply(float* x, long long rows, int n_feat, const float* mean, const float* inv_std,
                        cudaStream_t st) {
    const long long total = rows * n_feat;
    if (total == 0) return SELD_OK;
    const int sms = current_sms();
    if (n_feat % 4 == 0 && aligned16(x) && aligned16(mean) && aligned16(inv_std)) {
        scaler_apply_kernel<float4><<<grid_for(total / 4, 256, sms), 256, 0, st>>>(
            reinterpret_cast<float4*>(x), total / 4, n_feat / 4, reinterpret_cast<const float4*>(mean),
            reinterpret_cast<const float4*>(inv_std));
    } else {
        scaler_apply_kernel<float><<<grid_for(total, 256, sms), 256, 0, st>>>(x, total, n_feat, mean, inv_std);
    }
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

// ---- PCM16 ingest: int16 samples -> float32 in [-1, 1), x / 32768 (what torchaudio.load returns for 16-bit WAV, the
// input of reference dataset.py:18-25 load_audio).  Uploading int16 and converting here halves the PCIe bytes. ----
__global__ void pcm16_to_float_kernel(const short* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n8 = n / 8;
    const uint4* in8 = reinterpret_cast<const uint4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    constexpr float k = 1.0f / 32768.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 v = __ldg(in8 + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = k * (float)(short)(w[j] & 0xffffu);
            f[2 * j + 1] = k * (float)(short)(w[j] >> 16);
        }
        __stcs(out4 + 2 * i, make_float4(f[0], f[1], f[2], f[3]));
        __stcs(out4 + 2 * i + 1, make_float4(f[4], f[5], f[6], f[7]));
    }
    for (long long i = 8 * n8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = k * (float)in[i];
}

int launch_pcm16_to_float(const short* in, float* out, long long n, cudaStream_t st) {
    if (n == 0) return SELD_OK;
    if (!aligned16(in) || !aligned16(out)) {
        set_error("seld_pcm16_to_float: buffers must be 16-byte aligned");
        return SELD_ERR_BAD_ARG;
    }
    pcm16_to_float_kernel<<<grid_for((n + 7) / 8, 256, current_sms()), 256, 0, st>>>(in, out, n);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

}  // namespace seld

// ---- on-device batch assembly (replaces reference dataset.py:267-330 _create_windows +
// __getitem__ + the DataLoader collate of main.py:60-74 for one training batch) ------------------------------------
// ONE launch per batch.  CTA (w, s) owns rows [s * kSliceRows, ...) of window w of the batch:
//   features: out_spec[w, f, :] = feat[start_w + f, :] (zeros past the end of the corpus, dataset.py:290-296)
//   labels  : one-hot background over its rows (dataset.py:114-117 / :297-300), __syncthreads, then the events of
//             the window (a precomputed range [lo, hi) of the table sorted by first row) clipped to its rows:
//             pass 0 clears the background class, __syncthreads, pass 1 sets the event class (labels_paint_kernel's
//             order, so a class M-1 event leaves the background at 1 like the reference).
// Nothing crosses PCIe and the host does no per-window work: the epoch's permutation, the window starts and the
// per-window event ranges are resident in HBM.
namespace seld {
#ifndef SELD_SLICE_ROWS
#define SELD_SLICE_ROWS 5
#endif
constexpr int kSliceRows = SELD_SLICE_ROWS;

__global__ void __launch_bounds__(256) loader_batch_kernel(
    const float4* __restrict__ feat, long long rows, int row_len4, const int* __restrict__ order, int first, int n_win,
    const int* __restrict__ win_start, const int* __restrict__ win_lo, const int* __restrict__ win_hi, int win_len,
    float4* __restrict__ out_spec, const int4* __restrict__ events, const double2* __restrict__ centres, int I, int J, int M,
    double two_s_az, double two_s_el, float* __restrict__ out_lab) {
    const int w = blockIdx.y, s = blockIdx.x;
    const int widx = order ? order[first + w] : first + w;
    const long long start = win_start[widx];
    const int f0 = s * kSliceRows, f1 = min(win_len, f0 + kSliceRows);
    if (f0 >= f1) return;
    const int cells = I * J;
    // features
    {
        const int n4 = (f1 - f0) * row_len4;
        float4* dst = out_spec + ((long long)w * win_len + f0) * row_len4;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const int f = i / row_len4, j = i - f * row_len4;
            const long long r = start + f0 + f;
            dst[i] = r < rows ? __ldg(feat + r * row_len4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (!out_lab) return;
    // labels: background fill of this slice (the slice starts at a multiple of cells * M floats, so the period-M
    // pattern starts at phase 0; 16-byte stores need cells * M % 4 == 0, checked by the launcher)
    float* lab = out_lab + ((long long)w * win_len + f0) * cells * M;
    {
        const long long n_vec = (long long)(f1 - f0) * cells * M / 4;
        float4* lab4 = reinterpret_cast<float4*>(lab);
        int r = int((4ll * threadIdx.x) % M);
        const int step = int((4ll * blockDim.x) % M);
        for (long long i = threadIdx.x; i < n_vec; i += blockDim.x) {
            int r1 = r + 1; if (r1 >= M) r1 -= M;
            int r2 = r1 + 1; if (r2 >= M) r2 -= M;
            int r3 = r2 + 1; if (r3 >= M) r3 -= M;
            __stcs(lab4 + i, make_float4(r == M - 1 ? 1.f : 0.f, r1 == M - 1 ? 1.f : 0.f, r2 == M - 1 ? 1.f : 0.f,
                                         r3 == M - 1 ? 1.f : 0.f));
            r += step; if (r >= M) r -= M;
        }
    }
    const int lo = win_lo[widx], hi = win_hi[widx];
    const long long g0 = start + f0, g1 = start + f1;  // absolute rows of this slice
    // events of the window that touch this slice: every thread tests one event of the window's range (one coalesced
    // load instead of a serial scan) and appends the hits to a short list; chunks of blockDim.x events
    __shared__ int s_hits[256];
    __shared__ int s_n;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)  // pass 0 clears the background class (all events), pass 1 sets the event class
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
            const int n = s_n;
            for (int h = 0; h < n; ++h) {
                const int e2 = s_hits[h];
                const int4 ev = events[e2];  // {row0, row1, cls, cell}, absolute rows
                const long long r0 = max((long long)ev.x, g0), r1 = min((long long)ev.y, g1);
                const int col = pass == 0 ? M - 1 : ev.z;
                const float val = pass == 0 ? 0.f : 1.f;
                if (ev.w >= 0) {
                    for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x)
                        lab[((r - g0) * cells + ev.w) * M + col] = val;
                } else if (centres) {
                    const double2 c = centres[e2];
                    const RegionBox bx = region_box(I, J, c.x, c.y, two_s_az, two_s_el);
                    for (int q = threadIdx.x; q < bx.ni * bx.nj; q += blockDim.x) {
                        const int gi = bx.i0 + q / bx.nj, gj = (bx.j0 + q % bx.nj) % J;
                        if (!cell_in_region(gi, gj, I, J, c.x, c.y, two_s_az, two_s_el)) continue;
                        const int cell = gi * J + gj;
                        for (long long r = r0; r < r1; ++r) lab[((r - g0) * cells + cell) * M + col] = val;
                    }
                }
            }
        }
}

int launch_loader_batch(const float* feat, long long rows, int row_len, const int* order, int first, int n_win,
                        const int* win_start, const int* win_lo, const int* win_hi, int win_len, float* out_spec,
                        const int* events, const double* centres, int I, int J, int M, double sigma_az, double sigma_el,
                        float* out_lab, cudaStream_t st) {
    if (n_win == 0 || win_len == 0) return SELD_OK;
    if (row_len % 4 != 0 || !aligned16(feat) || !aligned16(out_spec)) {
        set_error("seld_loader_batch: feature rows must be 16-byte aligned multiples of 4 floats");
        return SELD_ERR_BAD_ARG;
    }
    if (out_lab && (((long long)I * J * M) % 4 != 0 || !aligned16(out_lab))) {
        set_error("seld_loader_batch: I * J * n_classes must be a multiple of 4 and the label buffer 16-byte aligned");
        return SELD_ERR_BAD_ARG;
    }
    dim3 grid((unsigned)((win_len + kSliceRows - 1) / kSliceRows), (unsigned)n_win);
    loader_batch_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(feat), rows, row_len / 4, order, first, n_win,
                                              win_start, win_lo, win_hi, win_len, reinterpret_cast<float4*>(out_spec),
                                              reinterpret_cast<const int4*>(events), reinterpret_cast<const double2*>(centres),
                                              I, J, M, 2 * sigma_az, 2 * sigma_el, out_lab);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}
}  // namespace seld
