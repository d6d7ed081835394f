/* libseld_cuda — C ABI of the B200-native SELD feature front-end (hand-written sm_100a kernels).
 *
 * The reference (Zeudon/sound-event-localization-detection) has no FFI: its boundary for this path is the
 * set of plain Python callables in dataset.py / smrl_seld_gaussian.py.  Each entry point below names the
 * reference function it replaces.  The Python host layer (sound-event-localization-detection_b200/*.py)
 * binds these with ctypes and re-exposes the reference's names and signatures; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - All d_* pointers are DEVICE pointers owned by the caller (PyTorch allocates them); h_* are host
 *    pointers.  The library never allocates or frees caller memory; a plan owns only its constant tables.
 *  - Every compute call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*;
 *    NULL = legacy default stream).  No hidden synchronisation, no host allocation on the hot calls.
 *  - Return value: 0 = OK, negative = seld_status.  Nothing aborts or throws across the boundary;
 *    seld_last_error() returns a thread-local message for the last failing call on this thread.
 *  - There is no CPU fallback: without a CUDA device every compute call returns SELD_ERR_CUDA.
 *  - Every entry point runs on the device of its plan (or of its output pointer) and RESTORES the caller's current
 *    device before returning: a process that drives several GPUs can call from any thread with any current device.
 *  - The hot calls do no host work besides argument checks and the launch: environment switches (SELD_FEAT_IMPL,
 *    SELD_V3_CFG) are read and kernel attributes are set once, in seld_plan_create.
 */
#ifndef SELD_CUDA_H
#define SELD_CUDA_H

#include <stdint.h>

#if defined(_WIN32)
#define SELD_API
#else
#define SELD_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum seld_status {
    SELD_OK = 0,
    SELD_ERR_BAD_ARG = -1,
    SELD_ERR_UNSUPPORTED = -2, /* e.g. n_fft other than 960 / 1024, n_mels > 64 */
    SELD_ERR_CUDA = -3,
    SELD_ERR_ALLOC = -4
} seld_status;

typedef struct seld_plan seld_plan;

/* feature sets (seld_features `mode`) */
#define SELD_MODE_LOGMEL 0     /* C log-mel channels             (reference dataset.py:27-58)            */
#define SELD_MODE_LOGMEL_IV 1  /* 4 log-mel + 3 FOA intensity     (north_star kernel 2; not in reference) */
#define SELD_MODE_LOGMEL_GCC 2 /* 4 log-mel + 6 GCC-PHAT x n_mels (north_star kernel 3; not in reference) */

SELD_API int seld_version(void);
SELD_API const char* seld_last_error(void);

/* Number of STFT frames torch.stft(center=True) produces: 1 + n_samples / hop
 * (torchaudio/functional/functional.py:123-134 via reference dataset.py:49). */
SELD_API int64_t seld_num_frames(int64_t n_samples, int hop);

/* Output channels of a mode for C input channels (C, 7, 10). */
SELD_API int seld_out_channels(int mode, int n_channels);

/* Build the constant tables of one (n_fft, hop, n_mels) configuration on `device`.
 * Replaces the per-call construction in reference dataset.py:38-43
 * (torchaudio.transforms.MelSpectrogram: Hann window + HTK filterbank).
 *   h_window : n_fft floats, the analysis window (torch.hann_window(n_fft), float32)
 *   h_fb     : (n_fft/2+1, n_mels) row-major float32 filterbank (melscale_fbanks(...))
 * n_fft in {960, 1024}; n_mels <= 64.  GCC-PHAT mode: n_mels == 64 (= lags). */
SELD_API int seld_plan_create(seld_plan** plan, int device, int n_fft, int hop, int n_mels, const float* h_window,
                     const float* h_fb);
SELD_API int seld_plan_destroy(seld_plan* plan);

/* Fused framing + window + real FFT + power + mel + 10*log10 (+ IV | + GCC-PHAT).
 * Replaces reference dataset.py:27-58 audio_to_mel_spectrogram for a whole batch of clips.
 *   d_audio      : float32, clip b channel c sample n at d_audio[b*clip_stride + c*chan_stride + n]
 *   d_lengths    : int64[B] valid samples per clip, or NULL (all = n_samples); each must be > n_fft/2 (a shorter
 *                  clip is flagged on the device, see seld_plan_status, and its rows are written as 0)
 *   d_out        : float32 (B, T_out, C_out, n_mels), frame-major ("TCM") — the layout the reference's
 *                  models consume after dataset.py:303; rows t >= 1 + len_b/hop are written as 0
 *   T_out        : frame capacity per clip (>= seld_num_frames(max length))
 *   C_out        : channel count of d_out (>= c_off + seld_out_channels(mode, C))
 *   c_off        : first output channel to write
 *   d_stats      : optional float64[2 * C_out * n_mels]: the call ADDS per-feature sum and sum of squares
 *                  of the values it writes for frames t < d_stat_frames[b] (or all valid frames when
 *                  d_stat_frames is NULL) — the normalisation-scaler partials.  On the fast
 *                  path the sums are formed inside the feature kernel (fp32 partials of <= 64 rows, flushed into
 *                  the float64 sums: relative error ~1e-7); otherwise by a second kernel over d_out.
 *   d_spec       : optional float32 complex (B, C, T_out, n_fft/2+1) dump of the STFT (parity tests)
 */
SELD_API int seld_features(seld_plan* plan, int mode, const float* d_audio, int64_t clip_stride, int64_t chan_stride,
                  int64_t n_samples, const int64_t* d_lengths, int B, int C, float* d_out, int64_t T_out,
                  int C_out, int c_off, double* d_stats, const int32_t* d_stat_frames, float* d_spec,
                  void* stream);

/* seld_features with input / output options (fast path only: 4 channels, the reference's 64-mel HTK bank, no spectrum
 * dump, 16-byte aligned output; SELD_ERR_UNSUPPORTED otherwise).
 *   in_dtype   : SELD_DTYPE_F32, or SELD_DTYPE_I16 — d_audio holds 16-bit PCM and the kernel converts x / 32768 while
 *                loading (what torchaudio.load returns for a 16-bit WAV, the decode half of reference dataset.py:18-25
 *                load_audio): half the PCIe / HBM bytes per sample, no separate conversion pass
 *   out_layout : SELD_LAYOUT_TCF (B, T, C, F) as seld_features, or SELD_LAYOUT_CTF (B, C, T, F) — the layout every
 *                reference model permutes to first (model_crnn.py:106, model_conformer.py:191, resnet50_model.py:179)
 *   out_dtype  : SELD_DTYPE_F32 or SELD_DTYPE_BF16
 *   d_mean, d_inv_std : float32[C_out * n_mels] or both NULL; when given the kernel writes (x - mean) * inv_std of the
 *                valid rows (rows past a clip's end stay 0) — the scaler apply step fused into the row copy-out.
 * d_stats cannot be combined with normalisation / layout / dtype options (the partials are those of the raw features).
 * opts == NULL is seld_features.  d_out is float32 or bfloat16 according to out_dtype. */
#define SELD_DTYPE_F32 0
#define SELD_DTYPE_I16 1
#define SELD_DTYPE_BF16 2
#define SELD_LAYOUT_TCF 0
#define SELD_LAYOUT_CTF 1
typedef struct seld_feat_opts {
    int in_dtype;
    int out_layout;
    int out_dtype;
    const float* d_mean;
    const float* d_inv_std;
} seld_feat_opts;
SELD_API int seld_features_ex(seld_plan* plan, int mode, const void* d_audio, int64_t clip_stride, int64_t chan_stride,
                     int64_t n_samples, const int64_t* d_lengths, int B, int C, void* d_out, int64_t T_out,
                     int C_out, int c_off, double* d_stats, const int32_t* d_stat_frames, float* d_spec,
                     const seld_feat_opts* opts, void* stream);

/* 1 if the plan's filterbank is the reference's (64 HTK mels of a 24 kHz / n_fft 960|1024 STFT, bit for bit) so that
 * 4-channel calls run on the fast kernel — the precondition of seld_features_ex options; 0 otherwise. */
SELD_API int seld_plan_has_fast_path(const seld_plan* plan);

/* Device-side status of the calls issued so far on `stream` with this plan: synchronises the stream, copies the
 * status word back, clears it.  *h_status bit 0: some clip of a ragged batch (d_lengths) had <= n_fft/2 samples —
 * torch.stft(center=True, pad_mode="reflect") raises for those (reference dataset.py:49); the kernels read nothing
 * from such a clip and write its rows as 0.  Returns SELD_ERR_BAD_ARG when a bit is set, SELD_OK otherwise. */
SELD_API int seld_plan_status(seld_plan* plan, void* stream, int* h_status);

/* Scaler partials of an already computed feature tensor (the accumulation seld_features performs when given
 * d_stats, as a call of its own): for the channels [c_off, c_off + n_channels) of d_feat (B, T_out, C_out, n_mels)
 * ADD per-feature sum and sum of squares over the frames t < d_stat_frames[b] (NULL: t < 1 + len_b / hop with
 * len_b = d_lengths[b], or n_samples when d_lengths is NULL) into d_stats (float64[2 * C_out * n_mels]).
 * Not in the reference. */
SELD_API int seld_feature_stats(seld_plan* plan, const float* d_feat, int B, int64_t T_out, int C_out, int c_off,
                       int n_channels, int64_t n_samples, const int64_t* d_lengths, const int32_t* d_stat_frames,
                       double* d_stats, void* stream);

/* (x - mean) * inv_std in place over (rows, n_feat) float32; mean/inv_std float32[n_feat].
 * The scaler apply step. */
SELD_API int seld_scaler_apply(float* d_x, int64_t rows, int n_feat, const float* d_mean, const float* d_inv_std,
                      void* stream);

/* PCM16 ingest: d_out[i] = d_pcm[i] / 32768 (float32), the values torchaudio.load returns for a 16-bit WAV —
 * the decode half of reference dataset.py:18-25 load_audio, done on the device so that only 2 bytes per sample
 * cross PCIe.  Both buffers 16-byte aligned. */
SELD_API int seld_pcm16_to_float(const int16_t* d_pcm, float* d_out, int64_t n, void* stream);

/* Dense SELD grid labels (T, cells, classes) float32 for a batch of label tensors.
 * Replaces reference dataset.py:60-119 metadata_to_labels and smrl_seld_gaussian.py:397-534
 * augment_with_gaussian_noise.  The host parses the CSV exactly like the reference (pandas + int()) and
 * hands over compact event tables; the kernels do the dense encoding.
 *
 * seld_labels_fill   : every (t, cell) <- one-hot background (class n_classes-1), i.e. the state the
 *                      reference reaches for cells with no event (dataset.py:114-117); rows total.
 * seld_labels_paint  : n_events events; event e covers rows [row0[e], row1[e]) of d_out (absolute row
 *                      index = frame index inside the concatenated (rows, cells, classes) tensor) and
 *                      class cls[e] (already wrapped to [0, n_classes)).
 *                      point events : cell[e] >= 0 is the single grid cell (utils.py:77-90 polar_to_grid)
 *                      region events: cell[e] < 0; centre[e] = (azimuth + noise, elevation + noise) in
 *                      float64 and every cell whose centre passes the reference's +-2 sigma test
 *                      (smrl_seld_gaussian.py:474-518, same float64 operations and <= comparisons) is painted.
 *                      Painting sets (t, cell, cls) = 1 and clears the background class of that cell unless
 *                      some event of the cell has cls == n_classes-1 (then it stays 1, as in the reference).
 *   d_events : int32 (n_events, 4) rows {row0, row1, cls, cell}
 *   d_centres: float64 (n_events, 2) or NULL when there are no region events
 */
SELD_API int seld_labels_fill(float* d_out, int64_t rows, int cells, int n_classes, void* stream);
SELD_API int seld_labels_paint(float* d_out, int64_t rows, int I, int J, int n_classes, const int32_t* d_events,
                      const double* d_centres, int n_events, double sigma_az, double sigma_el, void* stream);

/* Window/batch assembly (reference dataset.py:267-330 _create_windows + __getitem__, for a batch):
 * out[w, f, :] = src[start[w] + f, :] for start[w] + f < rows, else `pad_row` (row_len floats; zeros for
 * features, one-hot background for labels).  src (rows, row_len), out (n_win, win_len, row_len). */
SELD_API int seld_window_gather(const float* d_src, int64_t rows, int64_t row_len, const int64_t* d_starts, int n_win,
                       int win_len, const float* d_pad_row, float* d_out, void* stream);

/* One training batch assembled on the device in ONE launch: replaces reference dataset.py:267-330
 * (_create_windows + __getitem__) plus the DataLoader collate and the 145 MB host->device label copy of
 * trainer.py:167-168.  Everything it reads is resident in HBM; the host passes only the position in the epoch.
 *   d_feat       : (rows, row_len) float32 concatenated features, frame-major (row_len = C * n_mels)
 *   d_order      : int32 permutation of the dataset's window indices for this epoch, or NULL (identity)
 *   first, n_win : the batch is windows d_order[first .. first + n_win)
 *   d_win_start  : int32[n_windows] first frame of every window (50 k)
 *   d_win_lo/hi  : int32[n_windows] range of the event table (sorted by first row) that can touch the window
 *   d_spec_out   : (n_win, win_len, row_len) float32; frames past the end of the corpus are 0 (dataset.py:290-296)
 *   d_events     : int32 (n_events, 4) {row0, row1, cls, cell} with absolute rows; d_centres: float64 (n_events, 2)
 *                  for region events (cell < 0), as seld_labels_paint
 *   d_labels_out : (n_win, win_len, I * J, n_classes) float32 dense targets, or NULL (features only)           */
SELD_API int seld_loader_batch(const float* d_feat, int64_t rows, int row_len, const int32_t* d_order, int first, int n_win,
                      const int32_t* d_win_start, const int32_t* d_win_lo, const int32_t* d_win_hi, int win_len,
                      float* d_spec_out, const int32_t* d_events, const double* d_centres, int I, int J, int n_classes,
                      double sigma_az, double sigma_el, float* d_labels_out, void* stream);

/* Class loss from compact targets (reference loss.py:27-54 SMRSELDLoss.class_mse_loss /
 * class_ce_loss, which take the dense (B, T, I*J, n_classes) float32 targets — 145 MB per batch, > 99.7 % background).
 *
 * seld_batch_class_mask: the targets of one batch as a class-SET mask per (window, frame, cell), uint16 (n_win, win_len,
 *   I*J): 0 = no event (one-hot background), bit c = an event of class c covers the cell.  Same arguments and event
 *   selection as seld_loader_batch; 5 MB instead of 145 MB.
 * seld_class_loss: softmax-MSE (SELD_LOSS_MSE: sum over cells and classes of (softmax(z) - y)^2) or cross entropy
 *   (SELD_LOSS_CE: sum of w[t] * (logsumexp(z) - z[t]) with t = lowest class of the mask, torch.argmax's choice for a
 *   multi-hot row, and the sum of w[t]) of d_logits (n_cells, n_classes) float32.
 *   d_sums      : float64[2], the call ADDS {loss sum, weight sum (CE)}; NULL = backward only
 *   d_grad      : NULL, or (n_cells, n_classes) float32 <- d(loss sum)/d(logits) * (*d_grad_scale); the scale is a DEVICE
 *                 float (upstream gradient x the mean's 1/N), so autograd never synchronises
 *   d_class_weight : float32[n_classes] or NULL (CE only).   n_classes == 14 (the reference's NUM_CLASSES).            */
#define SELD_LOSS_MSE 0
#define SELD_LOSS_CE 1
SELD_API int seld_batch_class_mask(const int32_t* d_order, int first, int n_win, const int32_t* d_win_start,
                          const int32_t* d_win_lo, const int32_t* d_win_hi, int win_len, const int32_t* d_events,
                          const double* d_centres, int I, int J, int n_classes, double sigma_az, double sigma_el,
                          uint16_t* d_mask, void* stream);
SELD_API int seld_class_loss(int loss_type, const float* d_logits, const uint16_t* d_mask, int64_t n_cells, int n_classes,
                    const float* d_class_weight, double* d_sums, float* d_grad, const float* d_grad_scale, void* stream);

/* The reference loss's two dormant terms from the same compact targets (reference loss.py:56-88 aiur_loss, :90-146
 * converging_localization_loss; both are commented out of SMRSELDLoss.forward, loss.py:158-165, and take softmax
 * probabilities there — here the softmax of d_logits (n_frames, I*J, n_classes) is taken inside).
 *   d_sums : float64[3], the call ADDS {sum over event frames and cells of p_nonbackground * y_at,  number of frames with
 *            events,  sum over frames of the IoU of predicted and true event cells};
 *            cl = sums[0] / (sums[1] * I * J + 1e-10), aiur = 1 - sums[2] / n_frames.  NULL = backward only.
 *   d_grad : NULL, or (n_frames, I*J, n_classes) float32 <- d(sums[0])/d(logits) * (*d_grad_scale) (device scalar); the AIUR
 *            term is an argmax statistic and has no gradient.      n_classes == 14, I*J <= 4096.                       */
SELD_API int seld_aux_losses(const float* d_logits, const uint16_t* d_mask, int64_t n_frames, int I, int J, int n_classes,
                    double* d_sums, float* d_grad, const float* d_grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELD_CUDA_H */
