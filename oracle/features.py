"""fp64 numpy restatement of the reference feature path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows, step by step:

* ``audio_to_mel_spectrogram``                       reference dataset.py:27-58
* ``torchaudio.transforms.MelSpectrogram`` defaults  torchaudio/transforms/_transforms.py:566-631
  (win_length = n_fft, periodic Hann, center=True, pad_mode="reflect", power=2, onesided,
  normalized=False, f_min=0, f_max=sr//2, norm=None, mel_scale="htk")
* ``F.spectrogram`` → ``torch.stft``                 torchaudio/functional/functional.py:116-145
* ``melscale_fbanks``                                torchaudio/functional/functional.py:492-587
* ``AmplitudeToDB()`` (power, amin 1e-10, ref 1, no top_db)  torchaudio/functional/functional.py:390-391

All arithmetic is float64 except the filterbank, which torchaudio builds in float32
(``torch.linspace`` float32) — the oracle reproduces that float32 construction so that the weights are
the reference's weights, then promotes them to float64.

IV / GCC-PHAT / scaler have no reference code (parity unpinned).
"""
from __future__ import annotations

import math

import numpy as np

EPS_IV = 1e-8
AMIN = 1e-10  # torchaudio AmplitudeToDB default amin


def num_frames(n_samples: int, hop: int) -> int:
    """torch.stft(center=True): T = 1 + N // hop."""
    return 1 + n_samples // hop


def hann_periodic(n_fft: int) -> np.ndarray:
    """The reference's window: torch.hann_window(n_fft) — periodic Hann 0.5 - 0.5 cos(2 pi n / N) built by
    ATen in FLOAT32 (arange * float(2pi/N) -> cos -> *(-0.5) + 0.5).  Near the edges the float32
    cancellation leaves up to ~1 % relative error in w[n] (w[N-1] ~ 1e-5), which is visible in dB for a
    signal that lives only there (golden case impulse_last: 8e-3 dB vs an exact-Hann fp64 DFT), so the
    oracle uses the reference's float32 table promoted to float64 — same policy as the filterbank."""
    import torch

    return torch.hann_window(n_fft, dtype=torch.float32).double().numpy()


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: step=(end-start)/(steps-1); first half start+i*step,
    second half end-(steps-1-i)*step (ATen RangeFactories kernel)."""
    start32, end32 = np.float32(start), np.float32(end)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (start32 + step * i.astype(np.float32)).astype(np.float32)
    hi = (end32 - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def mel_filterbank_f32(n_freqs: int, sample_rate: int, n_mels: int) -> np.ndarray:
    """(n_freqs, n_mels) float32 HTK triangles, f_min 0, f_max sr//2, norm None.

    Restates melscale_fbanks + _create_triangular_filterbank in float32 numpy.  tests/ check it
    against torchaudio's own table (exact or within 1 ulp of the float32 pow)."""
    all_freqs = _linspace_f32(0.0, float(sample_rate // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + float(sample_rate // 2) / 700.0)
    m_pts = _linspace_f32(m_min, m_max, n_mels + 2)
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0), dtype=np.float32)
                                  - np.float32(1.0))).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (np.float32(-1.0) * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


def stft(audio: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """(C, N) -> complex128 (C, T, n_fft//2+1); reflect pad n_fft//2, periodic Hann, rfft."""
    x = np.asarray(audio, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("audio must be (C, N)")
    pad = n_fft // 2
    if x.shape[1] <= pad:
        raise ValueError("reflect padding needs N > n_fft//2")
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    T = num_frames(x.shape[1], hop)
    idx = (np.arange(T) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[:, idx] * hann_periodic(n_fft)[None, None, :]
    return np.fft.rfft(frames, n=n_fft, axis=-1)


def power_to_db(p: np.ndarray) -> np.ndarray:
    return 10.0 * np.log10(np.maximum(p, AMIN))


def logmel(audio: np.ndarray, sample_rate: int, n_fft: int, hop: int, n_mels: int,
           fb: np.ndarray | None = None) -> np.ndarray:
    """Reference layout (C, n_mels, T), float64 dB."""
    if fb is None:
        fb = mel_filterbank_f32(n_fft // 2 + 1, sample_rate, n_mels)
    X = stft(audio, n_fft, hop)
    P = X.real ** 2 + X.imag ** 2  # (C, T, F)
    mel = P @ fb.astype(np.float64)  # (C, T, M)
    return power_to_db(mel).transpose(0, 2, 1)


def foa_iv(audio: np.ndarray, sample_rate: int, n_fft: int, hop: int, n_mels: int,
           fb: np.ndarray | None = None) -> np.ndarray:
    """FOA intensity vectors (3, n_mels, T).  NOT IN REFERENCE — parity unpinned.

    I_k = Re(conj(W) X_k), k=1..3 (file channel order, ch0 = W);
    E = eps + |W|^2 + (sum_k |X_k|^2)/3;  mel-project I_k/E with the log-mel filterbank; no log."""
    if audio.shape[0] != 4:
        raise ValueError("FOA intensity vectors need 4 channels")
    if fb is None:
        fb = mel_filterbank_f32(n_fft // 2 + 1, sample_rate, n_mels)
    X = stft(audio, n_fft, hop)
    W = X[0]
    I = np.real(np.conj(W)[None] * X[1:])  # (3, T, F)
    E = EPS_IV + np.abs(W) ** 2 + (np.abs(X[1:]) ** 2).sum(0) / 3.0
    return ((I / E[None]) @ fb.astype(np.float64)).transpose(0, 2, 1)


def logmel_iv(audio, sample_rate, n_fft, hop, n_mels, fb=None) -> np.ndarray:
    """7-channel FOA feature (7, n_mels, T): 4 log-mel then 3 IV."""
    return np.concatenate([logmel(audio, sample_rate, n_fft, hop, n_mels, fb),
                           foa_iv(audio, sample_rate, n_fft, hop, n_mels, fb)], axis=0)


def gcc_phat(audio: np.ndarray, n_fft: int, hop: int, n_lags: int = 64) -> np.ndarray:
    """GCC-PHAT (n_pairs, n_lags, T).  NOT IN REFERENCE — parity unpinned.

    Pairs (m<n) in order 01,02,03,12,13,23; R = conj(X_m) X_n; R/|R| (R == 0 -> 1, i.e. exp(j*angle(0)));
    cc = irfft(., n_fft); keep lags [-n_lags/2 .. -1, 0 .. n_lags/2-1]."""
    X = stft(audio, n_fft, hop)
    C = X.shape[0]
    out = []
    for m in range(C):
        for n in range(m + 1, C):
            R = np.conj(X[m]) * X[n]
            mag = np.abs(R)
            ph = np.where(mag > 0, R / np.where(mag > 0, mag, 1.0), 1.0 + 0.0j)
            cc = np.fft.irfft(ph, n=n_fft, axis=-1)  # (T, n_fft)
            cc = np.concatenate([cc[:, -(n_lags // 2):], cc[:, : n_lags // 2]], axis=-1)
            out.append(cc.T)
    return np.stack(out, axis=0)


def mic_features(audio, sample_rate, n_fft, hop, n_mels, fb=None) -> np.ndarray:
    """10-channel MIC feature (10, n_mels, T): 4 log-mel then 6 GCC-PHAT (n_lags == n_mels)."""
    return np.concatenate([logmel(audio, sample_rate, n_fft, hop, n_mels, fb),
                           gcc_phat(audio, n_fft, hop, n_mels)], axis=0)


def scaler_stats(features_tcm: np.ndarray) -> tuple[float, np.ndarray, np.ndarray]:
    """(T, C, M) -> (count, sum[C*M], sumsq[C*M]) in float64.  NOT IN REFERENCE — parity unpinned."""
    x = np.asarray(features_tcm, dtype=np.float64).reshape(features_tcm.shape[0], -1)
    return float(x.shape[0]), x.sum(0), (x * x).sum(0)


def scaler_mean_std(count, s, ss):
    mean = s / count
    var = np.maximum(ss / count - mean * mean, 0.0)
    return mean, np.sqrt(var)
