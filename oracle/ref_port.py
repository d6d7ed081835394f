"""fp32 torch/torchaudio port of the reference CPU feature path (TEST INFRASTRUCTURE — oracle/__init__.py).

This is the CPU arm that ``bench.py`` times (``cpu_baseline`` and ``--impl reference``): the same
torchaudio call sequence as reference dataset.py:38-56 (``MelSpectrogram`` per channel, ``torch.cat``,
``AmplitudeToDB``), running on the host cores through torch's MKL/pocketfft ``torch.stft``.  The
reference tree itself cannot travel to the GPU box, so this port is what runs there; in the build
container it is checked bit-for-bit against the real ``dataset.audio_to_mel_spectrogram``
(tests/golden/make_golden.py, tests/test_oracle.py).

The FOA-IV part has no reference code; ``logmel_iv_port`` adds it with the same torch primitives
(``torch.stft`` complex output) so the 7-channel CPU baseline does the same work as the GPU kernel.
"""
from __future__ import annotations

import torch


def audio_to_mel_spectrogram_port(waveform: torch.Tensor, sample_rate: int, n_fft: int, hop_length: int,
                                  n_mels: int) -> torch.Tensor:
    """dataset.py:27-58, same calls in the same order.  (C, N) fp32 -> (C, n_mels, T) fp32 dB."""
    import torchaudio

    mel_transform = torchaudio.transforms.MelSpectrogram(
        sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)
    mel_specs = []
    for channel_idx in range(waveform.shape[0]):
        mel_specs.append(mel_transform(waveform[channel_idx:channel_idx + 1, :]))
    mel = torch.cat(mel_specs, dim=0)
    return torchaudio.transforms.AmplitudeToDB()(mel)


def _fb(n_fft, sample_rate, n_mels):
    import torchaudio

    return torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels,
                                                 sample_rate, norm=None, mel_scale="htk")


def logmel_iv_port(waveform: torch.Tensor, sample_rate: int, n_fft: int, hop_length: int, n_mels: int,
                   fb: torch.Tensor | None = None) -> torch.Tensor:
    """7-channel FOA feature (7, n_mels, T) fp32 on the CPU: one batched torch.stft, power->mel->dB for
    the 4 channels, IV from the complex spectra."""
    if fb is None:
        fb = _fb(n_fft, sample_rate, n_mels)
    win = torch.hann_window(n_fft)
    X = torch.stft(waveform, n_fft, hop_length, n_fft, win, center=True, pad_mode="reflect",
                   normalized=False, onesided=True, return_complex=True)  # (4, F, T)
    P = X.real * X.real + X.imag * X.imag
    mel = torch.matmul(P.transpose(-1, -2), fb).transpose(-1, -2)
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    W = X[0]
    I = (W.real.unsqueeze(0) * X[1:].real + W.imag.unsqueeze(0) * X[1:].imag)
    E = 1e-8 + P[0] + P[1:].sum(0) / 3.0
    iv = torch.matmul((I / E.unsqueeze(0)).transpose(-1, -2), fb).transpose(-1, -2)
    return torch.cat([db, iv], dim=0)


def gcc_phat_port(waveform: torch.Tensor, n_fft: int, hop_length: int, n_lags: int = 64) -> torch.Tensor:
    """Second, independent restatement of the GCC-PHAT feature (no reference code exists): torch
    primitives in float64 — ``torch.stft`` (the transform the reference's log-mel path uses, dataset.py:38-50) and
    ``torch.fft.irfft`` — instead of the numpy framing + ``np.fft`` of oracle/features.py::gcc_phat.  (n_pairs, n_lags, T),
    pairs 01 02 03 12 13 23, lags [-n_lags/2, n_lags/2)."""
    x = waveform.to(torch.float64)
    win = torch.hann_window(n_fft, dtype=torch.float32).to(torch.float64)  # the reference's float32 table
    X = torch.stft(x, n_fft, hop_length, n_fft, win, center=True, pad_mode="reflect", normalized=False, onesided=True,
                   return_complex=True)  # (C, F, T)
    out = []
    C = X.shape[0]
    for m in range(C):
        for n in range(m + 1, C):
            R = torch.conj(X[m]) * X[n]
            mag = R.abs()
            ph = torch.where(mag > 0, R / torch.where(mag > 0, mag, torch.ones_like(mag)), torch.ones_like(R))
            cc = torch.fft.irfft(ph, n=n_fft, dim=0)  # (n_fft, T)
            out.append(torch.cat([cc[-(n_lags // 2):], cc[: n_lags // 2]], dim=0))
    return torch.stack(out, dim=0)


def metadata_to_labels_port(metadata_path, audio_duration, I=18, J=36, num_classes=14):
    """The reference's label encoder with the reference's COST structure (dataset.py:60-119): a torch tensor filled one
    element at a time from Python — per CSV row the five 20 ms frames of its cell, then the (frame, cell) double loop with
    a per-frame ``set`` lookup that writes the background class (1.94 M iterations for a 60 s clip, the >99 % of the
    reference's front-end time).  oracle/labels.py vectorises that last loop for the parity tests; this
    port keeps it so that bench.py's CPU label leg times what the reference really does.  (T, I*J, M) float32 tensor."""
    import pandas as pd

    n_frames = int((audio_duration * 1000) / 20)
    cells = I * J
    out = torch.zeros((n_frames, cells, num_classes), dtype=torch.float32)
    table = pd.read_csv(metadata_path, header=None)
    busy = [set() for _ in range(n_frames)]
    for _, rec in table.iterrows():
        f100, cls, az, el = int(rec.iloc[0]), int(rec.iloc[1]), int(rec.iloc[3]), int(rec.iloc[4])
        j = int(min(max((az + 180.0) / 360.0 * J, 0), J - 1))
        i = int(min(max((el + 90.0) / 180.0 * I, 0), I - 1))
        cell = i * J + j
        for t in range(5 * f100, min(5 * f100 + 5, n_frames)):
            out[t, cell, cls] = 1.0
            busy[t].add(cell)
    for t in range(n_frames):
        taken = busy[t]
        for cell in range(cells):
            if cell not in taken:
                out[t, cell, num_classes - 1] = 1.0
    return out


def class_mse_loss_port(y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
    """loss.py:43-54: softmax over the classes, mean squared error against the dense targets."""
    return torch.nn.functional.mse_loss(torch.softmax(y_pred, dim=-1), y_true)


def class_ce_loss_port(y_pred: torch.Tensor, y_true: torch.Tensor, class_weights: torch.Tensor | None = None) -> torch.Tensor:
    """loss.py:27-41 with the constructor's ``nn.CrossEntropyLoss(weight=class_weights)`` (loss.py:21-25): targets are
    the argmax of the dense rows."""
    M = y_pred.shape[-1]
    target = torch.argmax(y_true, dim=-1).view(-1)
    return torch.nn.CrossEntropyLoss(weight=class_weights)(y_pred.reshape(-1, M), target)


def aiur_loss_port(y_prob: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
    """loss.py:56-88: per frame, IoU between the cells whose most probable class is not the background (the last class) and
    the cells whose target argmax is not the background; a frame with neither counts as IoU 1; 1 - mean IoU."""
    bg = y_prob.shape[-1] - 1
    pred = (y_prob.argmax(dim=-1) != bg).to(y_prob.dtype)
    true = (y_true.argmax(dim=-1) != bg).to(y_prob.dtype)
    inter = (pred * true).sum(dim=-1)
    union = pred.sum(dim=-1) + true.sum(dim=-1) - inter
    iou = torch.where(union > 0, inter / (union + 1e-8), torch.ones_like(inter))
    return 1.0 - iou.mean()


def cl_loss_port(y_prob: torch.Tensor, y_true: torch.Tensor, I: int, J: int, eps: float = 1e-10) -> torch.Tensor:
    """loss.py:90-146 (eps = SMRSELDLoss.eps, loss.py:15): target map y' = 1 on background cells and -N_bac / (N_non + eps)
    on event cells of a frame; y_at = y' + (sum over the 8 neighbours on the circular (I, J) grid of (neighbour - y')) / 8,
    neighbours visited row by row from (-1, -1) to (1, 1); the loss is the sum over frames WITH events of
    (probability of any event class) * y_at, over (number of such frames * I * J + eps)."""
    B, T, G, M = y_prob.shape
    act_true = y_true.view(B, T, I, J, M)[..., :-1].sum(dim=-1)
    act_pred = y_prob.view(B, T, I, J, M)[..., :-1].sum(dim=-1)
    n_bac = (act_true < 0.01).sum(dim=(2, 3), keepdim=True).to(y_prob.dtype)
    n_non = (act_true > 0.01).sum(dim=(2, 3), keepdim=True).to(y_prob.dtype)
    yp = torch.where(act_true > 0.01, (-(n_bac / (n_non + eps))).expand_as(act_true), torch.ones_like(act_true))
    acc = torch.zeros_like(yp)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di or dj:
                acc = acc + (torch.roll(yp, shifts=(-di, -dj), dims=(2, 3)) - yp)  # element (i, j) sees (i + di, j + dj)
    y_at = yp + acc / 8.0
    has = (n_non > 0).to(y_prob.dtype)
    return ((act_pred * y_at) * has).sum() / (has.sum() * I * J + eps)
