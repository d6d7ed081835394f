"""Feature front-end: host side of the fused STFT + log-mel (+ IV | + GCC-PHAT) kernels.

Mirrors the reference's ``audio_to_mel_spectrogram`` (dataset.py:27-58) — same name, arguments and
result shape — on top of ``seld_features`` in libseld_cuda.  PyTorch is used for device memory, streams
and the two constant tables (window, filterbank); all arithmetic on the audio happens in the CUDA kernels.
"""
from __future__ import annotations

import ctypes
import math
import os
import threading

import torch

from . import _lib
from ._lib import SELD_MODE_LOGMEL, SELD_MODE_LOGMEL_GCC, SELD_MODE_LOGMEL_IV

MODES = {"logmel": SELD_MODE_LOGMEL, "logmel_iv": SELD_MODE_LOGMEL_IV, "foa_iv": SELD_MODE_LOGMEL_IV,
         "logmel_gcc": SELD_MODE_LOGMEL_GCC, "mic_gcc": SELD_MODE_LOGMEL_GCC}


def hann_window(n_fft: int) -> torch.Tensor:
    """The window torchaudio.transforms.MelSpectrogram uses (torchaudio/transforms/_transforms.py:604-616):
    ``torch.hann_window(n_fft)``, periodic, built by ATen in float32."""
    return torch.hann_window(n_fft, dtype=torch.float32)


def mel_filterbank(n_fft: int, sample_rate: int, n_mels: int) -> torch.Tensor:
    """(n_fft//2+1, n_mels) float32 HTK triangular filterbank, f_min 0, f_max sr//2, norm None — the table
    ``melscale_fbanks`` builds for MelSpectrogram's defaults (torchaudio/functional/functional.py:492-587,
    restated with the same float32 torch ops so the weights are bit-identical; tests compare with the
    table dumped from torchaudio)."""
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + float(sample_rate // 2) / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


class FeaturePlan:
    """Owns one ``seld_plan`` (constant tables on one device).  Replaces the MelSpectrogram object the
    reference rebuilds on every call (dataset.py:38-43)."""

    def __init__(self, n_fft: int, hop: int, n_mels: int, sample_rate: int, device: torch.device):
        self.n_fft, self.hop, self.n_mels, self.sample_rate = int(n_fft), int(hop), int(n_mels), int(sample_rate)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SeldError("seld_cuda has no CPU path: a CUDA device is required")
        self.window = hann_window(self.n_fft)
        self.fb = mel_filterbank(self.n_fft, self.sample_rate, self.n_mels)
        handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().seld_plan_create(ctypes.byref(handle), dev_index, self.n_fft, self.hop, self.n_mels,
                                               self.window.data_ptr(), self.fb.data_ptr()), "seld_plan_create")
        self._handle = handle
        self.device = torch.device("cuda", dev_index)
        self.fast = bool(_lib.lib().seld_plan_has_fast_path(handle))  # reference filterbank -> fast kernel for C == 4

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _lib.lib().seld_plan_destroy(h)
            except Exception:
                pass
            self._handle = None

    def num_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def out_channels(self, mode: int, C: int) -> int:
        return {SELD_MODE_LOGMEL: C, SELD_MODE_LOGMEL_IV: 7, SELD_MODE_LOGMEL_GCC: 10}[mode]

    def run(self, audio: torch.Tensor, mode="logmel", lengths: torch.Tensor | None = None, out: torch.Tensor | None = None,
            c_off: int = 0, stats: torch.Tensor | None = None, stat_frames: torch.Tensor | None = None,
            spec: torch.Tensor | None = None, T_out: int | None = None, isolate_channels: bool = False,
            mean: torch.Tensor | None = None, inv_std: torch.Tensor | None = None, layout: str = "tcf",
            out_dtype: torch.dtype = torch.float32, check: bool | None = None) -> torch.Tensor:
        """audio (B, C, N) float32 — or int16 PCM, converted as x / 32768 inside the kernel — CUDA (any strides with
        unit sample stride) -> (B, T_out, C_out, n_mels).

        ``mean`` / ``inv_std`` (float32 (C_out, n_mels) CUDA): write (x - mean) * inv_std, the scaler apply step fused
        into the kernel; ``layout="ctf"``: write (B, C_out, T_out, n_mels), what every reference model permutes to first
        (model_conformer.py:191); ``out_dtype=torch.bfloat16``.  These options and int16 input need the fast path
        (4 channels, 64 HTK mels).
        ``check``: after a ragged call (``lengths``), read the device status word back and raise if a clip was too
        short for reflect padding (len <= n_fft/2; torch.stft raises there).  Default: only when ``lengths`` is given.
        ``isolate_channels`` (log-mel mode only): transform every channel on its own through the generic kernel.  The
        default path no longer needs it: channel pairs are level-equalised per frame (block floating point), which
        removes the rounding-noise coupling between a loud and a quiet channel of one packed FFT."""
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        if isolate_channels:
            if mode != SELD_MODE_LOGMEL or spec is not None:
                raise ValueError("isolate_channels is available for mode='logmel' without a spectrum dump")
            Cn = audio.shape[1]
            if T_out is None:
                T_out = self.num_frames(audio.shape[2])
            if out is None:
                out = torch.empty((audio.shape[0], T_out, c_off + Cn, self.n_mels), dtype=torch.float32, device=self.device)
            for c in range(Cn):  # a one-channel group pairs the channel with an exact zero: nothing can leak into it
                self.run(audio[:, c:c + 1], mode=mode, lengths=lengths, out=out, c_off=c_off + c, T_out=T_out, check=check)
            if stats is not None:
                self.accumulate_stats(out, stats, n_samples=audio.shape[2], lengths=lengths, stat_frames=stat_frames,
                                      c_off=c_off, n_channels=Cn)
            return out
        if audio.dim() != 3:
            raise ValueError("audio must be (B, C, N)")
        if audio.dtype not in (torch.float32, torch.int16) or not audio.is_cuda:
            raise ValueError("audio must be a float32 (or int16 PCM) CUDA tensor")
        if audio.device != self.device:
            raise ValueError(f"audio on {audio.device}, plan on {self.device}")
        if audio.stride(2) != 1:
            audio = audio.contiguous()
        B, Cn, N = audio.shape
        if audio.dtype == torch.int16 and not (self.fast and Cn == 4 and spec is None and mode != SELD_MODE_LOGMEL_GCC):
            # outside the fast path the kernels take float32: convert on the device first (x / 32768)
            f32 = torch.empty((B, Cn, N), dtype=torch.float32, device=self.device)
            src = audio.contiguous()
            _lib.check(_lib.lib().seld_pcm16_to_float(src.data_ptr(), f32.data_ptr(), src.numel(),
                                                      torch.cuda.current_stream(self.device).cuda_stream), "seld_pcm16_to_float")
            audio = f32
        if T_out is None:
            T_out = self.num_frames(N)
        n_out = self.out_channels(mode, Cn)
        if layout not in ("tcf", "ctf"):
            raise ValueError("layout must be 'tcf' (B, T, C, F) or 'ctf' (B, C, T, F)")
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
        ctf = layout == "ctf"
        if out is None:
            shape = (B, c_off + n_out, T_out, self.n_mels) if ctf else (B, T_out, c_off + n_out, self.n_mels)
            out = torch.empty(shape, dtype=out_dtype, device=self.device)
        t_axis, c_axis = (2, 1) if ctf else (1, 2)
        if (out.dtype != out_dtype or not out.is_contiguous() or out.dim() != 4 or out.shape[0] != B
                or out.shape[t_axis] != T_out or out.shape[3] != self.n_mels or out.device != self.device):
            raise ValueError(f"out must be a contiguous {out_dtype} (B, T_out, C_out, n_mels) tensor ('ctf': (B, C_out, T_out, "
                             "n_mels)) on the plan's device")
        C_out = out.shape[c_axis]
        if (mean is None) != (inv_std is None):
            raise ValueError("mean and inv_std go together")
        for name, t, dt in (("lengths", lengths, torch.int64), ("stats", stats, torch.float64),
                            ("stat_frames", stat_frames, torch.int32), ("spec", spec, torch.complex64),
                            ("mean", mean, torch.float32), ("inv_std", inv_std, torch.float32)):
            if t is not None and (t.dtype != dt or not t.is_contiguous() or t.device != self.device):
                raise ValueError(f"{name} must be a contiguous {dt} tensor on {self.device}")
        if stats is not None and stats.numel() != 2 * C_out * self.n_mels:
            raise ValueError("stats must hold 2 * C_out * n_mels float64 values")
        if mean is not None and (mean.numel() != C_out * self.n_mels or inv_std.numel() != C_out * self.n_mels):
            raise ValueError("mean / inv_std must hold C_out * n_mels float32 values")
        if spec is not None and tuple(spec.shape) != (B, Cn, T_out, self.n_fft // 2 + 1):
            raise ValueError("spec must be (B, C, T_out, n_fft//2+1) complex64")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        opts = None
        if audio.dtype == torch.int16 or ctf or out_dtype != torch.float32 or mean is not None:
            opts = _lib.FeatOpts(_lib.SELD_DTYPE_I16 if audio.dtype == torch.int16 else _lib.SELD_DTYPE_F32,
                                 _lib.SELD_LAYOUT_CTF if ctf else _lib.SELD_LAYOUT_TCF,
                                 _lib.SELD_DTYPE_BF16 if out_dtype == torch.bfloat16 else _lib.SELD_DTYPE_F32,
                                 _lib.ptr(mean), _lib.ptr(inv_std))
        # (no torch.cuda.device guard needed: every ABI entry point runs on its plan's device and restores the caller's)
        _lib.check(_lib.lib().seld_features_ex(
            self._handle, mode, audio.data_ptr(), audio.stride(0), audio.stride(1), N, _lib.ptr(lengths), B, Cn,
            out.data_ptr(), T_out, C_out, c_off, _lib.ptr(stats), _lib.ptr(stat_frames), _lib.ptr(spec),
            ctypes.byref(opts) if opts is not None else None, stream), "seld_features")
        if check or (check is None and lengths is not None):
            self.check_status()
        return out

    def check_status(self) -> None:
        """Synchronise the current stream and raise ``SeldError`` if a kernel flagged a clip as too short for reflect
        padding (``seld_plan_status``)."""
        st = ctypes.c_int(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().seld_plan_status(self._handle, stream, ctypes.byref(st)), "seld_plan_status")


    def accumulate_stats(self, feats: torch.Tensor, stats: torch.Tensor, n_samples: int = 0,
                         lengths: torch.Tensor | None = None, stat_frames: torch.Tensor | None = None, c_off: int = 0,
                         n_channels: int | None = None) -> torch.Tensor:
        """Scaler partials of a finished feature tensor (B, T_out, C_out, n_mels): what ``run(stats=...)`` adds,
        as a call of its own (``seld_feature_stats``)."""
        if feats.dtype != torch.float32 or not feats.is_contiguous() or feats.dim() != 4 or feats.device != self.device:
            raise ValueError("feats must be a contiguous float32 (B, T_out, C_out, n_mels) tensor on the plan's device")
        B, T_out, C_out, M = feats.shape
        if M != self.n_mels or stats.dtype != torch.float64 or stats.numel() != 2 * C_out * M or stats.device != self.device:
            raise ValueError("stats must hold 2 * C_out * n_mels float64 values on the plan's device")
        if stat_frames is None and lengths is None and n_samples <= 0:
            raise ValueError("give stat_frames, lengths or n_samples")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().seld_feature_stats(
            self._handle, feats.data_ptr(), B, T_out, C_out, c_off, C_out - c_off if n_channels is None else n_channels,
            int(n_samples), _lib.ptr(lengths), _lib.ptr(stat_frames), stats.data_ptr(), stream), "seld_feature_stats")
        return stats


_plans: dict = {}
_plans_lock = threading.Lock()


def get_plan(n_fft: int, hop: int, n_mels: int, sample_rate: int, device=None) -> FeaturePlan:
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    # the library reads its A/B switches once per plan (seld_plan_create): a changed switch means a new plan
    key = (int(n_fft), int(hop), int(n_mels), int(sample_rate), str(device), os.environ.get("SELD_FEAT_IMPL", ""),
           os.environ.get("SELD_V3_CFG", ""))
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = FeaturePlan(n_fft, hop, n_mels, sample_rate, device)
    return p


def extract_features(audio: torch.Tensor, sample_rate: int = 24000, n_fft: int = 1024, hop_length: int = 480,
                     n_mels: int = 64, mode: str = "logmel", **kw) -> torch.Tensor:
    """Batch API: (B, C, N) float32 CUDA -> (B, T, C_out, n_mels) frame-major features."""
    return get_plan(n_fft, hop_length, n_mels, sample_rate, audio.device).run(audio, mode=mode, **kw)


def _default(name):
    from .config import get_config
    return getattr(get_config(), name)


def audio_to_mel_spectrogram(waveform: torch.Tensor, sample_rate: int, n_fft=None, hop_length=None, n_mels=None,
                             device=None, isolate_channels: bool = False) -> torch.Tensor:
    """Drop-in for reference dataset.py:27-58: (C, N) waveform -> (C, n_mels, 1 + N//hop) float32 dB.

    ``None`` arguments fall back to the Config values like the reference (dataset.py:30-35).  A CPU waveform
    is copied to the GPU, transformed there and the result returned as a contiguous CPU tensor (what
    unmodified main.py / DataLoader workers need); a CUDA waveform returns a CUDA tensor that is a
    (C, n_mels, T) view of the kernel's frame-major output."""
    if n_fft is None:
        n_fft = _default("SPECTROGRAM_N_FFT")
    if hop_length is None:
        hop_length = _default("SPECTROGRAM_HOP_LENGTH")
    if n_mels is None:
        n_mels = _default("N_MELS")
    if waveform.dim() != 2:
        raise ValueError("waveform must be (channels, samples)")
    was_cpu = not waveform.is_cuda
    dev = torch.device(device) if device is not None else (waveform.device if waveform.is_cuda else torch.device("cuda"))
    x = waveform.to(device=dev, dtype=torch.float32, non_blocking=True)
    plan = get_plan(n_fft, hop_length, n_mels, sample_rate, x.device)
    out = plan.run(x.unsqueeze(0), mode="logmel", isolate_channels=isolate_channels)[0]  # (T, C, M)
    res = out.permute(1, 2, 0)
    return res.contiguous().cpu() if was_cpu else res


def extract_features_host(audio_host: torch.Tensor, out_host: torch.Tensor, plan: FeaturePlan, mode="logmel",
                          chunk: int = 16, n_streams: int = 3) -> torch.Tensor:
    """End-to-end path for HOST buffers: (B, C, N) float32 — or int16 PCM, converted as x / 32768 inside the feature
    kernel's loads like torchaudio.load does for 16-bit WAV (reference dataset.py:18-25) — (ideally pinned) -> out_host
    (B, T, C_out, n_mels).

    Clips are streamed through the GPU in chunks on ``n_streams`` CUDA streams so that the host->device
    copy of chunk i+1, the kernel of chunk i and the device->host copy of chunk i-1 overlap (PCIe is full
    duplex).  Returns when the features are in ``out_host``."""
    if audio_host.is_cuda or out_host.is_cuda:
        raise ValueError("extract_features_host takes host tensors; use FeaturePlan.run for device tensors")
    if audio_host.dtype not in (torch.float32, torch.int16):
        raise ValueError("audio_host must be float32 or int16 (PCM16)")
    pcm = audio_host.dtype == torch.int16
    B, Cn, N = audio_host.shape
    T = plan.num_frames(N)
    m = MODES[mode] if isinstance(mode, str) else int(mode)
    n_out = plan.out_channels(m, Cn)
    if tuple(out_host.shape) != (B, T, n_out, plan.n_mels) or out_host.dtype != torch.float32:
        raise ValueError(f"out_host must be float32 {(B, T, n_out, plan.n_mels)}")
    key = (chunk, Cn, N, n_out, n_streams, pcm)
    cache = plan.__dict__.setdefault("_host_pipes", {})
    pipe = cache.get(key)
    if pipe is None:
        direct = pcm and plan.fast and Cn == 4 and m != SELD_MODE_LOGMEL_GCC  # the fast kernel loads int16 itself
        pipe = cache[key] = [(torch.cuda.Stream(plan.device),
                              None if direct else torch.empty((chunk, Cn, N), dtype=torch.float32, device=plan.device),
                              torch.empty((chunk, T, n_out, plan.n_mels), dtype=torch.float32, device=plan.device),
                              torch.empty((chunk, Cn, N), dtype=torch.int16, device=plan.device) if pcm else None)
                             for _ in range(n_streams)]
    cur = torch.cuda.current_stream(plan.device)
    for s, *_ in pipe:
        s.wait_stream(cur)
    for i, b0 in enumerate(range(0, B, chunk)):
        s, d_in, d_out, d_pcm = pipe[i % n_streams]
        nb = min(chunk, B - b0)
        with torch.cuda.stream(s):
            if pcm:
                d_pcm[:nb].copy_(audio_host[b0:b0 + nb], non_blocking=True)
                if d_in is not None:  # configurations outside the fast path: separate conversion pass
                    _lib.check(_lib.lib().seld_pcm16_to_float(d_pcm.data_ptr(), d_in.data_ptr(), nb * Cn * N, s.cuda_stream),
                               "seld_pcm16_to_float")
                src = d_pcm if d_in is None else d_in
            else:
                d_in[:nb].copy_(audio_host[b0:b0 + nb], non_blocking=True)
                src = d_in
            plan.run(src[:nb], mode=m, out=d_out[:nb])
            out_host[b0:b0 + nb].copy_(d_out[:nb], non_blocking=True)
    for s, *_ in pipe:
        s.synchronize()
    return out_host
