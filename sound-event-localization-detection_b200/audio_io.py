"""Audio ingest: drop-in for the reference's ``load_audio`` (dataset.py:18-25, ``torchaudio.load``).

``torchaudio.load`` needs torchcodec in torchaudio >= 2.9 (absent from this image), so RIFF/WAVE files —
the only format of the STARSS22/23 ``foa_dev`` / ``mic_dev`` sets — are decoded here with numpy to the same
result: float32 ``(channels, samples)`` in [-1, 1) (integer PCM divided by 2^(bits-1))."""
from __future__ import annotations

import logging
import struct

import numpy as np
import torch

logger = logging.getLogger("SMR_SELD")


def read_wav(path: str, keep_pcm16: bool = False):
    """(channels, samples) float32 in [-1, 1) and the sample rate.  ``keep_pcm16``: a 16-bit PCM file comes back as the
    int16 samples themselves (the feature kernel converts them while loading, x / 32768 exactly like here)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            pcm = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, sr, _br, _ba, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the real tag
        tag = struct.unpack("<H", fmt[24:26])[0]
    n = len(pcm) // (ch * bits // 8)
    pcm = pcm[: n * ch * bits // 8]
    if tag == 1:
        if bits == 16 and keep_pcm16:
            return np.ascontiguousarray(np.frombuffer(pcm, dtype="<i2").reshape(n, ch).T), sr
        if bits == 16:
            x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 32:
            x = (np.frombuffer(pcm, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        elif bits == 24:
            b = np.frombuffer(pcm, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = (v.astype(np.float64) / 8388608.0).astype(np.float32)
        elif bits == 8:
            x = (np.frombuffer(pcm, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(pcm, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    return np.ascontiguousarray(x.reshape(n, ch).T), sr


def write_wav_pcm16(path: str, x: np.ndarray, sr: int) -> None:
    """(channels, samples) float -> 16-bit PCM WAVE (used by tests and examples)."""
    x = np.asarray(x)
    ch, n = x.shape
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, sr, sr * ch * 2, ch * 2, 16) + b"data" + struct.pack("<I", len(pcm))
    with open(path, "wb") as f:
        f.write(hdr + pcm)


def load_audio(audio_path):
    """Drop-in for reference dataset.py:18-25: returns (float32 tensor (channels, samples), sample_rate) and
    warns when the file does not have 4 channels."""
    x, sr = read_wav(str(audio_path))
    waveform = torch.from_numpy(x)
    if waveform.shape[0] != 4:
        logger.warning(f"Expected 4 channels but got {waveform.shape[0]} channels in {audio_path}")
    return waveform, sr


def load_audio_pcm16(audio_path):
    """Ingest form of ``load_audio`` used by ``SELDDataset``: 16-bit PCM files — the STARSS sets — stay
    int16 ``(channels, samples)``, half the bytes over PCIe and no float pass on the host; the feature kernel divides by
    32768 while loading, bit-identical to ``load_audio``'s result.  Other encodings come back as float32 like ``load_audio``."""
    x, sr = read_wav(str(audio_path), keep_pcm16=True)
    waveform = torch.from_numpy(x)
    if waveform.shape[0] != 4:
        logger.warning(f"Expected 4 channels but got {waveform.shape[0]} channels in {audio_path}")
    return waveform, sr
