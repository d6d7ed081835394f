"""Label encoders: host side of the dense SELD grid label kernels (seld_labels_fill / seld_labels_paint).

Drop-ins for the reference's ``metadata_to_labels`` (dataset.py:60-119), ``polar_to_grid``
(utils.py:77-90) and ``augment_with_gaussian_noise`` (smrl_seld_gaussian.py:397-534): same names, arguments,
return values and exceptions.  The host does what must stay in Python to be bit-exact — the CSV parse with
pandas + ``int()``, ``total_frames`` in Python floats, the global-RNG draws — and turns each CSV row into a
compact event ``{row0, row1, class, cell}``; the dense ``(T, I*J, M)`` float32 tensor (108.9 MB per minute
of audio) is written by the CUDA kernels.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import _lib

FRAMES_PER_METADATA_FRAME = 100 // 20  # dataset.py:68-70


def polar_to_grid(phi, theta, I=None, J=None, cell_size_deg=None):
    """utils.py:77-90 — (azimuth, elevation) in degrees -> (i, j) grid indices, float64 arithmetic."""
    if (I is None or J is None) and cell_size_deg is not None:
        I = int(180 // cell_size_deg)
        J = int(360 // cell_size_deg)
    elif I is None or J is None:
        raise ValueError("Either provide (I, J) or cell_size_deg for polar_to_grid")
    phi_norm = (phi + 180.0) / 360.0
    theta_norm = (theta + 90.0) / 180.0
    j = int(np.clip(phi_norm * J, 0, J - 1))
    i = int(np.clip(theta_norm * I, 0, I - 1))
    return i, j


def _cells_of(az: np.ndarray, el: np.ndarray, I: int, J: int) -> np.ndarray:
    """Vectorised polar_to_grid + flattening (dataset.py:105-106): identical float64 operations."""
    phi_norm = (az.astype(np.float64) + 180.0) / 360.0
    theta_norm = (el.astype(np.float64) + 90.0) / 180.0
    j = np.clip(phi_norm * J, 0, J - 1).astype(np.int64)
    i = np.clip(theta_norm * I, 0, I - 1).astype(np.int64)
    return i * J + j


def total_frames_of(audio_duration: float) -> int:
    """dataset.py:73 — ``int((audio_duration * 1000) / 20)`` in Python floats (differs from N // 480 for
    some lengths, e.g. N = 97 440 -> 202, not 203)."""
    return int((audio_duration * 1000) / 20)


def _grid(I, J, cell_size_deg):
    if (I is None or J is None) and cell_size_deg is not None:
        I = int(180 // cell_size_deg)
        J = int(360 // cell_size_deg)
    elif I is None or J is None:
        raise ValueError("Either provide (I, J) or cell_size_deg for grid dimensions")
    return I, J


def read_metadata(metadata_path):
    """``pd.read_csv(header=None)`` + the ``int(row.iloc[c])`` casts of dataset.py:86-97, vectorised.
    ``iterrows`` upcasts a row to the common dtype of ALL columns, so one float column (STARSS23 distance)
    makes every cell a float64 that ``int()`` truncates toward zero.  Returns (df, int64 array (rows, 5))."""
    df = pd.read_csv(metadata_path, header=None)
    if df.shape[1] < 5:
        if len(df):
            raise IndexError("single positional indexer is out-of-bounds")  # row.iloc[4] in the reference
        return df, np.zeros((0, 5), dtype=np.int64)
    if all(pd.api.types.is_numeric_dtype(t) and not pd.api.types.is_bool_dtype(t) for t in df.dtypes):
        arr = df.to_numpy()
        first5 = arr[:, :5]
        if np.issubdtype(first5.dtype, np.floating):
            if np.isnan(first5).any():
                raise ValueError("cannot convert float NaN to integer")
            if np.isinf(first5).any():
                raise OverflowError("cannot convert float infinity to integer")
            first5 = np.trunc(first5)
        return df, first5.astype(np.int64)
    rows = [[int(row.iloc[c]) for c in range(5)] for _, row in df.iterrows()]  # exotic dtypes: literal path
    return df, np.asarray(rows, dtype=np.int64).reshape(-1, 5)


def _row_ranges(meta_frames: np.ndarray, total_frames: int):
    """Frames ``range(5f, min(5f+5, T))`` of each CSV row as [row0, row1) ranges, including Python's
    negative-index wrap-around for negative metadata frames.  Returns (event index, row0, row1)."""
    start = meta_frames * FRAMES_PER_METADATA_FRAME
    end = np.minimum(start + FRAMES_PER_METADATA_FRAME, total_frames)
    idx = np.arange(len(start))
    keep = start < end
    if (start[keep] < -total_frames).any():
        raise IndexError(f"index out of range for dimension 0 with size {total_frames}")
    pos = keep & (start >= 0)
    neg = keep & (start < 0)
    ev, r0, r1 = [idx[pos]], [start[pos]], [end[pos]]
    if neg.any():
        ev.append(idx[neg]); r0.append(start[neg] + total_frames); r1.append(np.minimum(end[neg], 0) + total_frames)
        cross = neg & (end > 0)
        ev.append(idx[cross]); r0.append(np.zeros(int(cross.sum()), dtype=np.int64)); r1.append(end[cross])
    return np.concatenate(ev), np.concatenate(r0), np.concatenate(r1)


def _wrap_classes(cls: np.ndarray, num_classes: int) -> np.ndarray:
    cls = np.where(cls < 0, cls + num_classes, cls)
    if ((cls < 0) | (cls >= num_classes)).any():
        raise IndexError(f"index out of range for dimension 2 with size {num_classes}")
    return cls


def point_events(metadata_path, audio_duration, I, J, num_classes=14):
    """Compact events of ``metadata_to_labels``: int32 (E, 4) rows {row0, row1, class, cell}, and T."""
    total_frames = total_frames_of(audio_duration)
    _, rows = read_metadata(metadata_path)
    if total_frames <= 0 or len(rows) == 0:
        return np.zeros((0, 4), dtype=np.int32), max(total_frames, 0)
    ev, r0, r1 = _row_ranges(rows[:, 0], total_frames)
    cls = _wrap_classes(rows[ev, 1], num_classes)  # only rows that paint at least one frame can raise
    cells = _cells_of(rows[:, 3], rows[:, 4], I, J)
    events = np.stack([r0, r1, cls, cells[ev]], axis=1).astype(np.int32)
    return events, total_frames


def draw_source_noise(df, sigma_azimuth, sigma_elevation):
    """smrl_seld_gaussian.py:427-440: one (azimuth, elevation) offset per (class, source) in pandas groupby
    order, azimuth first, from numpy's GLOBAL legacy RNG — exactly the reference's draws."""
    unique_sources = df.groupby([1, 2]).first().reset_index()
    noise = {}
    for _, source_row in unique_sources.iterrows():
        key = (int(source_row.iloc[0]), int(source_row.iloc[1]))
        az = np.random.normal(0, sigma_azimuth)
        el = np.random.normal(0, sigma_elevation)
        noise[key] = (az, el)
    return noise


def region_events(metadata_path, audio_duration, I, J, num_classes=14, sigma_azimuth=5.0, sigma_elevation=5.0):
    """Compact events of ``augment_with_gaussian_noise``: events with cell = -1 plus float64 centres
    (azimuth + noise, elevation + noise); the +-2 sigma cell test runs on the GPU."""
    total_frames = total_frames_of(audio_duration)
    df, rows = read_metadata(metadata_path)
    if len(rows) == 0:
        return np.zeros((0, 4), dtype=np.int32), np.zeros((0, 2), dtype=np.float64), max(total_frames, 0)
    noise = draw_source_noise(df, sigma_azimuth, sigma_elevation)  # drawn even if no frame is painted
    if total_frames <= 0:
        return np.zeros((0, 4), dtype=np.int32), np.zeros((0, 2), dtype=np.float64), 0
    nz = np.array([noise[(int(c), int(s))] for c, s in rows[:, 1:3]], dtype=np.float64).reshape(-1, 2)
    centres = np.stack([rows[:, 3] + nz[:, 0], rows[:, 4] + nz[:, 1]], axis=1)  # int + float64, as in :469-470
    ev, r0, r1 = _row_ranges(rows[:, 0], total_frames)
    cls = _wrap_classes(rows[ev, 1], num_classes)
    events = np.stack([r0, r1, cls, np.full(len(ev), -1)], axis=1).astype(np.int32)
    return events, np.ascontiguousarray(centres[ev]), total_frames


def encode_dense(out: torch.Tensor, events: np.ndarray, centres: np.ndarray | None, I: int, J: int,
                 sigma_azimuth: float = 5.0, sigma_elevation: float = 5.0, fill: bool = True) -> torch.Tensor:
    """Run the label kernels on ``out`` (rows, I*J, M) float32 CUDA: background fill, then paint events
    (rows are absolute row indices into ``out``)."""
    if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.dim() != 3:
        raise ValueError("out must be a contiguous float32 CUDA tensor (rows, cells, classes)")
    rows, cells, M = out.shape
    if cells != I * J:
        raise ValueError("out.shape[1] must equal I*J")
    lib = _lib.lib()
    stream = torch.cuda.current_stream(out.device).cuda_stream
    with torch.cuda.device(out.device):
        if fill:
            _lib.check(lib.seld_labels_fill(out.data_ptr(), rows, cells, M, stream), "seld_labels_fill")
        n = int(len(events))
        if n:
            ev = torch.from_numpy(np.ascontiguousarray(events, dtype=np.int32)).to(out.device, non_blocking=False)
            ce = None
            if centres is not None and len(centres):
                ce = torch.from_numpy(np.ascontiguousarray(centres, dtype=np.float64)).to(out.device)
            elif (events[:, 3] < 0).any():
                raise ValueError("region events need centres")
            _lib.check(lib.seld_labels_paint(out.data_ptr(), rows, I, J, M, ev.data_ptr(), _lib.ptr(ce), n,
                                             float(sigma_azimuth), float(sigma_elevation), stream),
                       "seld_labels_paint")
            # ev / ce are freed by the caching allocator in stream order (same stream as the kernels)
    return out


def _result(labels: torch.Tensor, device):
    return labels.cpu() if device is None else labels


def _cuda_device(device):
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.SeldError("seld_cuda has no CPU path: a CUDA device is required")
    return device


def metadata_to_labels(metadata_path, audio_duration, sample_rate=24000, I=None, J=None, cell_size_deg=None,
                       num_classes=14, device=None):
    """Drop-in for reference dataset.py:60-119.  Returns ``(labels (T, I*J, M) float32, I, J)``.

    ``device=None`` keeps the reference's contract (a CPU tensor; the encoding still runs on the GPU and is
    copied back); ``device="cuda"`` returns the CUDA tensor."""
    if cell_size_deg is None:
        from .config import get_config
        cell_size_deg = get_config().GRID_CELL_DEGREES
    I, J = _grid(I, J, cell_size_deg)
    events, T = point_events(metadata_path, audio_duration, I, J, num_classes)
    out = torch.empty((T, I * J, num_classes), dtype=torch.float32, device=_cuda_device(device))
    encode_dense(out, events, None, I, J)
    return _result(out, device), I, J


def augment_with_gaussian_noise(metadata_path, audio_duration, sample_rate=24000, I=None, J=None, cell_size_deg=None,
                                num_classes=14, sigma_azimuth=5.0, sigma_elevation=5.0, device=None):
    """Drop-in for reference smrl_seld_gaussian.py:397-534 (consumes numpy's global RNG identically)."""
    if cell_size_deg is None:
        from .config import get_config
        cell_size_deg = get_config().GRID_CELL_DEGREES
    I, J = _grid(I, J, cell_size_deg)
    events, centres, T = region_events(metadata_path, audio_duration, I, J, num_classes, sigma_azimuth, sigma_elevation)
    out = torch.empty((T, I * J, num_classes), dtype=torch.float32, device=_cuda_device(device))
    encode_dense(out, events, centres, I, J, sigma_azimuth, sigma_elevation)
    return _result(out, device), I, J
