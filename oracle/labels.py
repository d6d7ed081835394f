"""numpy restatement of the reference label encoders (TEST INFRASTRUCTURE — see oracle/__init__.py).

* ``polar_to_grid``                 reference utils.py:77-90
* ``metadata_to_labels``            reference dataset.py:60-119
* ``augment_with_gaussian_noise``   reference smrl_seld_gaussian.py:397-534
* ``create_windows``                reference dataset.py:267-317

The Python loops of the reference are kept for the per-row part (≈10^3 rows) and vectorised only for the
(frame, cell) background fill, so every index, wrap-around and comparison follows the reference line by
line.  Pinned against outputs of the reference itself (tests/golden/labels_*.npz).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def polar_to_grid(phi, theta, I=None, J=None, cell_size_deg=None):
    """utils.py:77-90."""
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


def total_frames_of(audio_duration: float) -> int:
    """dataset.py:73 — int((audio_duration * 1000) / 20) in Python floats."""
    return int((audio_duration * 1000) / 20)


def _grid(I, J, cell_size_deg):
    if (I is None or J is None) and cell_size_deg is not None:
        I = int(180 // cell_size_deg)
        J = int(360 // cell_size_deg)
    elif I is None or J is None:
        raise ValueError("Either provide (I, J) or cell_size_deg for grid dimensions")
    return I, J


def _rows(metadata_path):
    """pd.read_csv(header=None) + iterrows + int(row.iloc[i]) for i in 0..4 (dataset.py:86-97)."""
    df = pd.read_csv(metadata_path, header=None)
    out = []
    for _, row in df.iterrows():
        out.append(tuple(int(row.iloc[c]) for c in range(5)))
    return df, out


def _finish_background(labels, active, num_classes):
    """dataset.py:114-117: every (t, cell) not in the per-frame active set gets class M-1 = 1.0."""
    labels[:, :, num_classes - 1][~active] = 1.0
    return labels


def metadata_to_labels(metadata_path, audio_duration, sample_rate=24000, I=None, J=None,
                       cell_size_deg=10, num_classes=14):
    """dataset.py:60-119.  Returns (float32 ndarray (T, I*J, M), I, J)."""
    frames_per_metadata_frame = 100 // 20
    total_frames = total_frames_of(audio_duration)
    I, J = _grid(I, J, cell_size_deg)
    total_cells = I * J
    labels = np.zeros((total_frames, total_cells, num_classes), dtype=np.float32)
    active = np.zeros((total_frames, total_cells), dtype=bool)
    _, rows = _rows(metadata_path)
    for metadata_frame, active_class, _source, azimuth, elevation in rows:
        start_frame = metadata_frame * frames_per_metadata_frame
        end_frame = min(start_frame + frames_per_metadata_frame, total_frames)
        i, j = polar_to_grid(azimuth, elevation, I=I, J=J)
        cell_idx = i * J + j
        for t in range(start_frame, end_frame):
            labels[t, cell_idx, active_class] = 1.0  # negative t / class wrap like the reference's indexing
            active[t, cell_idx] = True
    return _finish_background(labels, active, num_classes), I, J


def draw_source_noise(df, sigma_azimuth, sigma_elevation):
    """smrl_seld_gaussian.py:427-440: one (az, el) draw per (class, source) in groupby order from the
    numpy GLOBAL legacy RNG (np.random.normal), azimuth first."""
    unique_sources = df.groupby([1, 2]).first().reset_index()
    noise = {}
    for _, source_row in unique_sources.iterrows():
        key = (int(source_row.iloc[0]), int(source_row.iloc[1]))
        az = np.random.normal(0, sigma_azimuth)
        el = np.random.normal(0, sigma_elevation)
        noise[key] = (az, el)
    return noise


def region_cells(center_azimuth, center_elevation, sigma_azimuth, sigma_elevation, I, J):
    """smrl_seld_gaussian.py:474-518: cells whose centre lies inside the +-2 sigma rectangle (float64,
    same operation order and <= comparisons as the reference)."""
    elevation_min = max(center_elevation - 2 * sigma_elevation, -90)
    elevation_max = min(center_elevation + 2 * sigma_elevation, 90)
    cells = []
    cell_size_elevation = 180.0 / I
    cell_size_azimuth = 360.0 / J
    for grid_i in range(I):
        cell_elevation = -90 + (grid_i + 0.5) * cell_size_elevation
        if not (elevation_min <= cell_elevation <= elevation_max):
            continue
        for grid_j in range(J):
            cell_azimuth = -180 + (grid_j + 0.5) * cell_size_azimuth
            diff = cell_azimuth - center_azimuth
            while diff > 180:
                diff -= 360
            while diff < -180:
                diff += 360
            if abs(diff) <= 2 * sigma_azimuth:
                cells.append(grid_i * J + grid_j)
    return cells


def augment_with_gaussian_noise(metadata_path, audio_duration, sample_rate=24000, I=None, J=None,
                                cell_size_deg=10, num_classes=14, sigma_azimuth=5.0, sigma_elevation=5.0):
    """smrl_seld_gaussian.py:397-534.  Consumes the numpy global RNG exactly like the reference."""
    frames_per_metadata_frame = 100 // 20
    total_frames = total_frames_of(audio_duration)
    I, J = _grid(I, J, cell_size_deg)
    total_cells = I * J
    labels = np.zeros((total_frames, total_cells, num_classes), dtype=np.float32)
    active = np.zeros((total_frames, total_cells), dtype=bool)
    df, rows = _rows(metadata_path)
    noise = draw_source_noise(df, sigma_azimuth, sigma_elevation)
    for metadata_frame, active_class, source_num, azimuth, elevation in rows:
        az_n, el_n = noise[(active_class, source_num)]
        start_frame = metadata_frame * frames_per_metadata_frame
        end_frame = min(start_frame + frames_per_metadata_frame, total_frames)
        cells = region_cells(azimuth + az_n, elevation + el_n, sigma_azimuth, sigma_elevation, I, J)
        for cell_idx in cells:
            for t in range(start_frame, end_frame):
                labels[t, cell_idx, active_class] = 1.0
                active[t, cell_idx] = True
    return _finish_background(labels, active, num_classes), I, J


def create_windows(spec_cft: np.ndarray, labels_tgm: np.ndarray, window_frames=250, hop_frames=50):
    """dataset.py:267-317 on the concatenated (C, F, T) features and (T, G, M) labels.
    Returns a list of (spec (W, C, F), labels (W, G, M), start, end)."""
    C, F, T = spec_cft.shape
    M = labels_tgm.shape[2]
    out = []
    start = 0
    while start < T:
        end = start + window_frames
        if end <= T:
            s = spec_cft[:, :, start:end]
            l = labels_tgm[start:end]
        else:
            pad = window_frames - (T - start)
            s = np.concatenate([spec_cft[:, :, start:], np.zeros((C, F, pad), spec_cft.dtype)], axis=2)
            lp = np.zeros((pad,) + labels_tgm.shape[1:], labels_tgm.dtype)
            lp[:, :, M - 1] = 1.0
            l = np.concatenate([labels_tgm[start:], lp], axis=0)
        out.append((s.transpose(2, 0, 1), l, start, min(end, T)))
        start += hop_frames
    return out
