"""CPU: host side of the label path (CSV parse, event tables, wrap-around, RNG draws) against the golden
vectors of the real reference.  The dense painting itself runs on the GPU (tests/test_labels_gpu.py); here the
compact events are expanded by a few lines of numpy so the host logic is covered without a device."""
import numpy as np
import torch
import pytest

import cases
from oracle import labels as ol


@pytest.fixture(scope="module")
def L():
    import seld_b200
    return seld_b200.labels


def expand_point_events(events, T, cells=648, M=14):
    lab = np.zeros((T, cells, M), np.float32)
    lab[:, :, M - 1] = 1.0
    for r0, r1, c, cell in events:      # pass 0: clear background
        lab[r0:r1, cell, M - 1] = 0.0
    for r0, r1, c, cell in events:      # pass 1: set classes
        lab[r0:r1, cell, c] = 1.0
    return lab


@pytest.mark.parametrize("name", list(cases.LABEL_CASES))
def test_point_events_expand_to_reference_labels(L, golden_labels, name):
    _csv, n = cases.LABEL_CASES[name]
    events, T = L.point_events(cases.csv_path(name), n / cases.SR, 18, 36, 14)
    assert events.dtype == np.int32 and events.shape[1] == 4
    assert (T, 648, 14) == tuple(golden_labels[f"{name}/shape"])
    lab = expand_point_events(events, T)
    assert (cases.pack_labels(lab) == golden_labels[f"{name}/point_bits"]).all()


def test_polar_to_grid_matches_reference(L, golden_labels):
    az = np.arange(-400, 401)
    el = np.arange(-200, 201)
    assert (np.array([L.polar_to_grid(int(a), 0, I=18, J=36)[1] for a in az]) == golden_labels["polar/j"]).all()
    assert (np.array([L.polar_to_grid(0, int(e), I=18, J=36)[0] for e in el]) == golden_labels["polar/i"]).all()
    cells = L._cells_of(np.zeros_like(az), np.zeros_like(az), 18, 36)
    assert (cells == 9 * 36 + 18).all()
    j = L._cells_of(az, np.full_like(az, -90), 18, 36)
    assert (j == golden_labels["polar/j"]).all()
    assert L.polar_to_grid(-98, -16, I=18, J=36) == (7, 8)  # SMR_SELD_2.ipynb:751
    with pytest.raises(ValueError):
        L.polar_to_grid(0, 0)


@pytest.mark.parametrize("name", list(cases.GAUSS_SEEDS))
def test_region_events_centres_follow_reference_rng(L, name):
    _csv, n = cases.LABEL_CASES[name]
    np.random.seed(cases.GAUSS_SEEDS[name])
    events, centres, T = L.region_events(cases.csv_path(name), n / cases.SR, 18, 36, 14)
    # same seed through the oracle's line-by-line restatement
    np.random.seed(cases.GAUSS_SEEDS[name])
    df, rows = ol._rows(cases.csv_path(name))
    noise = ol.draw_source_noise(df, 5.0, 5.0)
    assert (events[:, 3] == -1).all() and centres.shape == (len(events), 2)
    k = 0
    for (f, c, s, az, el) in rows:
        start, end = f * 5, min(f * 5 + 5, T)
        if start >= end:
            continue
        assert tuple(events[k, :3]) == (start, end, c)
        assert centres[k, 0] == az + noise[(c, s)][0] and centres[k, 1] == el + noise[(c, s)][1]
        k += 1
    assert k == len(events)
    # both consumed the global RNG identically
    a = np.random.normal()
    np.random.seed(cases.GAUSS_SEEDS[name])
    L.region_events(cases.csv_path(name), n / cases.SR, 18, 36, 14)
    assert np.random.normal() == a


def test_total_frames_python_float_expression(L):
    assert L.total_frames_of(97440 / 24000) == 202      # != 97440 // 480 == 203
    assert L.total_frames_of(2_145_600 / 24000) == 4470  # SMR_SELD_2.ipynb:663
    assert L.total_frames_of(1_440_000 / 24000) == 3000


def test_errors_match_reference(L, tmp_path):
    import pandas as pd
    p = tmp_path / "empty.csv"
    p.write_text("")
    with pytest.raises(pd.errors.EmptyDataError):
        L.point_events(str(p), 1.0, 18, 36)
    with pytest.raises(ValueError):
        L._grid(None, None, None)
    bad = tmp_path / "badclass.csv"
    bad.write_text("0,14,0,0,0\n")
    with pytest.raises(IndexError):
        L.point_events(str(bad), 1.0, 18, 36)
    late = tmp_path / "late.csv"
    late.write_text("999,99,0,0,0\n")       # row past the audio end never indexes the tensor -> no error
    ev, T = L.point_events(str(late), 1.0, 18, 36)
    assert len(ev) == 0 and T == 50
    neg = tmp_path / "neg.csv"
    neg.write_text("-100,1,0,0,0\n")        # start frame -500 < -T -> IndexError like labels[-500]
    with pytest.raises(IndexError):
        L.point_events(str(neg), 1.0, 18, 36)
    short = tmp_path / "short.csv"
    short.write_text("0,1,0,5\n")
    with pytest.raises(IndexError):
        L.point_events(str(short), 1.0, 18, 36)


def test_wav_roundtrip(tmp_path):
    import seld_b200
    x = cases.make_audio("int16", 4800, 3)
    p = tmp_path / "a.wav"
    seld_b200.audio_io.write_wav_pcm16(str(p), x, 24000)
    y, sr = seld_b200.load_audio(str(p))
    assert sr == 24000 and y.shape == (4, 4800) and y.dtype.is_floating_point
    assert np.array_equal(y.numpy(), x)  # int16-quantised input survives exactly (value / 32768)
    # the dataset's ingest form keeps the int16 samples (converted inside the feature kernel)
    z, sr2 = seld_b200.audio_io.load_audio_pcm16(str(p))
    assert sr2 == 24000 and z.dtype == torch.int16 and z.shape == (4, 4800)
    assert np.array_equal(z.numpy().astype(np.float32) / 32768.0, x)
