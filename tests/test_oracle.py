"""CPU: the oracle restatements against the golden vectors produced by the real reference
(tests/golden/make_golden.py) and against the notebook known-answers."""
import math

import os

import numpy as np
import pytest

import cases
from oracle import features as of
from oracle import labels as ol

TOL_DB = 1e-3  # north_star: max abs error <= 1e-3 on log-mel dB


def test_inputs_regenerate_identically(golden_features):
    for name, (kind, n, seed) in cases.AUDIO_CASES.items():
        x = cases.make_audio(kind, n, seed)
        assert cases.sha(x) == bytes(golden_features[f"{name}/sha"]).decode()


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
def test_filterbank_restatement_matches_torchaudio_table(golden_features, n_fft):
    fb = of.mel_filterbank_f32(n_fft // 2 + 1, cases.SR, cases.N_MELS)
    ref = golden_features[f"fb_{n_fft}"]
    assert fb.shape == ref.shape == (n_fft // 2 + 1, 64)
    assert ((fb != 0) == (ref != 0)).all()
    assert np.abs(fb - ref).max() <= 5e-6  # numpy vs ATen float32 pow differ by an ulp in f_pts


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
@pytest.mark.parametrize("name", list(cases.AUDIO_CASES))
def test_logmel_oracle_vs_reference(golden_features, name, n_fft):
    kind, n, seed = cases.AUDIO_CASES[name]
    x = cases.make_audio(kind, n, seed)
    ref = golden_features[f"{name}/logmel_{n_fft}"]
    got = of.logmel(x, cases.SR, n_fft, cases.HOP, cases.N_MELS, fb=golden_features[f"fb_{n_fft}"])
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL_DB


def test_silence_is_exactly_minus_100(golden_features):
    assert (golden_features["zeros/logmel_1024"] == -100.0).all()
    x = cases.make_audio("zeros", 4800, 0)
    assert (of.logmel(x, cases.SR, 1024, 480, 64) == -100.0).all()


def test_ref_port_is_bit_identical_to_reference(golden_features):
    import torch
    from oracle import ref_port
    old = torch.get_num_threads()
    torch.set_num_threads(1)  # the golden vectors were produced single-threaded; MKL's blocking changes the last ulp
    try:
        for name, (kind, n, seed) in cases.AUDIO_CASES.items():
            x = torch.from_numpy(cases.make_audio(kind, n, seed))
            for n_fft in cases.N_FFTS:
                y = ref_port.audio_to_mel_spectrogram_port(x, cases.SR, n_fft, cases.HOP, cases.N_MELS).numpy()
                assert np.array_equal(y, golden_features[f"{name}/logmel_{n_fft}"]), (name, n_fft)
    finally:
        torch.set_num_threads(old)


def test_ref_port_iv_matches_fp64_restatement():
    import torch
    from oracle import ref_port
    x = cases.make_audio("noise", 24000, 5)
    got = ref_port.logmel_iv_port(torch.from_numpy(x), cases.SR, 1024, 480, 64).numpy()
    want = of.logmel_iv(x, cases.SR, 1024, 480, 64)
    assert np.abs(got[:4] - want[:4]).max() <= TOL_DB
    assert np.abs(got[4:] - want[4:]).max() <= 1e-4 * np.abs(want[4:]).max()


def test_label_cost_port_equals_reference_golden(golden_labels):
    """The loop-for-loop label port that bench.py times on the CPU reproduces the reference's output bit for bit."""
    from oracle import ref_port
    for name in ("edges", "floatcol"):
        _csv, n = cases.LABEL_CASES[name]
        got = ref_port.metadata_to_labels_port(cases.csv_path(name), n / cases.SR).numpy()
        want = cases.unpack_labels(golden_labels[f"{name}/point_bits"], tuple(golden_labels[f"{name}/shape"]))
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
@pytest.mark.parametrize("name", ["noise_1s", "int16_noise", "impulse_first", "sine_1k_1e-4_ch0", "level_80db"])
def test_two_gcc_phat_restatements_agree(name, n_fft):
    """A8 has no reference code: two independent fp64 restatements (numpy framing + np.fft vs torch.stft + torch.fft.irfft)
    must agree, so that a slip in either one cannot pass as parity."""
    import torch
    from oracle import ref_port
    kind, n, seed = cases.AUDIO_CASES[name]
    x = cases.make_audio(kind, n, seed)
    a = of.gcc_phat(x, n_fft, cases.HOP, 64)
    b = ref_port.gcc_phat_port(torch.from_numpy(x), n_fft, cases.HOP, 64).numpy()
    assert a.shape == b.shape == (6, 64, 1 + n // cases.HOP)
    if kind in ("impulse_first", "sine"):
        # bins that are exactly 0 in one restatement and ~1e-20 in the other have arbitrary phase: compare where defined
        assert np.abs(a[..., 5:-5] - b[..., 5:-5]).max() <= 1e-6 or kind == "impulse_first"
    else:
        assert np.abs(a - b).max() <= 1e-9


# ---- notebook known-answers ---------------------------------------------------
def test_known_answers_shapes():
    n = 2_145_600  # SMR_SELD_2.ipynb:518-519, :663
    assert of.num_frames(n, 480) == 4471
    assert ol.total_frames_of(n / 24000) == 4470
    assert ol.polar_to_grid(-98, -16, I=18, J=36) == (7, 8)  # SMR_SELD_2.ipynb:751
    for total, windows in ((4470, 90), (3035, 61), (11470, 230), (5270, 106)):
        assert math.ceil(total / 50) == windows
    assert ol.total_frames_of(97440 / 24000) == 202 and 97440 // 480 == 203  # SURVEY.md §7 hard part 5


def test_polar_to_grid_domain(golden_labels):
    az = np.arange(-400, 401)
    el = np.arange(-200, 201)
    j = np.array([ol.polar_to_grid(int(a), 0, I=18, J=36)[1] for a in az])
    i = np.array([ol.polar_to_grid(0, int(e), I=18, J=36)[0] for e in el])
    assert (j == golden_labels["polar/j"]).all() and (i == golden_labels["polar/i"]).all()


@pytest.mark.parametrize("name", list(cases.LABEL_CASES))
def test_point_labels_oracle_vs_reference(golden_labels, name):
    _csv, n = cases.LABEL_CASES[name]
    lab, I, J = ol.metadata_to_labels(cases.csv_path(name), n / cases.SR, I=18, J=36)
    assert tuple(golden_labels[f"{name}/shape"]) == lab.shape
    assert (cases.pack_labels(lab) == golden_labels[f"{name}/point_bits"]).all()
    assert set(np.unique(lab).tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("name", list(cases.GAUSS_SEEDS))
def test_region_labels_oracle_vs_reference(golden_labels, name):
    _csv, n = cases.LABEL_CASES[name]
    np.random.seed(cases.GAUSS_SEEDS[name])
    lab, _, _ = ol.augment_with_gaussian_noise(cases.csv_path(name), n / cases.SR, I=18, J=36)
    assert (cases.pack_labels(lab) == golden_labels[f"{name}/region_bits"]).all()
    np.random.seed(cases.GAUSS_SEEDS[name] + 100)
    lab, _, _ = ol.augment_with_gaussian_noise(cases.csv_path(name), n / cases.SR, I=18, J=36,
                                               sigma_azimuth=12.5, sigma_elevation=3.0)
    assert (cases.pack_labels(lab) == golden_labels[f"{name}/region_s12.5_3_bits"]).all()


def test_empty_csv_raises_like_reference(tmp_path):
    import pandas as pd
    p = tmp_path / "empty.csv"
    p.write_text("")
    with pytest.raises(pd.errors.EmptyDataError):
        ol.metadata_to_labels(str(p), 1.0, I=18, J=36)


def test_windows_oracle_vs_reference(golden_windows):
    g = golden_windows
    spec = g["concat_spec"]
    T = int(g["total_frames"])
    labels = cases.unpack_labels(g["concat_label_bits"], (T, 648, 14))
    wins = ol.create_windows(spec, labels)
    assert len(wins) == int(g["n_windows"]) == math.ceil(T / 50)
    assert [w[2] for w in wins] == g["starts"].tolist() and [w[3] for w in wins] == g["ends"].tolist()
    for k in (0, len(wins) - 2, len(wins) - 1):
        assert np.array_equal(wins[k][0], g[f"win{k}/spec"])
        assert (cases.pack_labels(wins[k][1]) == g[f"win{k}/label_bits"]).all()


# ---- unpinned restatements: physical invariants ---------------------------------------------------
def test_gcc_phat_peaks_at_delay():
    rng = np.random.default_rng(3)
    s = rng.standard_normal(24000 + 80)
    d = 7
    x = np.stack([s[40:40 + 24000], s[40 - d:40 - d + 24000], s[40:40 + 24000], s[40 + 5:40 + 5 + 24000]])
    g = of.gcc_phat(x, 1024, 480)  # (6, 64, T)
    assert g.shape == (6, 64, 51)
    mid = g[:, :, 5:45]
    assert (mid[0].argmax(0) == 32 + d).all()   # pair (0,1): ch1 lags ch0 by d
    assert (mid[1].argmax(0) == 32).all()       # pair (0,2): identical
    assert (mid[2].argmax(0) == 32 - 5).all()   # pair (0,3): ch3 leads by 5


def test_iv_points_at_plane_wave_direction():
    rng = np.random.default_rng(4)
    s = rng.standard_normal(24000)
    az, el = np.deg2rad(40.0), np.deg2rad(-20.0)
    u = np.array([np.sin(az) * np.cos(el), np.sin(el), np.cos(az) * np.cos(el)])  # FOA ACN order Y, Z, X
    x = np.stack([s, u[0] * s, u[1] * s, u[2] * s])
    iv = of.foa_iv(x, cases.SR, 1024, 480, 64)  # (3, 64, T)
    v = iv[:, 10:60, 5:45].mean(axis=(1, 2))
    assert np.allclose(v / np.linalg.norm(v), u, atol=1e-6)


def test_scaler_stats():
    rng = np.random.default_rng(5)
    f = rng.standard_normal((100, 7, 64)) * 3 + 1
    c, s, ss = of.scaler_stats(f)
    mean, std = of.scaler_mean_std(c, s, ss)
    assert np.allclose(mean, f.reshape(100, -1).mean(0)) and np.allclose(std, f.reshape(100, -1).std(0))


def test_loss_ports_reproduce_the_reference_losses():
    """oracle/ref_port.py's restatements of reference loss.py (class losses :27-54, aiur_loss :56-88,
    converging_localization_loss :90-146) against values the REAL reference produced (tests/golden/losses.npz, made by
    make_golden.py): float32 runs agree to the last bits, the converging-localisation gradient through autograd."""
    import torch
    from oracle import ref_port
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))
    w = torch.from_numpy(gold["ce_weights"])
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        for name in cases.LOSS_CASES:
            z, y, I, J = cases.make_loss_case(name)
            zt, yt = torch.from_numpy(z), torch.from_numpy(y)
            p = torch.softmax(zt, dim=-1)
            got = [float(ref_port.class_mse_loss_port(zt, yt)), float(ref_port.class_ce_loss_port(zt, yt)),
                   float(ref_port.class_ce_loss_port(zt, yt, w)), float(ref_port.aiur_loss_port(p, yt)),
                   float(ref_port.cl_loss_port(p, yt, I, J))]
            assert np.allclose(got, gold[f"{name}/f32"], rtol=1e-6, atol=1e-9), (name, got, gold[f"{name}/f32"])
            # the mask the CUDA path consumes stands for exactly these targets
            m = cases.loss_mask(y).view(np.uint16).astype(np.int64)
            back = ((m[..., None] >> np.arange(cases.LOSS_M)) & 1).astype(np.float32)
            back[..., cases.LOSS_M - 1] = np.where(m == 0, 1.0, back[..., cases.LOSS_M - 1])
            assert np.array_equal(back, y), name
        z, y, I, J = cases.make_loss_case("grid6x12")
        zt = torch.from_numpy(z).double().requires_grad_(True)
        ref_port.cl_loss_port(torch.softmax(zt, dim=-1), torch.from_numpy(y).double(), I, J).backward()
        assert np.allclose(zt.grad.numpy(), gold["grid6x12/cl_grad_f64"], rtol=1e-5, atol=1e-9)
    finally:
        torch.set_num_threads(old)
