"""GPU parity, bit-exact: label kernels (through the C ABI) against the golden outputs of the reference's
metadata_to_labels / augment_with_gaussian_noise, the windowing of SELDDataset and the on-device loader."""
import math

import numpy as np
import pytest
import torch

import cases
from oracle import features as of
from oracle import labels as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import seld_b200
    return seld_b200


@pytest.mark.parametrize("name", list(cases.LABEL_CASES))
def test_point_labels_bit_exact(sb, golden_labels, name):
    _csv, n = cases.LABEL_CASES[name]
    lab, I, J = sb.metadata_to_labels(cases.csv_path(name), n / cases.SR, sample_rate=cases.SR, I=18, J=36,
                                      cell_size_deg=10, num_classes=14)
    assert (I, J) == (18, 36) and not lab.is_cuda and lab.dtype == torch.float32
    assert tuple(lab.shape) == tuple(golden_labels[f"{name}/shape"])
    want = cases.unpack_labels(golden_labels[f"{name}/point_bits"], lab.shape)
    assert torch.equal(lab, torch.from_numpy(want))


@pytest.mark.parametrize("name", list(cases.GAUSS_SEEDS))
def test_region_labels_bit_exact(sb, golden_labels, name):
    _csv, n = cases.LABEL_CASES[name]
    shape = tuple(golden_labels[f"{name}/shape"])
    np.random.seed(cases.GAUSS_SEEDS[name])
    lab, _, _ = sb.augment_with_gaussian_noise(cases.csv_path(name), n / cases.SR, sample_rate=cases.SR, I=18, J=36,
                                               cell_size_deg=10, num_classes=14)
    assert torch.equal(lab, torch.from_numpy(cases.unpack_labels(golden_labels[f"{name}/region_bits"], shape)))
    np.random.seed(cases.GAUSS_SEEDS[name] + 100)
    lab, _, _ = sb.augment_with_gaussian_noise(cases.csv_path(name), n / cases.SR, I=18, J=36, cell_size_deg=10,
                                               num_classes=14, sigma_azimuth=12.5, sigma_elevation=3.0, device="cuda")
    assert lab.is_cuda
    assert torch.equal(lab.cpu(), torch.from_numpy(cases.unpack_labels(golden_labels[f"{name}/region_s12.5_3_bits"], shape)))


def test_region_labels_random_centres_vs_oracle(sb, tmp_path):
    """Many random sources / sigmas, other grids: GPU float64 region test == the line-by-line restatement."""
    rng = np.random.default_rng(7)
    for trial, (I, J, cs) in enumerate([(18, 36, 10), (9, 18, 20), (36, 72, 5), (12, 24, 15)]):
        rows = [(int(f), int(rng.integers(0, 13)), int(rng.integers(0, 3)), int(rng.integers(-180, 181)),
                 int(rng.integers(-90, 91))) for f in range(30) for _ in range(3)]
        p = tmp_path / f"r{trial}.csv"
        p.write_text("\n".join(",".join(map(str, r)) for r in rows) + "\n")
        sa, se = float(rng.uniform(1, 30)), float(rng.uniform(1, 30))
        np.random.seed(trial)
        got, _, _ = sb.augment_with_gaussian_noise(str(p), 3.0, I=I, J=J, cell_size_deg=cs, sigma_azimuth=sa,
                                                   sigma_elevation=se)
        np.random.seed(trial)
        want, _, _ = ol.augment_with_gaussian_noise(str(p), 3.0, I=I, J=J, cell_size_deg=cs, sigma_azimuth=sa,
                                                    sigma_elevation=se)
        assert np.array_equal(got.numpy(), want)


def test_full_minute_labels_property(sb, tmp_path):
    """BASELINE size (60 s -> (3000, 648, 14), 108.9 MB): every (t, cell) holds background XOR events, and
    the painted set equals the events expanded on the host."""
    rng = np.random.default_rng(0)
    rows = [(f, int(rng.integers(0, 13)), s, int(rng.integers(-180, 181)), int(rng.integers(-90, 91)))
            for f in range(600) for s in range(int(rng.integers(0, 4)))]
    p = tmp_path / "minute.csv"
    p.write_text("\n".join(",".join(map(str, r)) for r in rows) + "\n")
    lab, _, _ = sb.metadata_to_labels(str(p), 60.0, I=18, J=36, device="cuda")
    assert tuple(lab.shape) == (3000, 648, 14)
    ev_sum = lab[:, :, :13].sum(-1)
    assert torch.equal((ev_sum > 0).float() + lab[:, :, 13], torch.ones_like(ev_sum))
    events, T = sb.labels.point_events(str(p), 60.0, 18, 36)
    want = torch.zeros((3000, 648, 14), dtype=torch.bool)
    for r0, r1, c, cell in events:
        want[r0:r1, cell, c] = True
    assert torch.equal(lab[:, :, :13].cpu() > 0, want[:, :, :13])


def _fake_loader(files):
    def load(path):
        kind, n, seed = files[path]
        return torch.from_numpy(cases.make_audio(kind, n, seed)), cases.SR
    return load


FILES = {"synthetic://a": ("noise", 97440, 31), "synthetic://b": ("noise", 60000, 32)}


@pytest.mark.parametrize("resident", ["cpu", "cuda"])
def test_dataset_windows_vs_reference(sb, golden_windows, resident):
    g = golden_windows
    ds = sb.SELDDataset(list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")], audio_loader=_fake_loader(FILES),
                        resident=resident)
    assert len(ds) == int(g["n_windows"]) == math.ceil(int(g["total_frames"]) / 50)
    assert ds.total_frames == int(g["total_frames"])
    assert [ds.I, ds.J, ds.total_cells, ds.window_length_frames, ds.hop_length_frames] == g["IJ"].tolist()
    assert [w["start_frame"] for w in ds.windows] == g["starts"].tolist()
    assert [w["end_frame"] for w in ds.windows] == g["ends"].tolist()
    assert tuple(ds.concatenated_spectrograms.shape) == g["concat_spec"].shape
    assert np.abs(ds.concatenated_spectrograms.cpu().numpy() - g["concat_spec"]).max() <= 1e-3
    T = ds.total_frames
    assert torch.equal(ds.concatenated_labels.cpu(), torch.from_numpy(cases.unpack_labels(g["concat_label_bits"], (T, 648, 14))))
    for k in (0, len(ds) - 2, len(ds) - 1):
        s, l = ds[k]
        assert tuple(s.shape) == (250, 4, 64) and tuple(l.shape) == (250, 648, 14)
        assert s.is_cuda == (resident == "cuda")
        assert np.abs(s.cpu().numpy() - g[f"win{k}/spec"]).max() <= 1e-3
        assert torch.equal(l.cpu(), torch.from_numpy(cases.unpack_labels(g[f"win{k}/label_bits"], (250, 648, 14))))
    # padded tail of the last window: exact zeros / background
    s, l = ds[len(ds) - 1]
    n_real = ds.windows[-1]["end_frame"] - ds.windows[-1]["start_frame"]
    assert (s[n_real:] == 0).all() and (l[n_real:, :, 13] == 1).all() and (l[n_real:, :, :13] == 0).all()


def test_dataset_works_with_torch_dataloader(sb):
    ds = sb.SELDDataset(list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")], audio_loader=_fake_loader(FILES))
    dl = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=2, pin_memory=True)
    spec, lab = next(iter(dl))
    assert tuple(spec.shape) == (4, 250, 4, 64) and tuple(lab.shape) == (4, 250, 648, 14)  # SMR_SELD_2.ipynb cell 20
    assert dl.dataset.I == 18 and dl.dataset.J == 36 and dl.dataset.total_cells == 648


@pytest.mark.parametrize("label_mode", ["dense", "compact"])
def test_device_loader_matches_dataset_items(sb, label_mode):
    ref = sb.SELDDataset(list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")], audio_loader=_fake_loader(FILES))
    ds = sb.SELDDataset(list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")], audio_loader=_fake_loader(FILES),
                        resident="cuda", labels=label_mode, feature_type="foa_iv")
    dl = sb.DeviceLoader(ds, batch_size=3)
    assert len(dl) == math.ceil(len(ds) / 3) and dl.dataset is ds
    k = 0
    for spec, lab in dl:
        assert spec.is_cuda and lab.is_cuda and spec.shape[1:] == (250, 7, 64) and lab.shape[1:] == (250, 648, 14)
        for i in range(spec.shape[0]):
            s_ref, l_ref = ref[k]
            assert torch.equal(spec[i, :, :4].cpu(), s_ref)   # log-mel channels identical to the 4-channel dataset
            assert torch.equal(lab[i].cpu(), l_ref)
            k += 1
    assert k == len(ds)


def test_gaussian_dataset_compact_equals_dense(sb):
    args = (list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")])
    np.random.seed(5)
    a = sb.SELDDataset(*args, use_gaussian_augmentation=True, audio_loader=_fake_loader(FILES), resident="cuda")
    np.random.seed(5)
    b = sb.SELDDataset(*args, use_gaussian_augmentation=True, audio_loader=_fake_loader(FILES), resident="cuda",
                       labels="compact")
    for k in range(len(a)):
        assert torch.equal(a[k][1], b[k][1])
    # and the dense one equals the oracle per file
    np.random.seed(5)
    l0, _, _ = ol.augment_with_gaussian_noise(cases.csv_path("edges"), 97440 / 24000, I=18, J=36)
    l1, _, _ = ol.augment_with_gaussian_noise(cases.csv_path("floatcol"), 60000 / 24000, I=18, J=36)
    want = np.concatenate([l0[:202], l1[:125]])
    assert np.array_equal(a.concatenated_labels.cpu().numpy(), want)


def test_scaler_stats_apply(sb):
    ds = sb.SELDDataset(list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")], audio_loader=_fake_loader(FILES),
                        resident="cuda", feature_type="foa_iv", compute_stats=True)
    f = ds._features_tcf.double().cpu().numpy().reshape(ds.total_frames, -1)
    sc = sb.FeatureScaler(7 * 64, "cuda")
    sc.merge(ds.stats, ds.total_frames)
    sc.sync()  # no process group: no-op
    mean, std = sc.finalize()
    c, s, ss = of.scaler_stats(ds._features_tcf.cpu().numpy())
    m_ref, s_ref = of.scaler_mean_std(c, s, ss)
    assert np.allclose(mean.cpu().numpy(), m_ref, rtol=1e-6, atol=1e-6)
    assert np.allclose(std.cpu().numpy(), s_ref, rtol=1e-5, atol=1e-6)
    sc2 = ds.scaler()  # the one-call form (merge + all-reduce + finalize)
    assert torch.equal(sc2.mean, mean) and torch.equal(sc2.std, std)
    x = ds._features_tcf.clone()
    sc.apply(x)
    want = (f - m_ref) / s_ref
    assert np.abs(x.cpu().numpy().reshape(ds.total_frames, -1) - want).max() <= 1e-4
