#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (read-only /root/reference) on the synthetic
cases of cases.py.  Run once in the build container:  python tests/golden/make_golden.py

The reference is imported from a scratch copy (its modules mkdir next to themselves at import time,
SURVEY.md §0); matplotlib/librosa (absent here, unused on the path) are stubbed for
smrl_seld_gaussian.py; torchaudio.load (needs torchcodec, absent) is replaced by a synthetic loader for
the SELDDataset case.  Nothing from the reference is copied into the repo — only its outputs.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

REF = os.environ.get("SELD_REFERENCE", "/root/reference")


def import_reference():
    scratch = tempfile.mkdtemp(prefix="seld_ref_")
    for f in os.listdir(REF):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, f), scratch)
    sys.path.insert(0, scratch)
    os.chdir(scratch)  # smrl_seld_gaussian creates gaussian/logs under the CWD
    class _Stub(types.ModuleType):  # absent, unused-on-the-path plotting/audio libs
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return _Stub(f"{self.__name__}.{item}")

        def __call__(self, *a, **k):
            return None

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "librosa", "librosa.display",
                 "seaborn"):
        sys.modules.setdefault(name, _Stub(name))
    import dataset  # noqa
    import smrl_seld_gaussian  # noqa
    import utils  # noqa
    return dataset, smrl_seld_gaussian, utils


def make_losses():
    """losses.npz: the reference's SMRSELDLoss (loss.py) on the seeded cases of cases.LOSS_CASES — the live class losses
    (loss.py:27-54) and the two dormant terms (aiur_loss :56-88, converging_localization_loss :90-146), float32 like the
    trainer computes them, plus float64 runs of the same code as the accuracy yardstick."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "loss.py"))
    ref_loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_loss)
    out = {}
    w = torch.from_numpy((np.random.default_rng(77).random(cases.LOSS_M) + 0.5).astype(np.float32))
    out["ce_weights"] = w.numpy()
    for name in cases.LOSS_CASES:
        z, y, I, J = cases.make_loss_case(name)
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            zt, yt = torch.from_numpy(z).to(dt), torch.from_numpy(y).to(dt)
            crit = ref_loss.SMRSELDLoss("mse", grid_size=(I, J))
            critw = ref_loss.SMRSELDLoss("ce", grid_size=(I, J), class_weights=w.to(dt))
            p = torch.softmax(zt, dim=-1)  # what the commented-out lines of forward pass to the two terms (loss.py:158)
            out[f"{name}/{tag}"] = np.array([float(crit.class_mse_loss(zt, yt)), float(crit.class_ce_loss(zt, yt)),
                                             float(critw.class_ce_loss(zt, yt)), float(crit.aiur_loss(p, yt)),
                                             float(crit.converging_localization_loss(p, yt))], dtype=np.float64)
        if name == "grid6x12":  # the gradient of the differentiable dormant term, through the reference's own autograd
            zt = torch.from_numpy(z).double().requires_grad_(True)
            crit = ref_loss.SMRSELDLoss("mse", grid_size=(I, J))
            crit.converging_localization_loss(torch.softmax(zt, dim=-1), torch.from_numpy(y).double()).backward()
            out[f"{name}/cl_grad_f64"] = zt.grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
    print("losses.npz:", len(out), "arrays")


def main():
    torch.set_num_threads(1)
    if "--only-losses" in sys.argv:
        return make_losses()
    cases.write_csvs()
    make_losses()
    dataset, gauss, utils = import_reference()

    # ---- features -------------------------------------------------------------------------
    out = {}
    for name, (kind, n, seed) in cases.AUDIO_CASES.items():
        x = cases.make_audio(kind, n, seed)
        out[f"{name}/sha"] = np.frombuffer(cases.sha(x).encode(), dtype=np.uint8)
        for n_fft in cases.N_FFTS:
            y = dataset.audio_to_mel_spectrogram(torch.from_numpy(x), cases.SR, n_fft=n_fft,
                                                 hop_length=cases.HOP, n_mels=cases.N_MELS)
            assert y.dtype == torch.float32 and y.shape == (4, 64, 1 + n // cases.HOP)
            out[f"{name}/logmel_{n_fft}"] = y.numpy()
    # BASELINE configs[0]: one 60 s 4-channel clip through the reference; every CONFIG0_STRIDE-th frame is kept
    kind, n, seed = cases.CONFIG0
    x = cases.make_audio(kind, n, seed)
    out["config0/sha"] = np.frombuffer(cases.sha(x).encode(), dtype=np.uint8)
    for n_fft in cases.N_FFTS:
        y = dataset.audio_to_mel_spectrogram(torch.from_numpy(x), cases.SR, n_fft=n_fft, hop_length=cases.HOP,
                                             n_mels=cases.N_MELS)
        assert y.shape == (4, 64, 1 + n // cases.HOP)
        out[f"config0/logmel_{n_fft}_strided"] = y.numpy()[:, :, ::cases.CONFIG0_STRIDE].copy()
    # channel counts other than 4 (the reference only warns in load_audio)
    for ch in (1, 2, 3, 6):
        x = cases.make_audio("noise", 4800, 100 + ch, channels=ch)
        out[f"noise_ch{ch}/sha"] = np.frombuffer(cases.sha(x).encode(), dtype=np.uint8)
        out[f"noise_ch{ch}/logmel_1024"] = dataset.audio_to_mel_spectrogram(
            torch.from_numpy(x), cases.SR, n_fft=1024, hop_length=480, n_mels=64).numpy()
    import torchaudio
    for n_fft in cases.N_FFTS:
        out[f"fb_{n_fft}"] = torchaudio.functional.melscale_fbanks(
            n_fft // 2 + 1, 0.0, float(cases.SR // 2), cases.N_MELS, cases.SR, norm=None, mel_scale="htk").numpy()
        out[f"win_{n_fft}"] = torch.hann_window(n_fft).numpy()
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)
    print("features.npz:", len(out), "arrays")

    # ---- labels ---------------------------------------------------------------------------
    out = {}
    for name, (_csv, n) in cases.LABEL_CASES.items():
        dur = n / cases.SR
        lab, I, J = dataset.metadata_to_labels(cases.csv_path(name), dur, sample_rate=cases.SR,
                                               I=18, J=36, cell_size_deg=10, num_classes=14)
        lab = lab.numpy()
        assert set(np.unique(lab).tolist()) <= {0.0, 1.0} and (I, J) == (18, 36)
        out[f"{name}/point_bits"] = cases.pack_labels(lab)
        out[f"{name}/shape"] = np.array(lab.shape)
        if name in cases.GAUSS_SEEDS:
            np.random.seed(cases.GAUSS_SEEDS[name])
            lab, I, J = gauss.augment_with_gaussian_noise(cases.csv_path(name), dur, sample_rate=cases.SR,
                                                          I=18, J=36, cell_size_deg=10, num_classes=14)
            lab = lab.numpy()
            assert set(np.unique(lab).tolist()) <= {0.0, 1.0}
            out[f"{name}/region_bits"] = cases.pack_labels(lab)
            # non-default sigmas
            np.random.seed(cases.GAUSS_SEEDS[name] + 100)
            lab, _, _ = gauss.augment_with_gaussian_noise(cases.csv_path(name), dur, I=18, J=36,
                                                          cell_size_deg=10, num_classes=14,
                                                          sigma_azimuth=12.5, sigma_elevation=3.0)
            out[f"{name}/region_s12.5_3_bits"] = cases.pack_labels(lab.numpy())
    # polar_to_grid over the whole integer domain + the notebook known-answer
    az = np.arange(-400, 401)
    el = np.arange(-200, 201)
    out["polar/j"] = np.array([utils.polar_to_grid(int(a), 0, I=18, J=36)[1] for a in az])
    out["polar/i"] = np.array([utils.polar_to_grid(0, int(e), I=18, J=36)[0] for e in el])
    assert utils.polar_to_grid(-98, -16, I=18, J=36) == (7, 8)  # SMR_SELD_2.ipynb:751
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **out)
    print("labels.npz:", len(out), "arrays")

    # ---- SELDDataset windows (dataset.py:167-330) with the reference's default n_fft = 960 -----
    files = {"synthetic://a": ("noise", 97440, 31), "synthetic://b": ("noise", 60000, 32)}

    def fake_load_audio(path):
        kind, n, seed = files[path]
        return torch.from_numpy(cases.make_audio(kind, n, seed)), cases.SR

    dataset.load_audio = fake_load_audio
    ds = dataset.SELDDataset(list(files), [cases.csv_path("edges"), cases.csv_path("floatcol")])
    out = {
        "n_windows": np.array(len(ds)),
        "total_frames": np.array(ds.total_frames),
        "starts": np.array([w["start_frame"] for w in ds.windows]),
        "ends": np.array([w["end_frame"] for w in ds.windows]),
        "IJ": np.array([ds.I, ds.J, ds.total_cells, ds.window_length_frames, ds.hop_length_frames]),
        "concat_spec": ds.concatenated_spectrograms.numpy(),
        "concat_label_bits": cases.pack_labels(ds.concatenated_labels.numpy()),
    }
    for k in (0, len(ds) - 2, len(ds) - 1):
        s, l = ds[k]
        assert s.shape == (250, 4, 64) and l.shape == (250, 648, 14)
        out[f"win{k}/spec"] = s.contiguous().numpy()
        out[f"win{k}/label_bits"] = cases.pack_labels(l.numpy())
    np.savez_compressed(os.path.join(HERE, "windows.npz"), **out)
    print("windows.npz: n_windows", len(ds), "total_frames", ds.total_frames)


if __name__ == "__main__":
    main()
