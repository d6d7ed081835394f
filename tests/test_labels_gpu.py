"""GPU parity, bit-exact: label kernels (through the C ABI) against the golden outputs of the reference's
metadata_to_labels / augment_with_gaussian_noise, the windowing of SELDDataset and the on-device loader."""
import math
import os

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


@pytest.mark.parametrize("gaussian", [False, True])
def test_device_loader_shuffled_one_launch_batches(sb, gaussian):
    """Shuffled epochs through the one-launch batch kernel (device-side event selection): every batch equals the dense
    dataset's items of the same permutation, point and Gaussian-region labels, partial last batch, ring reuse."""
    args = (list(FILES), [cases.csv_path("edges"), cases.csv_path("floatcol")])
    np.random.seed(7)
    dense = sb.SELDDataset(*args, use_gaussian_augmentation=gaussian, audio_loader=_fake_loader(FILES), resident="cuda",
                           feature_type="foa_iv")
    np.random.seed(7)
    ds = sb.SELDDataset(*args, use_gaussian_augmentation=gaussian, audio_loader=_fake_loader(FILES), resident="cuda",
                        labels="compact", feature_type="foa_iv")
    for bs, drop in ((2, False), (4, True), (16, False)):
        dl = sb.DeviceLoader(ds, batch_size=bs, shuffle=True, drop_last=drop, generator=torch.Generator().manual_seed(3), depth=2)
        for epoch in range(2):
            order = torch.randperm(len(ds), generator=torch.Generator().manual_seed(3)) if epoch == 0 else None
            seen = 0
            for bi, (spec, lab) in enumerate(dl):
                if order is not None:
                    for i in range(spec.shape[0]):
                        s_ref, l_ref = dense[int(order[bi * bs + i])]
                        assert torch.equal(spec[i], s_ref) and torch.equal(lab[i], l_ref)
                seen += spec.shape[0]
            assert seen == (len(ds) // bs * bs if drop else len(ds)) and bi + 1 == len(dl)


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


# ---- clip sharding with exact global windows (SURVEY.md §8(e) nuance; reference dataset.py:259, :274-314) ----
SHARD_FILES = {"synthetic://a": ("noise", 97440, 31), "synthetic://b": ("noise", 60000, 32), "synthetic://c": ("noise", 30000, 33),
               "synthetic://d": ("noise", 72000, 34)}
SHARD_CSVS = ["edges", "floatcol", "weird", "basic"]


def _shard_worker(rank, world, port, label_mode, gaussian, q):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests", "golden")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # both ranks share cuda:0 here; NCCL needs one GPU per rank
    import seld_b200 as sb
    files, csvs = list(SHARD_FILES), [cases.csv_path(c) for c in SHARD_CSVS]
    mine_a, mine_m = sb.shard_files(files, csvs, rank, world)
    np.random.seed(11)  # (rank 0 draws the noise of its files first; the unsharded run below draws all in order)
    if gaussian and rank == 1:  # consume the draws of rank 0's files so that both runs see the same noise per source
        from seld_b200 import labels as L
        for a, m in zip(*sb.shard_files(files, csvs, 0, world)):
            L.region_events(m, SHARD_FILES[a][1] / 24000, 18, 36, 14)
    ds = sb.SELDDataset(mine_a, mine_m, use_gaussian_augmentation=gaussian, audio_loader=_fake_loader(SHARD_FILES), resident="cuda",
                        labels=label_mode, feature_type="foa_iv", distributed=True)
    # (numpy arrays: pickled by value; torch tensors would travel as file descriptors of a process that has exited)
    items = [(ds.first_window + k, ds[k][0].cpu().numpy(), np.packbits(ds[k][1].cpu().numpy() != 0)) for k in range(len(ds))]
    batches = []
    if label_mode == "compact":
        for spec, lab in sb.DeviceLoader(ds, batch_size=4):
            batches.append((spec.cpu().numpy().copy(), np.packbits(lab.cpu().numpy() != 0, axis=None).reshape(spec.shape[0], -1)))
    q.put((rank, ds.global_frame_offset, ds.halo_frames, items, batches))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("label_mode,gaussian", [("dense", False), ("compact", True)])
def test_sharded_dataset_windows_equal_the_unsharded_ones(sb, label_mode, gaussian):
    """Two ranks (two processes, gloo, one GPU): rank r's window g is bit-identical to window g of the dataset built from
    ALL files in one process — the reference's semantics (windows of the global concatenation straddle file and shard
    boundaries).  Features, point labels and Gaussian-region labels, through __getitem__ and through DeviceLoader."""
    import socket
    import torch.multiprocessing as mp
    files, csvs = list(SHARD_FILES), [cases.csv_path(c) for c in SHARD_CSVS]
    np.random.seed(11)
    full = sb.SELDDataset(files, csvs, use_gaussian_augmentation=gaussian, audio_loader=_fake_loader(SHARD_FILES), resident="cuda",
                          feature_type="foa_iv")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, label_mode, gaussian, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    for rank, off, halo, items, batches in res:
        assert halo <= 249  # a window that starts on the last frames of a shard reads up to window - 1 frames of the next
        for g, spec, lab_bits in items:
            s_ref, l_ref = full[g]
            l_ref = l_ref.cpu().numpy()
            assert set(np.unique(l_ref).tolist()) <= {0.0, 1.0}
            assert np.array_equal(spec, s_ref.cpu().numpy()) and np.array_equal(lab_bits, np.packbits(l_ref != 0)), (rank, g)
            seen.append(g)
        k = 0
        for spec, lab_bits in batches:
            for i in range(spec.shape[0]):
                assert np.array_equal(spec[i], items[k][1]) and np.array_equal(lab_bits[i], items[k][2])
                k += 1
    assert seen == list(range(len(full)))
    assert res[0][1] == 0 and res[0][2] > 0 and res[1][2] == 0  # rank 0 reads a halo from rank 1; the last rank pads


# ---- N4: class loss from compact targets (reference loss.py:27-54) ----
@pytest.mark.parametrize("gaussian", [False, True])
def test_compact_loss_equals_reference_loss_on_dense_labels(sb, gaussian):
    """softmax-MSE and (weighted) cross entropy, value and gradient, from the int16 class-set masks == the reference's
    formulas (restated in oracle/ref_port.py) on the dense labels of the same batches — including multi-hot cells, class-13
    events (weird.csv) and padded tail windows."""
    from oracle import ref_port
    files = {"synthetic://a": ("noise", 97440, 31), "synthetic://b": ("noise", 48480, 32), "synthetic://c": ("noise", 60000, 33)}
    csvs = [cases.csv_path("edges"), cases.csv_path("weird"), cases.csv_path("floatcol")]
    np.random.seed(3)
    ds = sb.SELDDataset(list(files), csvs, use_gaussian_augmentation=gaussian, audio_loader=_fake_loader(files), resident="cuda",
                        labels="compact", feature_type="foa_iv")
    dense = sb.DeviceLoader(ds, batch_size=4)
    masks = sb.DeviceLoader(ds, batch_size=4, targets="mask")
    g = torch.Generator().manual_seed(1)
    w = (torch.rand(14, generator=g) + 0.5).cuda()
    for (spec_d, lab), (spec_m, mask) in zip(dense, masks):
        assert torch.equal(spec_d, spec_m) and mask.dtype == torch.int16 and mask.shape == lab.shape[:3]
        # the mask IS the dense label: bit c <=> lab[..., c] == 1, 0 <=> one-hot background
        bits = (lab[..., :] != 0).to(torch.int32) * (1 << torch.arange(14, device="cuda", dtype=torch.int32))
        want_mask = bits.sum(-1)
        want_mask = torch.where(want_mask == (1 << 13), torch.zeros_like(want_mask), want_mask)
        got_mask = mask.to(torch.int32) & 0xffff
        only_bg_event = (got_mask == (1 << 13))  # a class-13 event alone: mask bit 13 set, dense row identical to background
        assert torch.equal(torch.where(only_bg_event, torch.zeros_like(got_mask), got_mask), want_mask)
        z = (3.0 * torch.randn(lab.shape, generator=g)).cuda()
        for loss_type, weights in (("mse", None), ("ce", None), ("ce", w)):
            z1 = z.clone().requires_grad_(True)
            ref = (ref_port.class_mse_loss_port(z1, lab) if loss_type == "mse" else ref_port.class_ce_loss_port(z1, lab, weights))
            (2.5 * ref).backward()
            # the same formulas in float64: torch's float32 mean over 2.6 M cells is itself only good to ~1e-5 relative,
            # the kernel accumulates in float64
            ref64 = (ref_port.class_mse_loss_port(z.double(), lab.double()) if loss_type == "mse"
                     else ref_port.class_ce_loss_port(z.double(), lab.double(), None if weights is None else weights.double()))
            z2 = z.clone().requires_grad_(True)
            crit = sb.CompactSMRSELDLoss(loss_type=loss_type, w_class=2.5, grid_size=(18, 36), class_weights=weights)
            total, breakdown = crit(z2, mask)
            total.backward()
            assert abs(breakdown[f"class_{loss_type}"] - float(ref64)) <= 2e-6 * max(1.0, abs(float(ref64)))
            assert abs(breakdown[f"class_{loss_type}"] - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
            assert torch.allclose(total.double(), 2.5 * ref64, rtol=2e-6, atol=1e-9)
            assert torch.allclose(z2.grad, z1.grad, rtol=1e-4, atol=1e-12 + 1e-5 * float(z1.grad.abs().max()))


def test_dataset_ingests_pcm16_wav_files_without_a_host_float_pass(sb, tmp_path):
    """SELDDataset's default loader keeps 16-bit PCM WAV files as int16 (half the PCIe bytes, files read by a thread pool);
    the kernel's x / 32768 is exact, so features, labels and windows equal those of the float32 ``load_audio`` path bit for
    bit — for the FOA, log-mel and MIC (converted on the device: the MIC kernel takes float32) feature types."""
    files, csvs = [], []
    for i, (n, seed, csv) in enumerate(((97440, 41, "edges"), (60000, 42, "floatcol"), (48480, 43, "weird"))):
        p = tmp_path / f"clip{i}.wav"
        sb.audio_io.write_wav_pcm16(str(p), cases.make_audio("int16", n, seed), cases.SR)
        files.append(str(p))
        csvs.append(cases.csv_path(csv))
    for ft in ("foa_iv", "logmel", "mic_gcc"):
        a = sb.SELDDataset(files, csvs, resident="cuda", labels="compact", feature_type=ft)
        b = sb.SELDDataset(files, csvs, resident="cuda", labels="compact", feature_type=ft, audio_loader=sb.load_audio)
        assert len(a) == len(b) and a.total_frames == b.total_frames
        assert torch.equal(a._features_tcf, b._features_tcf), ft
        for k in (0, len(a) - 1):
            (sa, la), (sb_, lb) = a[k], b[k]
            assert torch.equal(sa, sb_) and torch.equal(la, lb)


def test_losses_from_masks_match_the_reference_golden_values(sb):
    """All five loss terms of reference loss.py from logits + int16 class-set masks against the values the REAL reference
    computed on dense targets (tests/golden/losses.npz): the live softmax-MSE / cross entropy / weighted cross entropy and
    the dormant AIUR and converging-localisation terms (loss.py:56-146), incl. frames without events, class-13 events and
    multi-hot cells; the converging-localisation gradient against the reference's autograd."""
    from oracle import ref_port
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))
    w = torch.from_numpy(gold["ce_weights"]).cuda()
    for name in cases.LOSS_CASES:
        z, y, I, J = cases.make_loss_case(name)
        zc = torch.from_numpy(z).cuda()
        mask = torch.from_numpy(cases.loss_mask(y)).cuda()
        want = gold[f"{name}/f64"]
        for k, (lt, ww) in enumerate((("mse", None), ("ce", None), ("ce", w))):
            crit = sb.CompactSMRSELDLoss(loss_type=lt, grid_size=(I, J), class_weights=ww)
            _, br = crit(zc, mask)
            assert abs(br[f"class_{lt}"] - want[k]) <= 2e-6 * abs(want[k]), (name, lt, br, want[k])
        crit = sb.CompactSMRSELDLoss(loss_type="mse", w_class=1.5, w_aiur=0.25, w_cl=2.0, grid_size=(I, J), use_aux_terms=True)
        z1 = zc.clone().requires_grad_(True)
        total, br = crit(z1, mask)
        assert abs(br["aiur"] - want[3]) <= 1e-6, (name, br["aiur"], want[3])
        assert abs(br["cl"] - want[4]) <= 1e-5 * abs(want[4]) + 1e-9, (name, br["cl"], want[4])
        assert abs(float(total) - (1.5 * want[0] + 0.25 * want[3] + 2.0 * want[4])) <= 1e-5
        assert abs(float(crit.aiur_loss(zc, mask)) - want[3]) <= 1e-6
        # gradient of the converging-localisation term alone
        z2 = zc.clone().requires_grad_(True)
        (3.0 * crit.converging_localization_loss(z2, mask)).backward()
        z3 = torch.from_numpy(z).double().requires_grad_(True)
        (3.0 * ref_port.cl_loss_port(torch.softmax(z3, dim=-1), torch.from_numpy(y).double(), I, J)).backward()
        ref_g = z3.grad.float().cuda()
        assert torch.allclose(z2.grad, ref_g, rtol=1e-4, atol=1e-6 * float(ref_g.abs().max()))
        if name == "grid6x12":
            gg = 3.0 * torch.from_numpy(gold["grid6x12/cl_grad_f64"]).cuda()
            assert torch.allclose(z2.grad, gg, rtol=1e-4, atol=1e-6 * float(gg.abs().max()))
        # the total's gradient = class part + w_cl * cl part (AIUR is an argmax statistic: no gradient)
        total.backward()
        z4 = zc.clone().requires_grad_(True)
        sb.CompactSMRSELDLoss(loss_type="mse", w_class=1.5, grid_size=(I, J))(z4, mask)[0].backward()
        assert torch.allclose(z1.grad, z4.grad + (2.0 / 3.0) * z2.grad, rtol=1e-4, atol=1e-9)


def test_region_candidate_box_equals_exhaustive_test(sb):
    """The paint kernels evaluate the exact float64 region test only on a box of candidate cells around the centre.  Against
    the oracle's exhaustive loop over all I*J cells: centres on cell-centre boundaries (the <= comparisons decide), at the
    azimuth wrap, beyond the poles, out of range, with sigmas from 0 to larger than the sphere, on four grids."""
    rng = np.random.default_rng(42)
    lib, check = sb._lib.lib(), sb._lib.check
    stream = torch.cuda.current_stream().cuda_stream
    for (I, J) in ((18, 36), (9, 18), (36, 72), (12, 24)):
        centres, sig = [], []
        special = [-185.0, -180.0, -175.0, -5.0, 0.0, 5.0, 175.0, 180.0, 185.0, 355.0, -400.0, 720.0]
        for az in special:
            for el in (-95.0, -90.0, -85.0, -5.0, 0.0, 85.0, 90.0, 100.0, 200.0):
                centres.append((az, el))
        for _ in range(150):
            centres.append((float(rng.uniform(-400, 400)), float(rng.uniform(-200, 200))))
        for sa, se in ((5.0, 5.0), (0.0, 0.0), (2.5, 2.5), (2.4999, 7.5), (12.5, 3.0), (45.0, 45.0), (95.0, 50.0), (0.01, 30.0)):
            ev = np.array([[k, k + 1, k % 13, -1] for k in range(len(centres))], dtype=np.int32)
            ce = np.array(centres, dtype=np.float64)
            out = torch.empty((len(centres), I * J, 14), dtype=torch.float32, device="cuda")
            check(lib.seld_labels_fill(out.data_ptr(), out.shape[0], I * J, 14, stream), "fill")
            ev_d, ce_d = torch.from_numpy(ev).cuda(), torch.from_numpy(ce).cuda()
            check(lib.seld_labels_paint(out.data_ptr(), out.shape[0], I, J, 14, ev_d.data_ptr(), ce_d.data_ptr(), len(ev), sa, se,
                                        stream), "paint")
            got = out.cpu().numpy()
            for k, (az, el) in enumerate(centres):
                want = np.zeros(I * J, bool)
                want[ol.region_cells(az, el, sa, se, I, J)] = True
                assert np.array_equal(got[k, :, k % 13] == 1.0, want), (I, J, az, el, sa, se)
                assert np.array_equal(got[k, :, 13] == 0.0, want)
