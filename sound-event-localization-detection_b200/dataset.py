"""Drop-in for the reference's ``dataset.py`` (dataset.py:1-330): ``load_audio``, ``audio_to_mel_spectrogram``,
``metadata_to_labels``, ``load_files`` and ``SELDDataset`` with the same names, arguments, attributes and log
lines, plus the Gaussian-label switch of smrl_seld_gaussian.py:540 (``use_gaussian_augmentation``).

B200-first differences (none visible through the reference API):
  * features and labels of every file are produced by the CUDA kernels and written in place into the two
    concatenated tensors (no per-file tensors, no torch.cat);
  * the concatenated features are kept frame-major ``(sum T, C, F)`` — what ``__getitem__`` hands out
    (dataset.py:303) — ``concatenated_spectrograms`` is the ``(C, F, sum T)`` view of it;
  * ``resident="cuda"`` keeps everything in HBM (65 GB of labels for a 10 h corpus fit in 180 GB) and
    ``DeviceLoader`` assembles batches there; ``resident="cpu"`` (default, what unmodified main.py with
    DataLoader workers + pin_memory needs) copies the two tensors back once;
  * ``labels="compact"`` keeps only the event tables and paints label windows per batch.
"""
from __future__ import annotations

import logging
import os
from glob import glob
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib, labels as L
from .audio_io import load_audio, load_audio_pcm16  # noqa: F401  (load_audio re-exported like the reference's dataset.load_audio)
from .config import get_config
from .features import audio_to_mel_spectrogram, get_plan  # noqa: F401
from .labels import augment_with_gaussian_noise, metadata_to_labels, polar_to_grid  # noqa: F401

logger = logging.getLogger("SMR_SELD")

FEATURE_MODES = {"logmel": "logmel", "foa_iv": "logmel_iv", "logmel_iv": "logmel_iv", "mic_gcc": "logmel_gcc",
                 "logmel_gcc": "logmel_gcc"}


def load_files():
    """Reference dataset.py:121-165: sorted wav lists of the four dev folders with matching CSVs."""
    config = get_config()
    if config.USE_FULL_DATASET:
        sony_train_audio = sorted(glob(str(config.SONY_TRAIN_DIR / "*.wav")))
        tau_train_audio = sorted(glob(str(config.TAU_TRAIN_DIR / "*.wav")))
        sony_test_audio = sorted(glob(str(config.SONY_TEST_DIR / "*.wav")))
        tau_test_audio = sorted(glob(str(config.TAU_TEST_DIR / "*.wav")))

        def get_matching_metadata(audio_files, meta_dir):
            meta_files = []
            for audio_file in audio_files:
                meta_file = meta_dir / f"{Path(audio_file).stem}.csv"
                if meta_file.exists():
                    meta_files.append(str(meta_file))
                else:
                    raise FileNotFoundError(f"Metadata file not found: {meta_file}")
            return meta_files

        train_audio_files = sony_train_audio + tau_train_audio
        train_meta_files = (get_matching_metadata(sony_train_audio, config.SONY_TRAIN_META_DIR)
                            + get_matching_metadata(tau_train_audio, config.TAU_TRAIN_META_DIR))
        test_audio_files = sony_test_audio + tau_test_audio
        test_meta_files = (get_matching_metadata(sony_test_audio, config.SONY_TEST_META_DIR)
                           + get_matching_metadata(tau_test_audio, config.TAU_TEST_META_DIR))
    else:
        train_audio_files = [str(config.TRAIN_AUDIO_PATH)]
        train_meta_files = [str(config.TRAIN_META_PATH)]
        test_audio_files = [str(config.TEST_AUDIO_PATH)]
        test_meta_files = [str(config.TEST_META_PATH)]
    return train_audio_files, train_meta_files, test_audio_files, test_meta_files


def shard_window_plan(frames_per_rank, rank: int, window: int, hop: int):
    """Windows of the GLOBAL concatenation that a rank owns when the files are sharded by rank.

    The reference windows the concatenation of all files (dataset.py:259, :274-314): window g covers global frames
    [g*hop, g*hop + window) and straddles file — hence shard — boundaries.  Rank r owns the frames [off, off + T_r) and
    every window that STARTS in them.  Returns ``(off, local_starts, halo, first_window)``: the global offset of the
    rank's first frame, the owned windows' start frames relative to it, the number of frames past the end of the shard
    that those windows read from the following ranks (0 for the last non-empty rank: the reference pads there) and the
    global index of the first owned window."""
    frames_per_rank = [int(t) for t in frames_per_rank]
    off, T, total = sum(frames_per_rank[:rank]), frames_per_rank[rank], sum(frames_per_rank)
    first = -(-off // hop)  # ceil
    starts = list(range(first * hop, off + T, hop))
    halo = 0
    if starts:
        halo = max(0, min(starts[-1] + window, total) - (off + T))
    return off, [g - off for g in starts], halo, first


def exchange_shard_halo(head: torch.Tensor, head_events: np.ndarray, head_centres, n_frames: int, window: int, hop: int,
                        group=None):
    """The neighbour exchange of SURVEY.md §8(e): every rank offers the first ``min(T_r, window)`` frames of its
    features (``head``, (h, row_len)) with the events that touch them (rows relative to the rank's first frame); one
    ``all_gather`` of the padded heads (window x row_len floats per rank, ~0.45 MB) and one ``all_gather_object`` of the
    small event tables give every rank the halo of the following ranks.  Works on any backend (NCCL for CUDA tensors,
    gloo for CPU tensors in the tests).  Returns ``(plan, halo_feat (halo, row_len), halo_events, halo_centres)`` with the
    halo events' rows relative to this rank's first frame."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    row_len = head.shape[1]
    h = head.shape[0]
    assert h == min(n_frames, window)
    meta = [None] * world
    dist.all_gather_object(meta, {"T": int(n_frames), "ev": np.asarray(head_events), "ce": head_centres}, group=group)
    # (a gloo group moves CUDA tensors through the host; NCCL gathers them in place over NVLink)
    xdev = torch.device("cpu") if dist.get_backend(group) == "gloo" else head.device
    padded = torch.zeros((window, row_len), dtype=head.dtype, device=xdev)
    padded[:h] = head.to(xdev)
    heads = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(heads, padded, group=group)
    heads = [t.to(head.device) for t in heads]
    frames = [m["T"] for m in meta]
    plan = shard_window_plan(frames, rank, window, hop)
    halo = plan[2]
    parts, evs, ces, got, r = [], [], [], 0, rank + 1
    base = frames[rank]  # local row of the next rank's first frame
    while got < halo and r < world:
        take = min(halo - got, min(frames[r], window))
        if take > 0:
            parts.append(heads[r][:take])
            ev = meta[r]["ev"]
            if len(ev):
                sel = ev[:, 0] < take
                e = ev[sel].copy()
                e[:, 1] = np.minimum(e[:, 1], take)
                e[:, 0] += base
                e[:, 1] += base
                evs.append(e)
                if meta[r]["ce"] is not None:
                    ces.append(np.asarray(meta[r]["ce"])[sel])
        got += take
        base += frames[r]
        if frames[r] > window and got < halo:  # cannot happen: a halo is shorter than a window
            break
        r += 1
    halo_feat = torch.cat(parts) if parts else head.new_zeros((0, row_len))
    halo_events = np.concatenate(evs) if evs else np.zeros((0, 4), np.int32)
    halo_centres = np.concatenate(ces) if ces else None
    return plan, halo_feat, halo_events, halo_centres


class SELDDataset(Dataset):
    """Reference dataset.py:167-330 (+ smrl_seld_gaussian.py:539-700 for ``use_gaussian_augmentation``)."""

    def __init__(self, audio_files, metadata_files, num_classes=14, use_gaussian_augmentation=False, *,
                 device=None, resident="cpu", labels="dense", feature_type=None, audio_loader=None,
                 compute_stats=False, distributed=False, group=None):
        assert len(audio_files) == len(metadata_files), \
            "Number of audio files must match number of metadata files"
        config = get_config()
        self.audio_files = audio_files
        self.metadata_files = metadata_files
        self.sample_rate = config.SR
        self.n_fft = config.SPECTROGRAM_N_FFT
        self.spectrogram_hop_length = config.SPECTROGRAM_HOP_LENGTH
        self.n_mels = config.N_MELS
        self.cell_size_deg = config.GRID_CELL_DEGREES
        self.num_classes = num_classes
        self.use_gaussian_augmentation = use_gaussian_augmentation
        self.I = int(180 // self.cell_size_deg)
        self.J = int(360 // self.cell_size_deg)
        self.total_cells = self.I * self.J
        self.window_length_samples = config.WINDOW_LENGTH
        self.hop_length_samples = config.HOP_LENGTH
        self.window_length_frames = int(self.window_length_samples / self.spectrogram_hop_length)
        self.hop_length_frames = int(self.hop_length_samples / self.spectrogram_hop_length)

        if resident not in ("cpu", "cuda") or labels not in ("dense", "compact"):
            raise ValueError("resident must be 'cpu' or 'cuda'; labels must be 'dense' or 'compact'")
        if labels == "compact" and resident != "cuda":
            raise ValueError("labels='compact' paints label windows on the GPU: use resident='cuda'")
        self.resident, self.label_mode = resident, labels
        self.feature_type = feature_type or getattr(config, "FEATURE_TYPE", "logmel")
        self._mode = FEATURE_MODES[self.feature_type]
        self.device = L._cuda_device(device)
        # default ingest keeps 16-bit PCM files as int16 (converted inside the feature kernel); a custom loader may return either
        self._load_audio = audio_loader or load_audio_pcm16
        self._compute_stats = compute_stats
        # distributed=True: ``audio_files`` is this rank's contiguous block of the global list (``shard_files``); windows
        # are those of the GLOBAL concatenation that start in the block, bit-identical to the unsharded dataset's — the
        # frames they read past the block's end (at most window - 1 = 249) come from the following ranks in one neighbour exchange
        self._distributed, self._group = bool(distributed), group
        self.halo_frames, self.global_frame_offset, self.first_window = 0, 0, 0
        self._window_starts = None

        logger.info(f"SELDDataset initialization started...")
        logger.info(f"  Files: {len(audio_files)} audio files")
        logger.info(f"  Grid: {self.I}x{self.J} = {self.total_cells} cells")
        logger.info(f"  Window: {self.window_length_frames} frames ({self.window_length_samples / self.sample_rate:.1f}s)")
        logger.info(f"  Hop: {self.hop_length_frames} frames ({self.hop_length_samples / self.sample_rate:.1f}s)")
        logger.info(f"  Label augmentation: {'Gaussian noise' if use_gaussian_augmentation else 'Standard (metadata_to_labels)'}")

        self._load_and_concatenate_all()
        self._create_windows()
        logger.info(f"SELDDataset initialized with {len(self.windows)} windows")

    # ------------------------------------------------------------------------------------------
    def _load_and_concatenate_all(self):
        """dataset.py:212-265, restructured: pass 1 reads audio + metadata on the host (sizes, events), then the
        GPU writes every file's features and labels straight into the concatenated tensors."""
        logger.info("Loading and processing all audio files...")
        waves, per_file = [], []
        total = 0
        # files are read and decoded by a few threads (file I/O and numpy release the GIL) while this thread parses the
        # metadata in file order — the Gaussian label noise draws from numpy's global RNG in that order
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
        pending = [pool.submit(self._load_audio, a) for a in self.audio_files]
        pool.shutdown(wait=False)
        for idx, (audio_path, metadata_path) in enumerate(zip(self.audio_files, self.metadata_files)):
            try:
                waveform, sr = pending[idx].result()
                pending[idx] = None
                audio_duration = waveform.shape[1] / sr
                if self.use_gaussian_augmentation:
                    events, centres, t_lab = L.region_events(metadata_path, audio_duration, self.I, self.J,
                                                             self.num_classes)
                else:
                    events, t_lab = L.point_events(metadata_path, audio_duration, self.I, self.J, self.num_classes)
                    centres = None
                t_mel = 1 + waveform.shape[1] // self.spectrogram_hop_length
                keep = min(t_mel, t_lab)  # dataset.py:243-249
                if len(events):  # events beyond the kept frames are cropped with the labels
                    sel = events[:, 0] < keep
                    events = events[sel].copy()
                    events[:, 1] = np.minimum(events[:, 1], keep)
                    if centres is not None:
                        centres = centres[sel]
                    events[:, 0] += total
                    events[:, 1] += total
                waves.append((waveform, sr))
                per_file.append((total, keep, events, centres))
                total += keep
            except Exception as e:
                logger.error(f"Error processing file {idx} ({audio_path}): {str(e)}")
                raise

        dev = self.device
        n_ch = None
        plan = None
        self.total_frames = total
        self.file_offsets = [p[0] for p in per_file]
        self.file_frames = [p[1] for p in per_file]
        feats = None
        stats = None
        for i, (off, keep, _ev, _ce) in enumerate(per_file):
            waveform, sr = waves[i]
            waves[i] = None  # the host copy is released as soon as the file is on its way to the GPU
            if plan is None or plan.sample_rate != sr:
                plan = get_plan(self.n_fft, self.spectrogram_hop_length, self.n_mels, sr, dev)
            c_out = plan.out_channels(_lib_mode(self._mode), waveform.shape[0])
            if feats is None:
                n_ch = c_out
                feats = torch.empty((total, n_ch, self.n_mels), dtype=torch.float32, device=dev)
                if self._compute_stats:
                    stats = torch.zeros(2 * n_ch * self.n_mels, dtype=torch.float64, device=dev)
            elif c_out != n_ch:
                raise RuntimeError(f"Sizes of tensors must match: {c_out} feature channels vs {n_ch}")
            if keep == 0:
                continue
            x = waveform.to(device=dev, non_blocking=True)
            if x.dtype == torch.int16 and not (plan.fast and x.shape[0] == 4 and self._mode != "logmel_gcc"):
                x = x.to(torch.float32) * (1.0 / 32768.0)  # outside the fast kernel's configurations: exact, on the device
            elif x.dtype != torch.int16:
                x = x.to(torch.float32)
            plan.run(x.unsqueeze(0), mode=self._mode, out=feats[off:off + keep].unsqueeze(0), T_out=keep, stats=stats)
        if feats is None:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
        self.n_channels = n_ch
        self.stats = stats  # [sum | sum of squares] per (channel, mel) over all kept frames, float64, or None
        self.events = (np.concatenate([p[2] for p in per_file]) if per_file else np.zeros((0, 4), np.int32))
        ce = [p[3] for p in per_file if p[3] is not None]
        self.centres = np.concatenate(ce) if ce else None
        rows = total
        if self._distributed:  # windows of the global concatenation: fetch the frames they read from the next ranks
            W, H = self.window_length_frames, self.hop_length_frames
            h = min(total, W)
            ev = self.events
            sel = ev[:, 0] < h if len(ev) else np.zeros(0, bool)
            head_ev = ev[sel].copy()
            if len(head_ev):
                head_ev[:, 1] = np.minimum(head_ev[:, 1], h)
            head_ce = self.centres[sel] if self.centres is not None and len(ev) else None
            plan_, halo_feat, halo_ev, halo_ce = exchange_shard_halo(feats[:h].reshape(h, -1), head_ev, head_ce, total, W, H,
                                                                   self._group)
            self.global_frame_offset, self._window_starts, self.halo_frames, self.first_window = plan_
            if self.halo_frames:
                feats = torch.cat([feats, halo_feat.view(-1, n_ch, self.n_mels)], dim=0)
                self.events = np.concatenate([self.events, halo_ev.astype(np.int32)]) if len(halo_ev) else self.events
                if halo_ce is not None:
                    self.centres = halo_ce if self.centres is None else np.concatenate([self.centres, halo_ce])
            rows = total + self.halo_frames
        self._rows = rows  # frames available to windows: the rank's own + the halo of the following ranks
        if self.label_mode == "dense":
            lab = torch.empty((rows, self.total_cells, self.num_classes), dtype=torch.float32, device=dev)
            L.encode_dense(lab, self.events, self.centres, self.I, self.J)
        else:
            lab = None
        torch.cuda.synchronize(dev)
        if self.resident == "cpu":
            feats = feats.cpu()
            lab = lab.cpu() if lab is not None else None
        self._features_tcf = feats
        self.concatenated_labels = lab
        logger.info(f"Concatenated data: {self.total_frames} total frames")
        logger.info(f"  Spectrograms shape: {self.concatenated_spectrograms.shape}")
        if lab is not None:
            logger.info(f"  Labels shape: {self.concatenated_labels.shape}")

    @property
    def concatenated_spectrograms(self) -> torch.Tensor:
        """(C, n_mels, sum T) view, the reference's layout (dataset.py:259)."""
        return self._features_tcf.permute(1, 2, 0)

    # ------------------------------------------------------------------------------------------
    def _create_windows(self):
        """dataset.py:267-317: windows [50k, 50k+250) while 50k < sum T; the tail is padded with 0 features and
        background labels.  Full windows are views, exactly like the reference."""
        self.windows = []
        W, H, T = self.window_length_frames, self.hop_length_frames, self._rows
        starts = self._window_starts if self._distributed else range(0, self.total_frames, H)
        for window_idx, start_frame in enumerate(starts):
            end_frame = start_frame + W
            if end_frame <= T:
                window_spec = self._features_tcf[start_frame:end_frame]
                window_labels = self.concatenated_labels[start_frame:end_frame] if self.label_mode == "dense" else None
            else:
                pad_frames = W - (T - start_frame)
                f = self._features_tcf
                spec_pad = torch.zeros((pad_frames, f.shape[1], f.shape[2]), dtype=f.dtype, device=f.device)
                window_spec = torch.cat([f[start_frame:], spec_pad], dim=0)
                if self.label_mode == "dense":
                    l = self.concatenated_labels
                    label_pad = torch.zeros((pad_frames, self.total_cells, self.num_classes), dtype=l.dtype,
                                            device=l.device)
                    label_pad[:, :, self.num_classes - 1] = 1.0
                    window_labels = torch.cat([l[start_frame:], label_pad], dim=0)
                else:
                    window_labels = None
            self.windows.append({'spectrogram': window_spec, 'labels': window_labels, 'window_idx': window_idx,
                                 'start_frame': start_frame, 'end_frame': min(end_frame, T)})
        logger.info(f"Created {len(self.windows)} windows")

    def __len__(self):
        return len(self.windows)

    def __getitem__(self, idx):
        """(spectrogram (250, C, n_mels), labels (250, I*J, M)) — dataset.py:322-330."""
        window = self.windows[idx]
        if window['labels'] is None:  # compact mode: paint this single window
            lab = self.paint_label_windows([window['start_frame']])[0]
            return window['spectrogram'], lab
        return window['spectrogram'], window['labels']

    # ------------------------------------------------------------------------------------------
    def scaler(self, group=None, device=None):
        """Normalisation scaler of the whole (possibly rank-sharded) split: this dataset's per-feature partials
        (``compute_stats=True``) merged, summed over all ranks with ONE all-reduce (NCCL on GPUs) and finalised."""
        from .scaler import FeatureScaler
        if self.stats is None:
            raise ValueError("SELDDataset(compute_stats=True) is needed for scaler()")
        sc = FeatureScaler(self.n_channels * self.n_mels, device if device is not None else self.device)
        sc.merge(self.stats, self.total_frames)
        sc.sync(group)
        sc.finalize()
        return sc

    def _events_sorted(self):
        """Event table sorted by first row (stable), built once; painting order inside a window does not matter
        (the paint kernel's two passes are order-independent)."""
        if getattr(self, "_ev_sorted", None) is None:
            ev, ce = self.events, self.centres
            order = np.argsort(ev[:, 0], kind="stable") if len(ev) else np.zeros(0, np.int64)
            self._ev_sorted = (np.ascontiguousarray(ev[order]), None if ce is None else np.ascontiguousarray(ce[order]))
            self._ev_maxlen = int((ev[:, 1] - ev[:, 0]).max()) if len(ev) else 1
        return self._ev_sorted

    def paint_label_windows(self, starts, out: torch.Tensor | None = None) -> torch.Tensor:
        """Dense labels (n, W, I*J, M) for windows starting at ``starts``, painted on the GPU from the compact
        event tables (background for frames past the end, like the reference's padding)."""
        W, T = self.window_length_frames, self._rows
        n = len(starts)
        if out is None:
            out = torch.empty((n, W, self.total_cells, self.num_classes), dtype=torch.float32, device=self.device)
        ev_all, ce_all = self._events_sorted()
        evs, ces = [], []
        if len(ev_all):
            # events are sorted by first row and last at most _ev_maxlen rows: a window only needs a binary search
            row0 = ev_all[:, 0]
            s_arr = np.asarray(starts, dtype=np.int64)
            lo = np.searchsorted(row0, s_arr - (self._ev_maxlen - 1), side="left")
            hi = np.searchsorted(row0, np.minimum(s_arr + W, T), side="left")
            for w, s in enumerate(starts):
                e = ev_all[lo[w]:hi[w]]
                sel = e[:, 1] > s
                e = e[sel].copy()
                e[:, 0] = np.maximum(e[:, 0], s) - s + w * W
                e[:, 1] = np.minimum(e[:, 1], s + W) - s + w * W
                evs.append(e)
                if ce_all is not None:
                    ces.append(ce_all[lo[w]:hi[w]][sel])
        events = np.concatenate(evs) if evs else np.zeros((0, 4), np.int32)
        centres = np.concatenate(ces) if ces else None
        L.encode_dense(out.view(n * W, self.total_cells, self.num_classes), events, centres, self.I, self.J)
        return out


def shard_clips(n_clips: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of the sorted file list owned by ``rank``: clips are independent
    (reference dataset.py:218-252), so ranks never exchange audio or features; contiguous blocks keep the reference's
    concatenation order (dataset.py:125-128) inside each rank.  Sizes differ by at most one clip."""
    base, extra = divmod(int(n_clips), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_files(audio_files, metadata_files, rank=None, world_size=None):
    """This rank's contiguous block of the (audio, metadata) file lists — ``load_files()`` output sharded for one
    process per GPU.  ``rank`` / ``world_size`` default to ``torch.distributed`` (if initialised) or RANK / WORLD_SIZE."""
    import os

    import torch.distributed as dist
    if rank is None or world_size is None:
        if dist.is_available() and dist.is_initialized():
            rank, world_size = dist.get_rank(), dist.get_world_size()
        else:
            rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert len(audio_files) == len(metadata_files), "Number of audio files must match number of metadata files"
    lo, hi = shard_clips(len(audio_files), rank, world_size)
    return list(audio_files[lo:hi]), list(metadata_files[lo:hi])


def _lib_mode(mode: str) -> int:
    from .features import MODES
    return MODES[mode]


class DeviceLoader:
    """On-device replacement for ``DataLoader(SELDDataset)`` (main.py:60-74 builds one with 2 workers and pinned
    memory; trainer.py only needs ``.dataset``, ``len()`` and iteration — trainer.py:42-47, :165-168).

    Everything a batch needs is resident in HBM: the concatenated features, the event table sorted by first row, the
    start frame and the event range of every window, and the epoch's permutation (uploaded once per epoch).  A batch
    ``(B, 250, C, 64)`` / ``(B, 250, I*J, M)`` is ONE kernel launch (``seld_loader_batch``: feature gather + background
    fill + event painting) into a preallocated ring of ``depth`` output buffers — no per-batch host work besides that
    call, no host->device copy, no synchronisation.  The tensors of a batch are overwritten ``depth`` batches later
    (the training loop consumes a batch before asking for the next, trainer.py:165-179).
    ``labels='dense'`` datasets (labels materialised in HBM) gather the label windows instead of painting them."""

    def __init__(self, dataset: SELDDataset, batch_size=16, shuffle=False, drop_last=False, generator=None, depth=3,
                 targets="dense"):
        if dataset.resident != "cuda":
            raise ValueError("DeviceLoader needs SELDDataset(resident='cuda')")
        if targets not in ("dense", "mask"):
            raise ValueError("targets must be 'dense' ((B, 250, I*J, M) float32) or 'mask' ((B, 250, I*J) int16 class sets)")
        if targets == "mask" and dataset.label_mode != "compact":
            raise ValueError("targets='mask' needs SELDDataset(labels='compact')")
        self.targets = targets
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.generator = generator
        ds, dev = dataset, dataset.device
        W = ds.window_length_frames
        starts = np.asarray([w['start_frame'] for w in ds.windows], dtype=np.int64)
        self._win_start = torch.from_numpy(starts.astype(np.int32)).to(dev)
        self._win_start64 = torch.from_numpy(starts).to(dev)
        self._identity = torch.arange(len(starts), dtype=torch.int32, device=dev)
        row = ds.n_channels * ds.n_mels
        self._ev = self._ce = self._lo = self._hi = None
        if ds.label_mode == "compact":
            ev, ce = ds._events_sorted()
            if len(ev):
                row0 = ev[:, 0]
                lo = np.searchsorted(row0, starts - (ds._ev_maxlen - 1), side="left")
                hi = np.searchsorted(row0, np.minimum(starts + W, ds._rows), side="left")
            else:
                lo = hi = np.zeros(len(starts), np.int64)
            self._lo = torch.from_numpy(lo.astype(np.int32)).to(dev)
            self._hi = torch.from_numpy(hi.astype(np.int32)).to(dev)
            self._ev = torch.from_numpy(np.ascontiguousarray(ev, dtype=np.int32)).to(dev) if len(ev) else None
            self._ce = torch.from_numpy(np.ascontiguousarray(ce, dtype=np.float64)).to(dev) if ce is not None and len(ce) else None
        else:
            pad_lab = torch.zeros((ds.total_cells, ds.num_classes), dtype=torch.float32, device=dev)
            pad_lab[:, ds.num_classes - 1] = 1.0
            self._pad_lab = pad_lab.reshape(-1)
        B = self.batch_size
        # targets='mask' (SURVEY.md §8(f) N4): int16 class sets for seld_b200.loss.CompactSMRSELDLoss instead of dense labels
        lab_shape, lab_dtype = (((B, W, ds.total_cells), torch.int16) if targets == "mask"
                                else ((B, W, ds.total_cells, ds.num_classes), torch.float32))
        self._ring = [(torch.empty((B, W, ds.n_channels, ds.n_mels), dtype=torch.float32, device=dev),
                       torch.empty(lab_shape, dtype=lab_dtype, device=dev))
                      for _ in range(max(2, int(depth)))]
        self._slot = 0
        self._args = (ds._features_tcf.data_ptr(), ds._features_tcf.shape[0], row)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        ds = self.dataset
        n, B, W = len(ds), self.batch_size, ds.window_length_frames
        if self.shuffle:  # the permutation is drawn like DataLoader's RandomSampler and uploaded once per epoch
            order = torch.randperm(n, generator=self.generator).to(torch.int32).to(ds.device, non_blocking=True)
        else:
            order = self._identity
        lib, check = _lib.lib(), _lib.check
        feat_ptr, rows, row = self._args
        compact = ds.label_mode == "compact"
        ev_ptr, ce_ptr = _lib.ptr(self._ev), _lib.ptr(self._ce)
        lo_ptr, hi_ptr = _lib.ptr(self._lo), _lib.ptr(self._hi)
        order_ptr, start_ptr = order.data_ptr(), self._win_start.data_ptr()
        stream = torch.cuda.current_stream(ds.device).cuda_stream
        for first in range(0, n, B):
            nb = min(B, n - first)
            if self.drop_last and nb < B:
                break
            spec, lab = self._ring[self._slot]
            self._slot = (self._slot + 1) % len(self._ring)
            masks = self.targets == "mask"
            check(lib.seld_loader_batch(feat_ptr, rows, row, order_ptr, first, nb, start_ptr, lo_ptr, hi_ptr, W, spec.data_ptr(),
                                        ev_ptr, ce_ptr, ds.I, ds.J, ds.num_classes, 5.0, 5.0,
                                        lab.data_ptr() if compact and not masks else None, stream), "seld_loader_batch")
            if masks:
                check(lib.seld_batch_class_mask(order_ptr, first, nb, start_ptr, lo_ptr, hi_ptr, W, ev_ptr, ce_ptr, ds.I, ds.J,
                                                ds.num_classes, 5.0, 5.0, lab.data_ptr(), stream), "seld_batch_class_mask")
            if not compact:  # dense labels resident in HBM: gather their windows
                l = ds.concatenated_labels
                starts_dev = self._win_start64[order[first:first + nb].long()]
                check(lib.seld_window_gather(l.data_ptr(), l.shape[0], l.shape[1] * l.shape[2], starts_dev.data_ptr(), nb, W,
                                             self._pad_lab.data_ptr(), lab.data_ptr(), stream), "seld_window_gather")
            yield (spec, lab) if nb == B else (spec[:nb], lab[:nb])
