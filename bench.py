#!/usr/bin/env python
"""Benchmark of the SELD feature front-end hot path (BASELINE.json configs[1]).

Step = one pass of the feature front-end over a batch of 256 synthetic 60 s 4-channel 24 kHz clips per GPU
(n_fft 1024, hop 480, 64 mel -> 7-channel log-mel + IV features), inputs resident in HBM: the fused feature
kernel, the scaler-partials kernel and (N > 1) the all-reduce of those partials.
Metric: audio clip-seconds per second, whole job (all ranks).  One JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the reference's CPU implementation of the same path on the host cores (the torch /
torchaudio call sequence of reference dataset.py:38-56, ported in oracle/ref_port.py because the reference
tree cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, N_FFT, HOP, N_MELS, CH = 24000, 1024, 480, 64, 4
CLIP_SECONDS = 60
BYTES_PER_CLIP_SECOND = 4 * SR * CH + 7 * N_MELS * (SR // HOP) * 4  # 473 600
METRIC, UNIT = "audio_clip_seconds_per_sec", "clip-s/s"
WORKLOAD = ("configs[1]: batch of 256 synthetic 60 s 4-ch FOA clips @24 kHz per GPU -> 7-ch log-mel+IV, "
            "n_fft 1024 / hop 480 / 64 mel")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _cpu_clip(seed=1234):
    import torch
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(CH, SR * CLIP_SECONDS, generator=g)


def _cpu_time_clips(threads: int, seconds_budget: float, min_clips: int):
    """clip-s/s of oracle.ref_port.logmel_iv_port (the reference's torchaudio call sequence, dataset.py:38-56, plus the
    same IV arithmetic) on one 60 s 4-ch clip, torch intra-op threads = ``threads``."""
    import torch

    from oracle import ref_port

    torch.set_num_threads(threads)
    x = _cpu_clip()
    fb = ref_port._fb(N_FFT, SR, N_MELS)
    ref_port.logmel_iv_port(x, SR, N_FFT, HOP, N_MELS, fb)  # warm-up (MKL plan, thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        ref_port.logmel_iv_port(x, SR, N_FFT, HOP, N_MELS, fb)
        n += 1
        el = time.perf_counter() - t0
        if n >= min_clips and el >= seconds_budget:
            break
    return n * CLIP_SECONDS / el, n


def _pool_worker(args):
    seconds_budget, seed = args
    import torch
    torch.set_num_threads(1)
    from oracle import ref_port
    x = _cpu_clip(seed)
    fb = ref_port._fb(N_FFT, SR, N_MELS)
    ref_port.logmel_iv_port(x, SR, N_FFT, HOP, N_MELS, fb)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds_budget:
        ref_port.logmel_iv_port(x, SR, N_FFT, HOP, N_MELS, fb)
        n += 1
    return n, time.perf_counter() - t0


def _cpu_pool_rate(workers: int, seconds_budget: float):
    """Process pool of single-thread workers over independent clips (BASELINE.md §4 leg iii)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        res = pool.map(_pool_worker, [(seconds_budget, 1234 + i) for i in range(workers)])
    return sum(n for n, _ in res) * CLIP_SECONDS / max(t for _, t in res), sum(n for n, _ in res)


LABEL_SAMPLE_SECONDS = 20  # bounded sample: the reference needs ~22 s of CPU per 60 s clip


def _cpu_label_rate():
    """The reference's label encoder on the CPU with its own cost structure (oracle/ref_port.metadata_to_labels_port:
    per-element torch writes from Python loops, dataset.py:60-119) on a bounded sample: one 20 s clip with 3 sources ->
    dense (1000, 648, 14) targets.  clip-s/s."""
    import tempfile

    import numpy as np

    from oracle import ref_port
    rng = np.random.default_rng(0)
    rows = []
    for src in range(3):
        az, el, cls = rng.integers(-180, 180), rng.integers(-60, 60), rng.integers(0, 13)
        for f in range(10 * LABEL_SAMPLE_SECONDS):
            rows.append((f, cls, src, ((az + f // 10 + 180) % 360) - 180, el))
    rows.sort()
    path = os.path.join(tempfile.mkdtemp(prefix="seld_cpu_"), "clip.csv")
    with open(path, "w") as fh:
        fh.writelines(",".join(str(int(v)) for v in r) + "\n" for r in rows)
    t0 = time.perf_counter()
    lab = ref_port.metadata_to_labels_port(path, float(LABEL_SAMPLE_SECONDS), I=18, J=36)
    el = time.perf_counter() - t0
    assert tuple(lab.shape) == (50 * LABEL_SAMPLE_SECONDS, 648, 14)
    return LABEL_SAMPLE_SECONDS / el, el


def cpu_baseline_legs(seconds_budget: float, labels: bool = True):
    """(i) all torch threads, (ii) one thread, (iii) a process pool of single-thread workers, best of the
    three reported as the reference CPU figure; plus the label encoder leg."""
    cores = os.cpu_count() or 1
    v_all, n_all = _cpu_time_clips(cores, seconds_budget, 4)
    v_one, n_one = _cpu_time_clips(1, min(seconds_budget, 5.0), 2)
    try:
        v_pool, n_pool = _cpu_pool_rate(cores, min(seconds_budget, 8.0))
    except Exception as e:  # noqa: BLE001 - a box that cannot spawn workers still gets the other legs
        v_pool, n_pool = None, str(e)
    legs = {"all_threads": {"value": v_all, "threads": cores, "clips": n_all},
            "one_thread": {"value": v_one, "threads": 1, "clips": n_one},
            "process_pool": {"value": v_pool, "workers": cores, "clips": n_pool}}
    best = max((k for k in legs if legs[k]["value"]), key=lambda k: legs[k]["value"])
    out = {"value": legs[best]["value"], "unit": UNIT, "cores": cores, "kind": "port", "best_leg": best, "legs": legs,
           "sample": f"60 s 4-ch clips -> 7-ch log-mel+IV with oracle/ref_port.py (torch.stft on MKL; the reference's call "
                     f"sequence dataset.py:38-56 as ONE batched stft, which favours the CPU), {cores} host threads; "
                     f"clips per leg: {n_all} / {n_one} / {n_pool}"}
    if labels:
        v_lab, sec = _cpu_label_rate()
        out["labels"] = {"value": v_lab, "unit": UNIT, "seconds": sec, "cores": 1,
                         "sample": f"oracle/ref_port.metadata_to_labels_port (the reference's per-element Python loops, "
                                   f"dataset.py:60-119; single-threaded by construction) on one {LABEL_SAMPLE_SECONDS} s CSV, 3 sources -> "
                                   f"({50 * LABEL_SAMPLE_SECONDS}, 648, 14)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import ref_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clips_per_step = args.ref_clips
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(CH, SR * CLIP_SECONDS, generator=g)
    fb = ref_port._fb(N_FFT, SR, N_MELS)

    def step():
        for _ in range(clips_per_step):
            ref_port.logmel_iv_port(x, SR, N_FFT, HOP, N_MELS, fb)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = args.steps * clips_per_step * CLIP_SECONDS / el
    sample = (f"{clips_per_step} x 60 s clips per step (bounded sample of the 256-clip batch; ONE cache-resident clip looped and one "
              f"batched torch.stft instead of the reference's per-channel loop — both favour the CPU), torch {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_aux(args):
    """Secondary rows of the hot-path table: same JSON contract, N = 1, device-resident inputs.
    mic    : configs[2], 4 log-mel + 6 GCC-PHAT x 64 lags (512 000 algorithmic bytes per clip-second)
    logmel : the reference parity path, 4 log-mel channels (435 200 B per clip-second)
    labels : dense (T, 648, 14) float32 grid targets, fill + paint (1 814 400 B per clip-second, write-only)"""
    import numpy as np
    import torch

    import seld_b200

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    N = SR * CLIP_SECONDS
    T = 1 + N // HOP
    peak, peak_src = peaks()
    if args.workload == "loader":
        return run_loader(args, dev, peak, peak_src)
    if args.workload in ("mic", "logmel"):
        B = args.clips
        mode, n_out, bytes_cs = (("logmel_gcc", 10, 4 * SR * CH + 10 * N_MELS * (SR // HOP) * 4) if args.workload == "mic"
                                 else ("logmel", 4, 4 * SR * CH + 4 * N_MELS * (SR // HOP) * 4))
        gen = torch.Generator(device=dev).manual_seed(1234)
        audio = torch.empty((B, CH, N), dtype=torch.float32, device=dev).normal_(0.0, 0.1, generator=gen)
        out = torch.empty((B, T, n_out, N_MELS), dtype=torch.float32, device=dev)
        plan = seld_b200.get_plan(N_FFT, HOP, N_MELS, SR, dev)
        step = lambda: plan.run(audio, mode=mode, out=out)
        units, launches = B * CLIP_SECONDS, (2 if args.workload == "mic" else 1)
        wl = (f"configs[2]: {B} synthetic 60 s 4-mic clips -> 4 log-mel + 6 GCC-PHAT x 64 lags" if args.workload == "mic"
              else f"{B} synthetic 60 s 4-ch clips -> 4 log-mel channels (reference dataset.py:27-58 parity path)")
        kernel = "seld::features_v3_kernel<32, false> + seld::gcc_phat_kernel" if args.workload == "mic" else "seld::features_v3_kernel<32, false>"
    else:
        from seld_b200 import labels as L
        B = min(args.clips, 64)
        frames = 3000
        rng = np.random.default_rng(0)
        ev = []
        for b in range(B):  # 4 sources per clip, one CSV row per 100 ms each -> 5 label frames per row
            for src in range(4):
                f100 = np.arange(600)
                cell = (rng.integers(0, 648) + f100 // 10) % 648
                cls = np.full(600, rng.integers(0, 13))
                r0 = b * frames + 5 * f100
                ev.append(np.stack([r0, np.minimum(r0 + 5, (b + 1) * frames), cls, cell], 1))
        ev = np.concatenate(ev).astype(np.int32)
        out = torch.empty((B * frames, 648, 14), dtype=torch.float32, device=dev)
        ev_d = torch.from_numpy(ev).to(dev)
        lib, chk = seld_b200._lib.lib(), seld_b200._lib.check
        st = torch.cuda.current_stream(dev).cuda_stream

        def step():
            chk(lib.seld_labels_fill(out.data_ptr(), out.shape[0], 648, 14, st), "fill")
            chk(lib.seld_labels_paint(out.data_ptr(), out.shape[0], 18, 36, 14, ev_d.data_ptr(), None, len(ev), 5.0, 5.0, st), "paint")
        units, launches, bytes_cs = B * CLIP_SECONDS, 3, 50 * 648 * 14 * 4
        wl = f"{B} x 60 s clips -> dense (T, 648, 14) float32 grid labels (reference dataset.py:60-119), {len(ev)} events"
        kernel = "seld::labels_fill_vec4 + seld::labels_paint_kernel (2 passes)"
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    achieved = bytes_cs * units / (ms / 1e3) / 1e9
    print(json.dumps({
        "metric": METRIC, "value": units / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "l2": "working set >> 126 MB L2", "timing": "CUDA events on the launch stream"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": kernel,
                     "algorithmic_bytes_per_launch": bytes_cs * units},
        "cpu_baseline": None, "e2e": None, "gpu_launches": launches * args.steps, "clocks": clocks}))


def _synthetic_dataset(dev, n_files=16, seed=0):
    """n_files x 60 s synthetic FOA clips + STARSS-style CSVs (3 moving sources each) -> SELDDataset resident in HBM with
    compact labels (Gaussian-region targets painted per batch) — BASELINE configs[4]'s front-end."""
    import tempfile

    import numpy as np
    import torch

    import seld_b200

    N = SR * CLIP_SECONDS
    tmp = tempfile.mkdtemp(prefix="seld_bench_")
    rng = np.random.default_rng(seed)
    csvs = []
    for i in range(n_files):  # rows: frame(100 ms), class, source, azimuth, elevation
        rows = []
        for src in range(3):
            az, el, cls = rng.integers(-180, 180), rng.integers(-60, 60), rng.integers(0, 13)
            for f in range(int(rng.integers(0, 100)), 600):
                rows.append((f, cls, src, ((az + f // 10 + 180) % 360) - 180, el))
        rows.sort()
        path = os.path.join(tmp, f"f{i}.csv")
        with open(path, "w") as fh:
            fh.writelines(",".join(str(int(v)) for v in r) + "\n" for r in rows)
        csvs.append(path)
    gen = torch.Generator().manual_seed(1234 + seed)

    def loader(path):
        return 0.1 * torch.randn(CH, N, generator=gen), SR

    np.random.seed(42 + seed)
    return seld_b200.SELDDataset([f"synthetic://{i}" for i in range(n_files)], csvs, use_gaussian_augmentation=True,
                                 resident="cuda", labels="compact", feature_type="foa_iv", audio_loader=loader, device=dev)


def run_train_step(args):
    """BASELINE configs[4]: the on-device front-end (features resident in HBM, batch = 16 windows x 250 frames gathered and
    Gaussian-region targets painted per batch, one launch) feeding the reference's own SELD_Conformer(n_channels=7)
    training step (trainer.py:165-179: forward, SMRSELDLoss, backward, Adam step, loss.item()), one process per GPU
    (DistributedDataParallel over NCCL when N > 1).  The model and loss are the reference's files, staged — never
    committed — under baseline/_ref/ by __graft_entry__.build().  Reports front-end ms, model ms and the fraction."""
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def emit(line):
        if rank == 0:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "model_conformer.py")):
        return emit({"metric": "train_steps_per_sec", "workload": "train_step",
                     "unavailable": "baseline/_ref/model_conformer.py is not staged (run __graft_entry__.build() where /root/reference exists)"})
    sys.path.insert(0, ref_dir)
    from loss import SMRSELDLoss  # reference loss.py:6
    from model_conformer import SELD_Conformer  # reference model_conformer.py:116

    import seld_b200

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ds = _synthetic_dataset(dev, n_files=16, seed=rank)  # every rank owns its shard of clips
    dl = seld_b200.DeviceLoader(ds, batch_size=16, shuffle=True, drop_last=True, generator=torch.Generator().manual_seed(rank),
                                targets="mask" if args.compact_loss else "dense")
    torch.manual_seed(0)
    model = SELD_Conformer(n_channels=7, n_mels=N_MELS, grid_size=(ds.I, ds.J), num_classes=14).to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    if args.compact_loss:  # N4: int16 class-set masks + seld_class_loss instead of dense targets + the reference loss
        criterion = seld_b200.CompactSMRSELDLoss(loss_type="mse", grid_size=(ds.I, ds.J))
    else:
        criterion = SMRSELDLoss(loss_type="mse", grid_size=(ds.I, ds.J))  # config.py:71 LOSS_TYPE = 'mse'
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()

    def train_step(spec, lab):  # trainer.py:170-179
        optimizer.zero_grad()
        pred = model(spec)
        loss, _ = criterion(pred, lab)
        loss.backward()
        optimizer.step()
        return loss

    it = iter(dl)
    def next_batch():
        nonlocal it
        try:
            return next(it)
        except StopIteration:
            it = iter(dl)
            return next(it)

    for _ in range(max(args.warmup, 3)):
        train_step(*next_batch()).item()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        spec, lab = next_batch()       # front-end: one launch (gather + fill + paint) into the ring buffer
        ev[i][1].record()
        loss = train_step(spec, lab)   # the reference's model step
        ev[i][2].record()
        lv = loss.item()               # trainer.py:182 reads the loss every step
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    clocks = sampler.stop()
    fe = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    md = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    if world > 1:
        t = torch.tensor([el, fe, md], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el, fe, md = (float(v) for v in t.tolist())
    line = {
        "metric": "train_steps_per_sec", "value": world * steps / el, "unit": "steps/s (global batch = 16 windows x n_gpus)",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: on-device front-end (features + Gaussian label targets per batch of 16 windows x "
                               "250 frames) feeding the reference SELD_Conformer(n_channels=7) train step, 16 x 60 s clips per GPU",
                   "model": ("reference model_conformer.SELD_Conformer + " + ("seld_b200.CompactSMRSELDLoss('mse') on int16 class-set masks (N4)"
                             if args.compact_loss else "loss.SMRSELDLoss('mse')") + " + Adam (staged in baseline/_ref, fp32)"),
                   "parallelism": f"ddp{world}" if world > 1 else "single GPU", "timing": "wall clock; per-part CUDA events"},
        "frontend_ms_per_step": fe, "model_ms_per_step": md, "frontend_fraction_of_step": fe / (fe + md),
        "reference_frontend_bytes_h2d_per_step": 16 * 250 * (7 * 64 + 648 * 14) * 4, "last_loss": lv,
        "gpu_launches": steps, "clocks": clocks,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_loader(args, dev, peak, peak_src):
    """configs[4] front-end part: per training batch of 16 windows x 250 frames, features (16, 250, 7, 64) gathered
    and Gaussian-region label targets (16, 250, 648, 14) painted on the device from compact event tables
    (SELDDataset(resident='cuda', labels='compact') + DeviceLoader); metric = batches per second of the front-end
    alone, reported in clip-seconds (16 windows x 5 s per batch)."""
    import torch

    import seld_b200

    n_files = 16
    ds = _synthetic_dataset(dev, n_files=n_files, seed=0)
    dl = seld_b200.DeviceLoader(ds, batch_size=16, shuffle=True, generator=torch.Generator().manual_seed(0))
    for _ in dl:  # warm-up epoch
        pass
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, nb, chk = time.perf_counter(), 0, 0.0
    e0.record()
    for _ in range(max(1, args.steps // 2)):
        for spec, lab in dl:
            nb += 1
    e1.record()
    chk = float(lab[0, 0, 0, 13]) + float(spec[0, 0, 0, 0])  # device -> host read of the last batch
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    bytes_batch = 16 * 250 * (7 * 64 + 648 * 14) * 4
    achieved = bytes_batch * nb / el / 1e9
    print(json.dumps({
        "metric": "frontend_batches_per_sec", "value": nb / el, "unit": "batches/s", "n_gpus": 1, "steps": nb,
        "warmup": len(dl), "ms_per_step": 1e3 * el / nb, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[4] front-end: batch = 16 windows x 250 frames from {n_files} x 60 s clips "
                               f"({len(ds)} windows): features gather + Gaussian-region label painting on device",
                   "timing": "wall clock around whole epochs incl. the host side of every batch (one C-ABI call)",
                   "device_ms_per_batch": dev_ms / nb, "check": chk},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "seld::loader_batch_kernel (gather + fill + paint, one launch per batch)",
                     "algorithmic_bytes_per_launch": bytes_batch},
        "cpu_baseline": None, "e2e": None, "gpu_launches": nb, "clocks": clocks}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-clips", type=int, default=8, help="clips per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--compact-loss", action="store_true", help="--workload train_step: class-set masks + seld_class_loss (N4)")
    ap.add_argument("--corpus-clips", type=int, default=600, help="--workload corpus: clips of the whole job (~10 h at 600)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="foa", choices=["foa", "mic", "logmel", "labels", "loader", "train_step", "corpus"],
                    help="foa = BASELINE configs[1] (default, the contract line); the others are secondary rows")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train_step":
        return run_train_step(args)
    if args.workload != "foa" and args.workload != "corpus":
        return run_aux(args)

    # stdout carries exactly ONE JSON line: libraries that write to fd 1 from C (NCCL prints its version there)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import seld_b200
    from seld_b200.dataset import shard_clips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    corpus = args.workload == "corpus"
    N = SR * CLIP_SECONDS
    T = 1 + N // HOP
    if corpus:  # BASELINE configs[3]: ~10 h (600 x 60 s clips), contiguous clip blocks per rank, strong scaling
        lo, hi = shard_clips(args.corpus_clips, rank, world)
        B = hi - lo
    else:       # BASELINE configs[1]: 256 clips per GPU, weak scaling
        B = args.clips
    chunk_clips = min(B, 256)

    # synthetic shard of this rank, generated on the device (seed 1234 + rank), resident in HBM
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    audio = torch.empty((B, CH, N), dtype=torch.float32, device=dev)
    audio.normal_(0.0, 0.1, generator=gen)
    out = torch.empty((B, T, 7, N_MELS), dtype=torch.float32, device=dev)
    stats = torch.zeros(2 * 7 * N_MELS + 1, dtype=torch.float64, device=dev)  # [sum | sum of squares | frame count]
    stat_frames = torch.full((B,), T - 1, dtype=torch.int32, device=dev)  # the frames SELDDataset keeps
    plan = seld_b200.get_plan(N_FFT, HOP, N_MELS, SR, dev)
    chunks = [(b0, min(B, b0 + chunk_clips)) for b0 in range(0, B, chunk_clips)]

    def step(kev=None):
        """One pass of the hot path over this rank's clips: per chunk of <= 256 clips the fused feature kernels (lean +
        block-floating redo list; bracketed by their own events) and the scaler-partials kernel."""
        for ci, (b0, b1) in enumerate(chunks):
            if kev is not None:
                kev[ci][0].record()
            plan.run(audio[b0:b1], mode="logmel_iv", out=out[b0:b1])
            if kev is not None:
                kev[ci][1].record()
            plan.accumulate_stats(out[b0:b1], stats[:-1], stat_frames=stat_frames[b0:b1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        stats.zero_()
        step()
    if world > 1:
        dist.all_reduce(stats)  # warm the communicator
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    kev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in chunks]
           for _ in range(args.steps)]
    stats.zero_()
    stats[-1] = float(args.steps) * B * (T - 1)  # frame count of the pass (known up front; no kernel of ours needed)
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(kev[i])
    ev[1].record()
    if world > 1:  # the path's only collective, ONCE per pass over the data: ~7 KB of fp64 partials (SELDDataset.scaler())
        dist.all_reduce(stats)
    ev[2].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[2])
    allreduce_ms = ev[1].elapsed_time(ev[2]) if world > 1 else 0.0
    kern_ms = sum(a.elapsed_time(b) for st in kev for a, b in st) / args.steps  # feature kernels alone, same timed region
    if world > 1:
        t = torch.tensor([total_ms, kern_ms, allreduce_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kern_ms, allreduce_ms = (float(v) for v in t.tolist())
    clip_s_per_step = B * CLIP_SECONDS
    job_clip_s = (args.corpus_clips if corpus else world * B) * CLIP_SECONDS
    value = job_clip_s * args.steps / (total_ms / 1e3)

    # dominant kernel: the fused feature kernel, timed by its own CUDA events inside the timed region
    peak, peak_src = peaks()
    achieved = BYTES_PER_CLIP_SECOND * clip_s_per_step / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "seld::features_fast_kernel<32, IV, f32, lean> (+ the block-floating launch over its redo list)",
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": BYTES_PER_CLIP_SECOND * min(B, chunk_clips) * CLIP_SECONDS}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tfile):
        try:
            roofline["traffic"] = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the public API with HOST buffers (pinned): H2D + kernels + D2H every step ----
    e2e = None
    if not args.no_e2e and not corpus:
        from seld_b200.features import extract_features_host
        Be = min(B, 256)
        h_out = torch.empty((Be, T, 7, N_MELS), dtype=torch.float32, pin_memory=True)
        chunk = 16

        def timed(fn, n):
            fn()  # warm-up (staging buffers)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            barrier()
            el_ = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([el_], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                el_ = float(tt.item())
            return el_ / n

        # contract figure: 16-bit PCM host buffers — what the WAV files of the reference's dataset hold (load_audio,
        # dataset.py:18-25, returns x / 32768 of them); the kernel converts while loading
        h_pcm = torch.empty((Be, CH, N), dtype=torch.int16, pin_memory=True)
        for b0 in range(0, Be, chunk):
            h_pcm[b0:b0 + chunk].copy_((audio[b0:b0 + chunk] * 32768.0).clamp_(-32768, 32767).to(torch.int16))
        torch.cuda.synchronize()
        sec = timed(lambda: extract_features_host(h_pcm, h_out, plan, mode="logmel_iv"), args.e2e_steps)
        h2d, d2h = h_pcm.numel() * 2, h_out.numel() * 4
        # the same pipeline without the kernels: what the host side / PCIe alone allows (names the limiter)
        d_stage = [torch.empty((chunk, CH, N), dtype=torch.int16, device=dev) for _ in range(3)]
        d_ostage = [torch.empty((chunk, T, 7, N_MELS), dtype=torch.float32, device=dev) for _ in range(3)]
        streams = [torch.cuda.Stream(dev) for _ in range(3)]

        def copies_only():
            for i, b0 in enumerate(range(0, Be, chunk)):
                with torch.cuda.stream(streams[i % 3]):
                    d_stage[i % 3].copy_(h_pcm[b0:b0 + chunk], non_blocking=True)
                    h_out[b0:b0 + chunk].copy_(d_ostage[i % 3], non_blocking=True)
            for st in streams:
                st.synchronize()
        sec_copy = timed(copies_only, args.e2e_steps)
        e2e = {"value": world * Be * CLIP_SECONDS / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": args.e2e_steps, "ms_per_step": 1e3 * sec, "host_dtype": "int16 PCM in (as stored in the dataset's WAV files), float32 features out",
               "api": "seld_b200.features.extract_features_host (pinned host in/out, 3-stream chunked pipeline, int16 converted inside the feature kernel)",
               "per_gpu_h2d_gbs": h2d / sec / 1e9, "per_gpu_d2h_gbs": d2h / sec / 1e9,
               "copies_only_ms_per_step": 1e3 * sec_copy,
               "limiter": ("host<->device copies (PCIe / host memory): the same H2D + D2H traffic without any kernel takes "
                           f"{100 * sec_copy / sec:.0f} % of the end-to-end time"),
               "check": float(h_out[0, 0, 0, 0])}
        del h_pcm, d_stage
        # same path fed with float32 host buffers (4 bytes per sample over PCIe)
        h_audio = torch.empty((Be, CH, N), dtype=torch.float32, pin_memory=True)
        for b0 in range(0, Be, chunk):
            h_audio[b0:b0 + chunk].copy_(audio[b0:b0 + chunk])
        torch.cuda.synchronize()
        sec2 = timed(lambda: extract_features_host(h_audio, h_out, plan, mode="logmel_iv"), args.e2e_steps)
        e2e["f32_host"] = {"value": world * Be * CLIP_SECONDS / sec2, "unit": UNIT, "h2d_bytes_per_step": h_audio.numel() * 4,
                           "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * sec2,
                           "note": "host input already decoded to float32 (round 1's contract figure)"}
        del h_out, h_audio

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_legs(args.cpu_seconds)

    if rank == 0:
        wl = (f"configs[3]: synthetic corpus of {args.corpus_clips} x 60 s 4-ch FOA clips (~{args.corpus_clips / 60:.0f} h) sharded by clip "
              f"over {world} GPU(s) -> 7-ch log-mel+IV + scaler partials, ONE all-reduce per pass" if corpus else WORKLOAD)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if corpus else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "clips_per_gpu": B, "clip_seconds": CLIP_SECONDS,
                       "l2": "inputs >= 5.9 GB per step >> 126 MB L2 (no flush needed)",
                       "collective": ("all_reduce(fp64 scaler partials, 7 KB) once per pass, inside the timed region"
                                      if world > 1 else "none (1 GPU)"),
                       "allreduce_ms": allreduce_ms,
                       "timing": "CUDA events on the launch stream, max over ranks"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": 3 * len(chunks) * args.steps, "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
