"""Deterministic synthetic inputs shared by make_golden.py (run once, in the build container, against the
real reference) and by the parity tests (run anywhere).  Inputs are regenerated from seeds; each golden
file stores the sha256 of the input bytes so a numpy RNG change cannot go unnoticed."""
from __future__ import annotations

import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSV_DIR = os.path.join(HERE, "csv")
SR = 24000
HOP = 480
N_MELS = 64

# name -> (kind, n_samples, seed)
AUDIO_CASES = {
    "noise_1s": ("noise", 24000, 1234),
    "noise_n1000": ("noise", 1000, 7),
    "noise_n24001": ("noise", 24001, 8),
    "noise_n24479": ("noise", 24479, 9),
    "noise_n24480": ("noise", 24480, 10),
    "noise_n97440": ("noise", 97440, 11),
    "zeros": ("zeros", 4800, 0),
    "impulse_first": ("impulse_first", 4800, 0),
    "impulse_last": ("impulse_last", 4800, 0),
    "sine_1k_1e-4_ch0": ("sine", 12000, 0),
    "int16_noise": ("int16", 12000, 21),
    "loud_noise": ("loud", 9600, 22),
    # unequal channel levels (round 2): the reference transforms every channel on its own (dataset.py:46-50), a
    # kernel that packs two channels per complex FFT must not let the loud one's rounding noise into the quiet one
    "level_60db": ("level60", 12000, 23),     # channels 0 and 3 are 60 dB below their FFT partners 1 and 2
    "level_80db": ("level80", 12000, 24),     # channels 1 and 2 are 80 dB below their partners 0 and 3
    "level_100db": ("level100", 9600, 25),    # channel 0: -100 dB, channel 2: -40 dB
    "level_ramp": ("ramp", 24000, 26),        # channel 1 fades from 0 dB to -100 dB over the clip, channel 3 fades in
}
# BASELINE.json configs[0]: one 60 s clip; the golden keeps every CONFIG0_STRIDE-th frame of the reference's output
CONFIG0 = ("noise", SR * 60, 1234)
CONFIG0_STRIDE = 37
N_FFTS = (1024, 960)


def make_audio(kind: str, n: int, seed: int, channels: int = 4) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "noise":
        x = 0.1 * rng.standard_normal((channels, n))
    elif kind == "loud":
        x = np.clip(0.5 * rng.standard_normal((channels, n)), -1.0, 1.0)
    elif kind == "zeros":
        x = np.zeros((channels, n))
    elif kind == "impulse_first":
        x = np.zeros((channels, n))
        x[:, 0] = 1.0
    elif kind == "impulse_last":
        x = np.zeros((channels, n))
        x[:, n - 1] = 1.0
    elif kind == "sine":
        x = np.zeros((channels, n))
        x[0] = 1e-4 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / SR)
    elif kind in ("level60", "level80", "level100"):
        x = 0.2 * rng.standard_normal((channels, n))
        gains = {"level60": (1e-3, 1.0, 1.0, 1e-3), "level80": (1.0, 1e-4, 1e-4, 1.0), "level100": (1e-5, 1.0, 1e-2, 1.0)}[kind]
        x *= np.asarray(gains)[:channels, None]
    elif kind == "ramp":
        x = 0.2 * rng.standard_normal((channels, n))
        fade = 10.0 ** (-5.0 * np.arange(n) / n)  # 0 dB -> -100 dB
        x[1] *= fade
        x[3] *= fade[::-1]
    elif kind == "int16":
        x = np.round(np.clip(0.1 * rng.standard_normal((channels, n)), -1, 1) * 32767.0) / 32768.0
    else:
        raise KeyError(kind)
    return np.ascontiguousarray(x.astype(np.float32))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------------------------
# metadata CSVs (STARSS format: frame100ms, class, source, azimuth, elevation[, distance])
# ---------------------------------------------------------------------------------------------
# name -> (csv file, n_samples of the matching audio)
LABEL_CASES = {
    "basic": ("basic.csv", 240000),      # 10 s -> T = 500
    "edges": ("edges.csv", 97440),       # 4.06 s -> int(4.06*1000/20) = 202 (not 203)
    "floatcol": ("floatcol.csv", 48000),  # 6th float column -> iterrows upcasts every cell to float64
    "weird": ("weird.csv", 48480),       # class-13 event, negative metadata frame, T % 5 != 0 (T = 101)
}
GAUSS_SEEDS = {"basic": 42, "edges": 43, "floatcol": 44}


def write_csvs() -> None:
    os.makedirs(CSV_DIR, exist_ok=True)
    rng = np.random.default_rng(0)
    # basic: 4 overlapping drifting sources over metadata frames 0..99
    rows = []
    srcs = [  # (class, source, start, stop, az0, el0, daz per frame, del per frame)
        (0, 0, 0, 80, -170, -30, 0.9, 0.3),
        (4, 0, 10, 100, 100, 50, -1.3, -0.5),
        (8, 1, 20, 60, 175, 0, 0.5, 0.0),
        (12, 0, 40, 95, -20, -80, 0.0, 0.4),
        (0, 1, 50, 70, 30, 20, 2.0, 0.1),
    ]
    for f in range(100):
        for (c, s, a, b, az0, el0, daz, de) in srcs:
            if a <= f < b:
                az = int(round(az0 + daz * (f - a)))
                az = (az + 180) % 360 - 180
                el = int(np.clip(round(el0 + de * (f - a)), -90, 90))
                rows.append((f, c, s, az, el))
    _write("basic.csv", rows)
    # edges
    rows = [
        (0, 1, 0, -180, -90), (0, 2, 0, 180, 90), (1, 3, 0, 179, 89), (1, 3, 1, -179, -89),
        (2, 5, 0, 14, 4), (2, 6, 0, 11, 9),      # two classes, one cell
        (3, 7, 0, 15, 5), (3, 7, 1, 19, 9),      # same class twice in one cell
        (4, 0, 0, 0, 0), (4, 0, 0, 0, 0),        # duplicate row
        (5, 9, 0, -98, -16),                     # notebook known-answer cell (7, 8)
        (39, 10, 0, 45, 45), (40, 10, 0, 45, 45),  # frame 40 -> start 200, clipped at T = 202
        (41, 11, 0, 50, 50), (60, 11, 0, 50, 50),  # rows past the audio end
        (7, 2, 0, 200, 100), (8, 2, 0, -300, -120),  # out-of-range angles clip
    ]
    _write("edges.csv", rows)
    # floatcol (STARSS23-like distance column)
    rows = []
    for f in range(0, 20):
        rows.append((f, int(rng.integers(0, 13)), 0, int(rng.integers(-180, 181)), int(rng.integers(-90, 91)),
                     float(np.round(rng.uniform(50, 400), 1))))
        if f % 3 == 0:
            rows.append((f, int(rng.integers(0, 13)), 1, int(rng.integers(-180, 181)), int(rng.integers(-90, 91)),
                         float(np.round(rng.uniform(50, 400), 1))))
    _write("floatcol.csv", rows)
    # weird
    rows = [
        (0, 13, 0, 10, 10), (0, 5, 0, 12, 12),  # class 13 (= background index) event sharing a cell
        (1, 13, 0, 100, 10),
        (-1, 3, 0, 0, 0),                       # negative metadata frame: t = -5..-1 wraps (T = 101)
        (20, 2, 0, -45, 30),                    # start 100, end min(105, 101)
        (2, -1, 0, 60, -60),                    # negative class index wraps to 13
    ]
    _write("weird.csv", rows)


def _write(name, rows):
    with open(os.path.join(CSV_DIR, name), "w") as f:
        for r in rows:
            f.write(",".join(str(v) for v in r) + "\n")


def csv_path(name: str) -> str:
    return os.path.join(CSV_DIR, LABEL_CASES[name][0])


# ---- loss cases (reference loss.py): logits and dense targets from seeds ------------------------------------------------
# name -> (B, T, I, J, events per frame, seed); frame 1 of every batch entry has no event at all, class 13 (the background
# index) appears as an event class, cells collect several classes (multi-hot rows)
LOSS_CASES = {"grid18x36": (2, 6, 18, 36, 7, 501), "grid6x12": (3, 5, 6, 12, 4, 502), "dense_events": (1, 4, 18, 36, 200, 503)}
LOSS_M = 14


def make_loss_case(name: str):
    """(logits (B, T, I*J, 14) float32, dense targets (B, T, I*J, 14) float32 built like dataset.py:100-117, I, J)."""
    B, T, I, J, n_ev, seed = LOSS_CASES[name]
    rng = np.random.default_rng(seed)
    G = I * J
    z = (3.0 * rng.standard_normal((B, T, G, LOSS_M))).astype(np.float32)
    y = np.zeros((B, T, G, LOSS_M), dtype=np.float32)
    for b in range(B):
        for t in range(T):
            if t == 1:
                continue
            cells = rng.integers(0, G, size=n_ev)
            cls = rng.integers(0, LOSS_M, size=n_ev)
            y[b, t, cells, cls] = 1.0
    y[..., LOSS_M - 1] = np.where(y.sum(-1) == 0, 1.0, y[..., LOSS_M - 1])  # cells without an event: one-hot background
    return z, y, I, J


def loss_mask(y: np.ndarray) -> np.ndarray:
    """The int16 class-set mask of dense targets: bit c <=> y[..., c] == 1; 0 for a one-hot background row."""
    bits = ((y != 0).astype(np.int64) << np.arange(y.shape[-1])).sum(-1)
    bits = np.where(bits == (1 << (y.shape[-1] - 1)), 0, bits)
    return bits.astype(np.uint16).view(np.int16)


def pack_labels(labels: np.ndarray) -> np.ndarray:
    """{0,1} float32 (T, G, M) -> packed bits; caller asserts the value set first."""
    return np.packbits((np.asarray(labels) != 0).reshape(-1))


def unpack_labels(bits: np.ndarray, shape) -> np.ndarray:
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(np.float32)
