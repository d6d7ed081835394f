"""GPU parity: the CUDA feature path (through the C ABI) against the golden vectors produced by the real
reference and against the fp64 oracle.  Tolerances are north_star's: log-mel max abs error <= 1e-3 dB;
IV <= 1e-4 relative (to the largest |IV| of the clip; the reference has no IV code — parity unpinned)."""
import numpy as np
import pytest
import torch

import cases
from oracle import features as of

pytestmark = pytest.mark.gpu
TOL_DB = 1e-3
TOL_REL = 1e-4


@pytest.fixture(scope="module")
def sb():
    import seld_b200
    return seld_b200


def _run(sb, x, n_fft, mode="logmel", **kw):
    xt = torch.from_numpy(x).cuda()
    if xt.dim() == 2:
        xt = xt.unsqueeze(0)
    out = sb.extract_features(xt, 24000, n_fft, 480, 64, mode=mode, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
@pytest.mark.parametrize("name", list(cases.AUDIO_CASES))
def test_logmel_vs_reference_golden(sb, golden_features, name, n_fft):
    kind, n, seed = cases.AUDIO_CASES[name]
    x = cases.make_audio(kind, n, seed)
    y = sb.audio_to_mel_spectrogram(torch.from_numpy(x), cases.SR, n_fft=n_fft, hop_length=480, n_mels=64)
    ref = golden_features[f"{name}/logmel_{n_fft}"]
    assert y.shape == ref.shape and y.dtype == torch.float32 and not y.is_cuda and y.is_contiguous()
    assert np.abs(y.numpy() - ref).max() <= TOL_DB


def test_silence_exactly_minus_100(sb):
    y = sb.audio_to_mel_spectrogram(torch.zeros(4, 4800), 24000, n_fft=1024, hop_length=480, n_mels=64)
    assert (y == -100.0).all()


@pytest.mark.parametrize("ch", (1, 2, 3, 6))
def test_other_channel_counts(sb, golden_features, ch):
    x = cases.make_audio("noise", 4800, 100 + ch, channels=ch)
    y = sb.audio_to_mel_spectrogram(torch.from_numpy(x), 24000, n_fft=1024, hop_length=480, n_mels=64)
    assert np.abs(y.numpy() - golden_features[f"noise_ch{ch}/logmel_1024"]).max() <= TOL_DB


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
def test_stft_dump_vs_oracle(sb, n_fft):
    x = cases.make_audio("noise", 24000, 1234)
    T = 1 + 24000 // 480
    spec = torch.zeros((1, 4, T, n_fft // 2 + 1), dtype=torch.complex64, device="cuda")
    _run(sb, x, n_fft, spec=spec)
    X = of.stft(x, n_fft, 480)
    got = spec[0].cpu().numpy()
    assert np.abs(got - X).max() <= 2e-6 * np.abs(X).max()
    # Nyquist and DC bins of a real signal are real
    assert np.abs(got[..., 0].imag).max() == 0 and np.abs(got[..., -1].imag).max() == 0


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
@pytest.mark.parametrize("name", ["noise_1s", "noise_n97440", "int16_noise", "loud_noise", "impulse_last"])
def test_logmel_iv_vs_oracle(sb, golden_features, name, n_fft):
    kind, n, seed = cases.AUDIO_CASES[name]
    x = cases.make_audio(kind, n, seed)
    y = _run(sb, x, n_fft, mode="logmel_iv")[0]  # (T, 7, 64)
    want = of.logmel_iv(x, 24000, n_fft, 480, 64, fb=golden_features[f"fb_{n_fft}"]).transpose(2, 0, 1)
    assert y.shape == want.shape
    assert np.abs(y[:, :4] - want[:, :4]).max() <= TOL_DB
    scale = max(np.abs(want[:, 4:]).max(), 1e-12)
    assert np.abs(y[:, 4:] - want[:, 4:]).max() <= TOL_REL * scale
    # the 4 log-mel channels of the 7-channel mode equal the reference's own output as well
    assert np.abs(y[:, :4].transpose(1, 2, 0) - golden_features[f"{name}/logmel_{n_fft}"]).max() <= TOL_DB


def test_iv_plane_wave_direction(sb):
    rng = np.random.default_rng(4)
    s = rng.standard_normal(24000)
    az, el = np.deg2rad(40.0), np.deg2rad(-20.0)
    u = np.array([np.sin(az) * np.cos(el), np.sin(el), np.cos(az) * np.cos(el)])
    x = np.stack([s, u[0] * s, u[1] * s, u[2] * s]).astype(np.float32)
    y = _run(sb, x, 1024, mode="logmel_iv")[0]
    v = y[5:45, 4:, 10:60].mean(axis=(0, 2))
    assert np.allclose(v / np.linalg.norm(v), u, atol=1e-4)


def test_batch_ragged_lengths_and_strides(sb, golden_features):
    """Batch of clips with different valid lengths inside one padded buffer; rows past a clip's end are 0."""
    names = ["noise_n24001", "noise_n1000", "noise_n24479", "zeros"]
    ns = [cases.AUDIO_CASES[k][1] for k in names]
    nmax = max(ns)
    buf = torch.zeros((len(names), 4, nmax + 37), dtype=torch.float32)
    for i, k in enumerate(names):
        kind, n, seed = cases.AUDIO_CASES[k]
        buf[i, :, :n] = torch.from_numpy(cases.make_audio(kind, n, seed))
    audio = buf.cuda()[:, :, :nmax]  # non-contiguous clip/channel strides
    lengths = torch.tensor(ns, dtype=torch.int64, device="cuda")
    out = sb.extract_features(audio, 24000, 1024, 480, 64, mode="logmel", lengths=lengths).cpu().numpy()
    assert out.shape == (4, 1 + nmax // 480, 4, 64)
    for i, k in enumerate(names):
        ref = golden_features[f"{k}/logmel_1024"]
        T = ref.shape[2]
        assert np.abs(out[i, :T].transpose(1, 2, 0) - ref).max() <= TOL_DB
        assert (out[i, T:] == 0).all()


def test_cuda_input_returns_cuda_view(sb, golden_features):
    kind, n, seed = cases.AUDIO_CASES["noise_1s"]
    x = torch.from_numpy(cases.make_audio(kind, n, seed)).cuda()
    y = sb.audio_to_mel_spectrogram(x, 24000, n_fft=960, hop_length=480, n_mels=64)
    assert y.is_cuda and y.shape == (4, 64, 51)
    assert np.abs(y.cpu().numpy() - golden_features["noise_1s/logmel_960"]).max() <= TOL_DB


def test_scaler_stats_fused(sb):
    x = np.stack([cases.make_audio("noise", 48000, 50 + i) for i in range(3)])  # (3, 4, N)
    T = 1 + 48000 // 480
    stats = torch.zeros(2 * 7 * 64, dtype=torch.float64, device="cuda")
    stat_frames = torch.tensor([T - 1, T, 10], dtype=torch.int32, device="cuda")
    out = _run(sb, x, 1024, mode="logmel_iv", stats=stats, stat_frames=stat_frames)
    s = stats.cpu().numpy().reshape(2, 7 * 64)
    rows = np.concatenate([out[0, :T - 1], out[1, :T], out[2, :10]]).astype(np.float64).reshape(-1, 7 * 64)
    assert np.allclose(s[0], rows.sum(0), rtol=2e-6, atol=1e-3)
    assert np.allclose(s[1], (rows * rows).sum(0), rtol=2e-6, atol=1e-3)


def test_full_size_clip_config0(sb):
    """BASELINE.json configs[0]: one 60 s 4-ch clip, n_fft 1024 / hop 480, 64 mel + 3 IV vs the oracle."""
    rng = np.random.default_rng(1234)
    x = (0.1 * rng.standard_normal((4, 1_440_000))).astype(np.float32)
    y = _run(sb, x, 1024, mode="logmel_iv")[0]
    assert y.shape == (3001, 7, 64)
    want = of.logmel_iv(x, 24000, 1024, 480, 64).transpose(2, 0, 1)
    assert np.abs(y[:, :4] - want[:, :4]).max() <= TOL_DB
    assert np.abs(y[:, 4:] - want[:, 4:]).max() <= TOL_REL * np.abs(want[:, 4:]).max()


@pytest.mark.parametrize("n_fft", cases.N_FFTS)
def test_full_size_clip_config0_vs_reference_golden(sb, golden_features, n_fft):
    """BASELINE.json configs[0] pinned to the REFERENCE (not only to the oracle): the golden file keeps every
    CONFIG0_STRIDE-th frame of dataset.audio_to_mel_spectrogram's output for the 60 s clip."""
    kind, n, seed = cases.CONFIG0
    x = cases.make_audio(kind, n, seed)
    assert cases.sha(x) == bytes(golden_features["config0/sha"]).decode()
    y = sb.audio_to_mel_spectrogram(torch.from_numpy(x), cases.SR, n_fft=n_fft, hop_length=480, n_mels=64)
    assert y.shape == (4, 64, 3001)
    ref = golden_features[f"config0/logmel_{n_fft}_strided"]
    assert np.abs(y.numpy()[:, :, ::cases.CONFIG0_STRIDE] - ref).max() <= TOL_DB
    yi = _run(sb, x, n_fft, mode="logmel_iv")[0]  # the 7-channel mode's log-mel channels are the same numbers
    assert np.abs(yi[::cases.CONFIG0_STRIDE, :4].transpose(1, 2, 0) - ref).max() <= TOL_DB


def test_errors_are_exceptions(sb):
    with pytest.raises(sb.SeldError):
        sb.extract_features(torch.zeros(1, 4, 100, device="cuda"), 24000, 1024, 480, 64)  # N <= n_fft/2
    with pytest.raises(sb.SeldError):
        sb.get_plan(512, 480, 64, 24000, "cuda")  # unsupported n_fft
    with pytest.raises(sb.SeldError):
        sb.extract_features(torch.zeros(1, 3, 4800, device="cuda"), 24000, 1024, 480, 64, mode="logmel_iv")
    with pytest.raises((ValueError, sb.SeldError)):
        sb.extract_features(torch.zeros(1, 4, 4800), 24000, 1024, 480, 64)  # CPU tensor: no CPU path


# ---- MIC format: 4 log-mel + 6 GCC-PHAT (north_star kernel 3; no reference code — parity unpinned) ----
@pytest.mark.parametrize("n_fft", cases.N_FFTS)
@pytest.mark.parametrize("name", ["noise_1s", "noise_n97440", "int16_noise", "impulse_first", "zeros", "sine_1k_1e-4_ch0",
                                  "level_60db", "level_80db"])
def test_mic_gcc_vs_oracle(sb, golden_features, name, n_fft):
    """One fused launch: 4 log-mel + 6 GCC-PHAT channels, at n_fft 1024 and at the reference's default 960."""
    kind, n, seed = cases.AUDIO_CASES[name]
    x = cases.make_audio(kind, n, seed)
    y = _run(sb, x, n_fft, mode="logmel_gcc")[0]  # (T, 10, 64)
    want = of.mic_features(x, 24000, n_fft, 480, 64, fb=golden_features[f"fb_{n_fft}"]).transpose(2, 0, 1)
    assert y.shape == want.shape
    assert np.abs(y[:, :4] - want[:, :4]).max() <= TOL_DB
    assert np.abs(y[:, :4].transpose(1, 2, 0) - golden_features[f"{name}/logmel_{n_fft}"]).max() <= TOL_DB  # the reference itself
    # tolerance: 1e-4 relative to the largest |cc| of the frame (PHAT-normalised correlations peak at <= 1)
    scale = np.abs(want[:, 4:]).max(axis=(1, 2), keepdims=True)
    if kind == "sine":
        # channels 1..3 are exactly 0 -> R == 0 -> phase 1 -> delta at lag 0 for every pair
        assert np.abs(y[:, 4:] - want[:, 4:]).max() <= 1e-6 and (y[:, 4:, 32] > 0.999999).all()
    elif kind == "impulse_first":
        # an impulse at sample 0 is seen only by the first frames; later frames are digitally silent -> delta at lag 0
        assert (np.abs(y[3:, 4:] - want[3:, 4:]) <= TOL_REL).all()
    else:
        assert (np.abs(y[:, 4:] - want[:, 4:]) <= TOL_REL * np.maximum(scale, 1e-12)).all()


def test_mic_gcc_second_oracle(sb):
    """The CUDA GCC-PHAT against the second, independent restatement (torch.stft + torch.fft.irfft, oracle/ref_port.py)."""
    from oracle import ref_port
    x = cases.make_audio("noise", 24000 + 200, 91)
    for n_fft in cases.N_FFTS:
        y = _run(sb, x, n_fft, mode="logmel_gcc")[0][:, 4:]
        want = ref_port.gcc_phat_port(torch.from_numpy(x), n_fft, 480, 64).numpy().transpose(2, 0, 1)
        scale = np.abs(want).max(axis=(1, 2), keepdims=True)
        assert (np.abs(y - want) <= TOL_REL * scale).all()


def test_mic_ragged_batch_and_padding_rows(sb):
    rng = np.random.default_rng(13)
    ns = [24000, 5000, 24479]
    buf = torch.zeros((3, 4, max(ns)), dtype=torch.float32)
    for i, n in enumerate(ns):
        buf[i, :, :n] = torch.from_numpy((0.1 * rng.standard_normal((4, n))).astype(np.float32))
    lengths = torch.tensor(ns, dtype=torch.int64, device="cuda")
    T_out = 1 + max(ns) // 480 + 2
    out = sb.extract_features(buf.cuda(), 24000, 960, 480, 64, mode="logmel_gcc", lengths=lengths, T_out=T_out).cpu().numpy()
    for i, n in enumerate(ns):
        want = of.mic_features(buf[i, :, :n].numpy(), 24000, 960, 480, 64).transpose(2, 0, 1)
        T = want.shape[0]
        assert np.abs(out[i, :T, :4] - want[:, :4]).max() <= TOL_DB
        scale = np.abs(want[:, 4:]).max(axis=(1, 2), keepdims=True)
        assert (np.abs(out[i, :T, 4:] - want[:, 4:]) <= TOL_REL * scale).all()
        assert (out[i, T:] == 0).all()


def test_gcc_peak_at_delay(sb):
    rng = np.random.default_rng(3)
    s = rng.standard_normal(24000 + 80)
    d = 7
    x = np.stack([s[40:40 + 24000], s[40 - d:40 - d + 24000], s[40:40 + 24000], s[45:45 + 24000]]).astype(np.float32)
    y = _run(sb, x, 1024, mode="logmel_gcc")[0][5:45, 4:]  # (frames, 6, 64)
    assert (y[:, 0].argmax(-1) == 32 + d).all() and (y[:, 1].argmax(-1) == 32).all() and (y[:, 2].argmax(-1) == 27).all()


def test_gcc_needs_4_channels(sb):
    with pytest.raises(sb.SeldError):
        sb.extract_features(torch.zeros(1, 2, 4800, device="cuda"), 24000, 1024, 480, 64, mode="logmel_gcc")


def test_dead_channel_next_to_full_scale_signal(sb):
    """A silent channel packed with a loud one must still read exactly -100 dB (the reference transforms every
    channel separately); its IV contribution is exactly 0."""
    n = 9600
    x = np.zeros((4, n), np.float32)
    x[0] = np.sin(2 * np.pi * 3000.0 * np.arange(n) / 24000).astype(np.float32)
    x[3] = 0.9 * np.sin(2 * np.pi * 5000.0 * np.arange(n) / 24000).astype(np.float32)
    for n_fft in (1024, 960):
        y = _run(sb, x, n_fft, mode="logmel_iv")[0]
        assert (y[:, 1] == -100.0).all() and (y[:, 2] == -100.0).all()
        assert (y[:, 4] == 0).all() and (y[:, 5] == 0).all()
        want = of.logmel_iv(x, 24000, n_fft, 480, 64).transpose(2, 0, 1)
        # full-scale pure tones: mel bands more than ~70 dB below the peak sit on the float32 rounding floor of
        # ANY float32 FFT (the reference's too), so parity is only meaningful above it
        loud = want[:, :4] > want[:, :4].max() - 70.0
        assert np.abs(y[:, :4] - want[:, :4])[loud].max() <= TOL_DB
