"""Differential fuzz of the v3 feature kernel against the generic kernel through the public API: random batch sizes,
ragged lengths, strides, frame capacities, silent channels, both n_fft.  usage (GPU box): python tools/fuzz_v3_vs_generic.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import seld_b200 as sb
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 123
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(seed)
worst = 0
for it in range(iters):
    n_fft = int(rng.choice([1024, 960]))
    B = int(rng.integers(1, 6))
    nmax = int(rng.integers(600, 60000))
    ns = [int(rng.integers(max(n_fft // 2 + 1, 520), nmax + 1)) for _ in range(B)]
    ns[int(rng.integers(0, B))] = nmax
    pad = int(rng.integers(0, 9)) * 4
    buf = torch.zeros((B, 4, nmax + pad), dtype=torch.float32)
    for i, n in enumerate(ns):
        buf[i, :, :n] = torch.from_numpy((rng.standard_normal((4, n)) * 10 ** rng.uniform(-4, 0)).astype(np.float32))
        if rng.random() < 0.3:
            buf[i, int(rng.integers(0, 4))] = 0
        if rng.random() < 0.2:  # a stretch of digital silence in all channels (a single almost-silent channel next to a
            s0 = int(rng.integers(0, n)); buf[i, :, s0:s0 + 3000] = 0  # loud one is the crosstalk case of tools/crosstalk_probe.py)
    audio = buf.cuda()[:, :, :nmax]
    lengths = torch.tensor(ns, dtype=torch.int64, device="cuda") if rng.random() < 0.8 else None
    if lengths is None:
        ns = [nmax] * B
    T_full = 1 + nmax // 480
    T_out = int(rng.integers(1, T_full + 1)) if rng.random() < 0.5 else T_full
    mode = str(rng.choice(["logmel", "logmel_iv"]))
    outs = {}
    for impl in ("v3", "v2"):
        os.environ["SELD_FEAT_IMPL"] = impl
        outs[impl] = sb.extract_features(audio, 24000, n_fft, 480, 64, mode=mode, lengths=lengths, T_out=T_out).cpu().numpy()
    a, b = outs["v3"], outs["v2"]
    assert a.shape == b.shape and np.isfinite(a).all(), (it, a.shape)
    d_db = np.abs(a[:, :, :4] - b[:, :, :4]).max()
    d_iv = np.abs(a[:, :, 4:] - b[:, :, 4:]).max() if mode == "logmel_iv" else 0.0
    for i, n in enumerate(ns):
        assert (a[i, 1 + n // 480:] == 0).all(), ("tail rows", it)
    worst = max(worst, d_db)
    if not (d_db <= 2e-4 and d_iv <= 2e-5):
        from oracle import features as of
        d = np.abs(a[:, :, :4] - b[:, :, :4])
        idx = np.unravel_index(d.argmax(), d.shape)
        i = idx[0]
        x = buf[i, :, :ns[i]].numpy()
        ref = (of.logmel_iv(x, 24000, n_fft, 480, 64) if mode == "logmel_iv" else of.logmel(x, 24000, n_fft, 480, 64)).transpose(2, 0, 1)
        print("FAIL", (it, n_fft, B, ns, T_out, mode, d_db, d_iv))
        print(" at (clip, frame, ch, mel)", idx, "v3", a[idx], "v2", b[idx], "oracle", ref[idx[1], idx[2], idx[3]])
        print(" frame samples", idx[1] * 480 - n_fft // 2, "..", idx[1] * 480 + n_fft // 2, "len", ns[i])
        xs = x[idx[2], max(0, idx[1] * 480 - n_fft // 2): idx[1] * 480 + n_fft // 2]
        print(" channel stats in frame: nonzero", int((xs != 0).sum()), "of", xs.size, "max", float(np.abs(xs).max()))
        sys.exit(1)
print("fuzz OK, worst dB diff", worst)
