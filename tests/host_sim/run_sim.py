"""Drive tests/host_sim/warp_sim (CPU emulation of the feature kernel) and compare with the oracle."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from oracle import features as of

def sim(x, n_fft, iv, window, fb, binary="/tmp/warp_sim"):
    C, N = x.shape
    NB = n_fft // 2 + 1
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fi, "wb") as f:
            f.write(window.astype(np.float32).tobytes()); f.write(fb.astype(np.float32).tobytes()); f.write(x.astype(np.float32).tobytes())
        r = subprocess.run([binary, str(n_fft), str(C), str(N), str(int(iv)), fi, fo], capture_output=True, text=True)
        if r.returncode: raise RuntimeError(r.stderr)
        T = 1 + N // 480
        C_out = 7 if iv else C
        raw = np.fromfile(fo, dtype=np.float32)
        feat = raw[: T * C_out * 64].reshape(T, C_out, 64)
        spec = raw[T * C_out * 64:].view(np.complex64).reshape(C, T, NB)
        return feat, spec, r.stderr

if __name__ == "__main__":
    g = np.load(os.path.join(ROOT, "tests/golden/features.npz"))
    worst = 0
    for n_fft in (1024, 960):
        win, fb = g[f"win_{n_fft}"], g[f"fb_{n_fft}"]
        for name, (kind, n, seed) in cases.AUDIO_CASES.items():
            x = cases.make_audio(kind, n, seed)
            feat, spec, log = sim(x, n_fft, True, win, fb)
            ref = g[f"{name}/logmel_{n_fft}"]          # (4, 64, T) from the real reference
            err = np.abs(feat[:, :4].transpose(1, 2, 0) - ref).max()
            X = of.stft(x, n_fft, 480)
            serr = np.abs(spec - X).max() / max(np.abs(X).max(), 1e-30)
            iv = of.foa_iv(x, 24000, n_fft, 480, 64, fb)
            iverr = np.abs(feat[:, 4:].transpose(1, 2, 0) - iv).max() / max(np.abs(iv).max(), 1e-30)
            print(f"n_fft={n_fft} {name:20s} logmel max|d|={err:.2e} dB  spec rel={serr:.2e}  iv rel={iverr:.2e}")
            worst = max(worst, err)
        for ch in (1, 2, 3, 6):
            if n_fft != 1024: continue
            x = cases.make_audio("noise", 4800, 100 + ch, channels=ch)
            feat, spec, log = sim(x, n_fft, False, win, fb)
            err = np.abs(feat.transpose(1, 2, 0) - g[f"noise_ch{ch}/logmel_1024"]).max()
            print(f"n_fft={n_fft} C={ch} logmel max|d|={err:.2e}")
            worst = max(worst, err)
    print(log.strip())
    print("worst logmel err", worst)
    sys.exit(0 if worst <= 1e-3 else 1)
