"""How the two-channels-per-complex-FFT packing limits log-mel accuracy when the paired channels differ in level:
rounding noise of the loud channel leaks into the quiet one at about -125 dB (fp32).  Prints max |error| in dB of the
quiet channel against the fp64 oracle for level differences of 0..100 dB.  usage (GPU box): python tools/crosstalk_probe.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import seld_b200 as sb
from oracle import features as of

rng = np.random.default_rng(5)
n = 24000
base = rng.standard_normal((4, n)).astype(np.float32) * 0.2
for n_fft in (1024, 960):
    for L in (0, 20, 30, 40, 50, 60, 80, 100):
        x = base.copy()
        x[0] *= 10 ** (-L / 20)   # channel 0 (paired with 1) is L dB below its partner
        x[3] *= 10 ** (-L / 20)   # channel 3 (paired with 2)
        y = sb.extract_features(torch.from_numpy(x).cuda().unsqueeze(0), 24000, n_fft, 480, 64, mode="logmel")[0].cpu().numpy()
        ref = of.logmel(x, 24000, n_fft, 480, 64).transpose(2, 0, 1)
        e = np.abs(y - ref)
        print(f"n_fft {n_fft} level difference {L:3d} dB: quiet channels max|err| {max(e[:, 0].max(), e[:, 3].max()):.2e} dB, "
              f"loud channels {max(e[:, 1].max(), e[:, 2].max()):.2e} dB")
