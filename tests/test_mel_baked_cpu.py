"""The compile-time filterbank tables of the v3 kernel (csrc/mel_baked.h) are bit-exact copies of the table the
reference builds through torchaudio (dataset.py:38-43 -> melscale_fbanks), and regenerate identically."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "sound-event-localization-detection_b200", "csrc", "mel_baked.h")


def _parse(n_fft):
    txt = open(HDR).read()
    body = txt[txt.index(f"struct MelBaked<{n_fft}>"):]
    body = body[: body.index("};\n\n") + 2]

    def arr(name, conv):
        m = re.search(r"%s\[(\d+)\] = \{(.*?)\};" % name, body, re.S)
        vals = [v.strip() for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
        assert len(vals) == int(m.group(1))
        return [conv(v) for v in vals]

    fl = lambda v: float.fromhex(v[:-1]) if v.startswith(("0x", "-0x")) else float(v.rstrip("f"))
    return dict(m0=arr("m0", int), m1=arr("m1", int), w0=arr("w0", fl), w1=arr("w1", fl), first=arr("first", int),
                last=arr("last", int), chunk_m=arr("chunk_m", int), k0=arr("chunk_k0", int), k1=arr("chunk_k1", int))


@pytest.mark.parametrize("n_fft", [1024, 960])
def test_baked_tables_equal_reference_filterbank(golden_features, n_fft):
    t = _parse(n_fft)
    fb = golden_features[f"fb_{n_fft}"]  # dumped from torchaudio by tests/golden/make_golden.py
    nb = n_fft // 2 + 1
    assert fb.shape == (nb, 64) and fb.dtype == np.float32
    dense = np.zeros_like(fb)
    for k in range(nb):
        if t["m0"][k] >= 0:
            dense[k, t["m0"][k]] = np.float32(t["w0"][k])
        if t["m1"][k] >= 0:
            dense[k, t["m1"][k]] = np.float32(t["w1"][k])
    nz = fb != 0  # (torchaudio leaves a few -0.0 entries; they are zeros)
    assert np.array_equal(dense != 0, nz) and np.array_equal(dense[nz].view(np.uint32), fb[nz].view(np.uint32))
    # structure the kernel relies on: contiguous supports, filters two apart never overlap, chunks cover everything
    for m in range(64):
        nz = np.nonzero(fb[:, m])[0]
        assert nz[0] == t["first"][m] and nz[-1] == t["last"][m] and len(nz) == nz[-1] - nz[0] + 1
        if m + 2 < 64:
            assert t["last"][m] < t["first"][m + 2]
    assert t["chunk_m"][0] == 0 and t["chunk_m"][-1] == 64 and t["chunk_m"] == sorted(t["chunk_m"])
    for j in range(4):
        lo, hi = t["chunk_m"][j], t["chunk_m"][j + 1]
        assert t["k0"][j] == t["first"][lo] and t["k1"][j] == t["last"][hi - 1] + 1


def test_generator_reproduces_committed_header(tmp_path):
    pytest.importorskip("torchaudio")
    before = open(HDR).read()
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_mel_baked.py")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert open(HDR).read() == before
