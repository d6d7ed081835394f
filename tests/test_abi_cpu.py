"""CPU: the C-ABI library loads and exports every symbol include/seld_cuda.h declares; host-side tables."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import seld_b200
    hdr = open(os.path.join(ROOT, "include", "seld_cuda.h")).read()
    declared = set(re.findall(r"SELD_API\s+[\w\s\*]+?\b(seld_\w+)\s*\(", hdr))
    assert declared and declared == set(seld_b200._lib.EXPORTS)
    l = ctypes.CDLL(seld_b200.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), name
    assert seld_b200._lib.lib().seld_version() >= 100


def test_pure_host_entry_points():
    import seld_b200
    l = seld_b200._lib.lib()
    assert l.seld_num_frames(2_145_600, 480) == 4471      # SMR_SELD_2.ipynb:518-519
    assert l.seld_num_frames(1_440_000, 480) == 3001
    assert l.seld_out_channels(0, 4) == 4 and l.seld_out_channels(1, 4) == 7 and l.seld_out_channels(2, 4) == 10


def test_no_cpu_fallback():
    import torch
    import seld_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((seld_b200.SeldError, RuntimeError, AssertionError)):
        seld_b200.audio_to_mel_spectrogram(torch.zeros(4, 4800), 24000, n_fft=1024, hop_length=480, n_mels=64)


@pytest.mark.parametrize("n_fft", (1024, 960))
def test_tables_bit_identical_to_torchaudio(golden_features, n_fft):
    import seld_b200
    assert np.array_equal(seld_b200.mel_filterbank(n_fft, 24000, 64).numpy(), golden_features[f"fb_{n_fft}"])
    assert np.array_equal(seld_b200.hann_window(n_fft).numpy(), golden_features[f"win_{n_fft}"])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sound-event-localization-detection_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
