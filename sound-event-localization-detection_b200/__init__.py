"""seld_b200 — B200-native SELD feature front-end (drop-in for the data path of
Zeudon/sound-event-localization-detection: dataset.py + smrl_seld_gaussian.augment_with_gaussian_noise).

The directory is named ``sound-event-localization-detection_b200``; import it as ``seld_b200`` (the
top-level ``seld_b200.py`` shim) or via ``importlib.import_module``."""
from . import _lib, config, features  # noqa: F401
from ._lib import LIB_PATH, SeldError  # noqa: F401
from .features import (FeaturePlan, audio_to_mel_spectrogram, extract_features, get_plan, hann_window,  # noqa: F401
                       mel_filterbank)

__all__ = ["FeaturePlan", "audio_to_mel_spectrogram", "extract_features", "get_plan", "hann_window",
           "mel_filterbank", "SeldError", "LIB_PATH"]
