"""seld_b200 — B200-native SELD feature front-end (drop-in for the data path of
Zeudon/sound-event-localization-detection: dataset.py + smrl_seld_gaussian.augment_with_gaussian_noise).

The directory is named ``sound-event-localization-detection_b200``; import it as ``seld_b200`` (the
top-level ``seld_b200.py`` shim) or via ``importlib.import_module``."""
from . import _lib, audio_io, config, dataset, features, labels, loss, scaler  # noqa: F401
from ._lib import LIB_PATH, SeldError  # noqa: F401
from .audio_io import load_audio  # noqa: F401
from .dataset import DeviceLoader, SELDDataset, load_files, shard_clips, shard_files  # noqa: F401
from .features import (FeaturePlan, audio_to_mel_spectrogram, extract_features, extract_features_host,  # noqa: F401
                       get_plan, hann_window, mel_filterbank)
from .labels import augment_with_gaussian_noise, metadata_to_labels, polar_to_grid  # noqa: F401
from .loss import CompactSMRSELDLoss  # noqa: F401
from .scaler import FeatureScaler  # noqa: F401

__all__ = ["FeaturePlan", "audio_to_mel_spectrogram", "extract_features", "extract_features_host", "get_plan",
           "hann_window", "mel_filterbank", "metadata_to_labels", "augment_with_gaussian_noise", "polar_to_grid",
           "SELDDataset", "DeviceLoader", "load_files", "shard_clips", "shard_files", "load_audio", "FeatureScaler", "CompactSMRSELDLoss", "SeldError", "LIB_PATH"]
