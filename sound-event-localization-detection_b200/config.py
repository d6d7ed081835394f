"""Configuration mirror of the reference's ``config.py`` (config.py:1-118): same attribute names and
values for everything the feature/label path reads, plus the front-end's own optional switches.

When this package runs inside the reference tree (drop-in under main.py / trainer.py) the reference's own
``Config`` is used so that a user's edits there are honoured; otherwise the defaults below apply.  Unlike
the reference, constructing it has no filesystem side effects."""
from __future__ import annotations

from pathlib import Path


class Config:
    BASE_PATH = Path.cwd()
    AUDIO_PATH = BASE_PATH / "foa_dev"
    METADATA_PATH = BASE_PATH / "metadata_dev"
    USE_FULL_DATASET = True
    TRAIN_AUDIO_FILE = "fold3_room21_mix001.wav"
    TRAIN_META_FILE = "fold3_room21_mix001.csv"
    TEST_AUDIO_FILE = "fold4_room23_mix001.wav"
    TEST_META_FILE = "fold4_room23_mix001.csv"

    NUM_CLASSES = 14
    N_CHANNELS = 4
    BATCH_SIZE = 16

    # Signal processing (config.py:84-88)
    SPECTROGRAM_N_FFT = int(0.04 * 24000)        # 960
    SPECTROGRAM_HOP_LENGTH = int(0.02 * 24000)   # 480
    N_MELS = 64
    SR = 24000
    # Dataset windowing (config.py:90-92)
    WINDOW_LENGTH = int(5 * 24000)
    HOP_LENGTH = int(1 * 24000)
    # Grid (config.py:94-97)
    I = None
    J = None
    GRID_CELL_DEGREES = 10

    # ---- additions of this front-end (absent from the reference; defaults keep reference behaviour) ----
    FEATURE_TYPE = "logmel"      # "logmel" (reference), "foa_iv" (7 ch), "mic_gcc" (10 ch)

    def __init__(self):
        self.SONY_TRAIN_DIR = self.AUDIO_PATH / "dev-train-sony"
        self.SONY_TEST_DIR = self.AUDIO_PATH / "dev-test-sony"
        self.SONY_TRAIN_META_DIR = self.METADATA_PATH / "dev-train-sony"
        self.SONY_TEST_META_DIR = self.METADATA_PATH / "dev-test-sony"
        self.TAU_TRAIN_DIR = self.AUDIO_PATH / "dev-train-tau"
        self.TAU_TEST_DIR = self.AUDIO_PATH / "dev-test-tau"
        self.TAU_TRAIN_META_DIR = self.METADATA_PATH / "dev-train-tau"
        self.TAU_TEST_META_DIR = self.METADATA_PATH / "dev-test-tau"
        self.TRAIN_AUDIO_PATH = self.AUDIO_PATH / "dev-train-sony" / self.TRAIN_AUDIO_FILE
        self.TRAIN_META_PATH = self.METADATA_PATH / "dev-train-sony" / self.TRAIN_META_FILE
        self.TEST_AUDIO_PATH = self.AUDIO_PATH / "dev-test-sony" / self.TEST_AUDIO_FILE
        self.TEST_META_PATH = self.METADATA_PATH / "dev-test-sony" / self.TEST_META_FILE


_config = None


def get_config():
    """The active configuration: the reference's ``config.Config()`` when importable (drop-in use inside the
    reference tree), else this module's mirror."""
    global _config
    if _config is None:
        try:
            from config import Config as RefConfig  # type: ignore  # the reference's top-level module
            if RefConfig is Config or not hasattr(RefConfig, "SPECTROGRAM_N_FFT"):
                raise ImportError
            _config = RefConfig()
        except Exception:
            _config = Config()
    return _config


def set_config(cfg) -> None:
    global _config
    _config = cfg
