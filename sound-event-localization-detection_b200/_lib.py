"""ctypes binding of libseld_cuda.so (C ABI in include/seld_cuda.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with ``python __graft_entry__.py`` (or ``make -C <package>/csrc``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SELD_CUDA_LIB: another build of the library (developer A/B runs, tools/build_variant.sh); the default is the in-tree one
LIB_PATH = os.environ.get("SELD_CUDA_LIB") or os.path.join(_HERE, "libseld_cuda.so")

SELD_MODE_LOGMEL, SELD_MODE_LOGMEL_IV, SELD_MODE_LOGMEL_GCC = 0, 1, 2
SELD_DTYPE_F32, SELD_DTYPE_I16, SELD_DTYPE_BF16 = 0, 1, 2
SELD_LAYOUT_TCF, SELD_LAYOUT_CTF = 0, 1


class FeatOpts(C.Structure):
    """``seld_feat_opts`` (include/seld_cuda.h)."""
    _fields_ = [("in_dtype", C.c_int), ("out_layout", C.c_int), ("out_dtype", C.c_int), ("d_mean", C.c_void_p),
                ("d_inv_std", C.c_void_p)]


class SeldError(RuntimeError):
    """A libseld_cuda call returned a negative status."""


_lib = None

_SIGS = {
    "seld_version": (C.c_int, []),
    "seld_last_error": (C.c_char_p, []),
    "seld_num_frames": (C.c_int64, [C.c_int64, C.c_int]),
    "seld_out_channels": (C.c_int, [C.c_int, C.c_int]),
    "seld_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "seld_plan_destroy": (C.c_int, [C.c_void_p]),
    "seld_features": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "seld_features_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(FeatOpts), C.c_void_p]),
    "seld_plan_status": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "seld_plan_has_fast_path": (C.c_int, [C.c_void_p]),
    "seld_feature_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "seld_scaler_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "seld_pcm16_to_float": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "seld_labels_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "seld_labels_paint": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p]),
    "seld_window_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "seld_loader_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "seld_batch_class_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "seld_class_loss": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "seld_aux_losses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built (run `python __graft_entry__.py` "
                "or `make -C sound-event-localization-detection_b200/csrc`).  There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().seld_last_error()
        raise SeldError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
