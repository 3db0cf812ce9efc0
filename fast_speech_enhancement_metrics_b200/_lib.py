"""ctypes binding of libfsem_b200.so (C ABI declared in include/fsem.h).

There is no fallback: if the shared library is missing or does not load, importing
a metric raises.  Build it with `python -m fast_speech_enhancement_metrics_b200.build`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

from .design import PesqDesign, StoiDesign

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSEM_B200_LIB", os.path.join(HERE, "libfsem_b200.so"))

FSEM_OK = 0
FSEM_E_INVALID = -1
FSEM_E_CUDA = -2
FSEM_E_WORKSPACE = -3
FSEM_E_TOO_SHORT = -4
ITEM_OK, ITEM_TOO_SHORT, ITEM_NAN = 0, 1, 2

# every symbol include/fsem.h declares
EXPORTS = [
    "fsem_version", "fsem_last_error", "fsem_launch_count",
    "fsem_profile_enable", "fsem_profile_reset", "fsem_profile_read",
    "fsem_pesq_create", "fsem_pesq_destroy", "fsem_pesq_workspace_bytes", "fsem_pesq_score_f32",
    "fsem_pesq_score_host_f32", "fsem_pesq_debug_taps",
    "fsem_stoi_create", "fsem_stoi_destroy", "fsem_stoi_workspace_bytes", "fsem_stoi_score_f32",
    "fsem_stoi_score_host_f32", "fsem_stoi_debug_taps", "fsem_pesq_stoi_score_host_f32",
    "fsem_pesq_stoi_score_f32", "fsem_lsd_create", "fsem_lsd_destroy", "fsem_lsd_workspace_bytes", "fsem_lsd_score_f32",
    "fsem_sdr_workspace_bytes", "fsem_sdr_score_f32", "fsem_ingest_f32", "fsem_score_host",
    "fsem_stoi_mask_margin", "fsem_resampler_create", "fsem_resampler_destroy", "fsem_resampled_len",
    "fsem_resample_f32", "fsem_pesq_score", "fsem_stoi_score", "fsem_pesq_stoi_score",
    "fsem_graph_create", "fsem_graph_launch", "fsem_graph_nodes", "fsem_graph_slices", "fsem_graph_destroy",
]

# ingest formats (include/fsem.h FSEM_DTYPE_*)
DTYPE_F32, DTYPE_I16, DTYPE_F16 = 0, 1, 2


def dtype_code(dtype) -> int:
    """torch dtype -> FSEM_DTYPE_*; raises like the reference for anything else (float64 fails inside
    lfilter / stft there with "expected scalar type Float")."""
    import torch
    codes = {torch.float32: DTYPE_F32, torch.int16: DTYPE_I16, torch.float16: DTYPE_F16}
    if dtype not in codes:
        raise RuntimeError("expected scalar type Float but found %s" % str(dtype).replace("torch.", ""))
    return codes[dtype]


class Batch(C.Structure):
    """Mirror of fsem_batch_t."""
    _fields_ = [
        ("clean", C.c_void_p), ("deg", C.c_void_p), ("lengths", C.c_void_p),
        ("batch", C.c_int64), ("n", C.c_int64), ("stride", C.c_int64),
    ]


class FsemError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libfsem_b200: %s (code %d)" % (message, code))
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: the CUDA extension has not been built (python -m "
            "fast_speech_enhancement_metrics_b200.build). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32p, fp = C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p
    lib.fsem_version.restype = C.c_int
    lib.fsem_last_error.restype = C.c_char_p
    lib.fsem_launch_count.restype = C.c_int64
    lib.fsem_profile_enable.argtypes = [C.c_int]
    lib.fsem_profile_read.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.fsem_pesq_create.argtypes = [C.POINTER(vp), C.POINTER(PesqDesign)]
    lib.fsem_pesq_destroy.argtypes = [vp]
    lib.fsem_pesq_workspace_bytes.argtypes = [vp, i64, i64]
    lib.fsem_pesq_workspace_bytes.restype = C.c_size_t
    lib.fsem_pesq_score_f32.argtypes = [vp, C.POINTER(Batch), fp, i32p, vp, C.c_size_t, vp]
    lib.fsem_pesq_score_host_f32.argtypes = [vp, C.POINTER(Batch), fp, i32p]
    lib.fsem_pesq_debug_taps.argtypes = [vp, i64, i64, vp, fp, vp, C.POINTER(i64), vp]
    lib.fsem_stoi_create.argtypes = [C.POINTER(vp), C.POINTER(StoiDesign)]
    lib.fsem_stoi_destroy.argtypes = [vp]
    lib.fsem_stoi_workspace_bytes.argtypes = [vp, i64, i64]
    lib.fsem_stoi_workspace_bytes.restype = C.c_size_t
    lib.fsem_stoi_score_f32.argtypes = [vp, C.POINTER(Batch), fp, fp, i32p, i32p, vp, C.c_size_t, vp]
    lib.fsem_stoi_score_host_f32.argtypes = [vp, C.POINTER(Batch), fp, fp, i32p, i32p]
    lib.fsem_pesq_stoi_score_host_f32.argtypes = [vp, vp, C.POINTER(Batch), fp, i32p, fp, fp, i32p, i32p]
    lib.fsem_pesq_stoi_score_f32.argtypes = [vp, vp, C.POINTER(Batch), fp, i32p, fp, fp, i32p, i32p, vp, C.c_size_t, vp,
                                             C.c_size_t, vp]
    lib.fsem_pesq_score.argtypes = [vp, C.POINTER(Batch), C.c_int, fp, i32p, vp, C.c_size_t, vp]
    lib.fsem_stoi_score.argtypes = [vp, C.POINTER(Batch), C.c_int, fp, fp, i32p, i32p, vp, C.c_size_t, vp]
    lib.fsem_pesq_stoi_score.argtypes = [vp, vp, C.POINTER(Batch), C.c_int, fp, i32p, fp, fp, i32p, i32p, vp, C.c_size_t,
                                         vp, C.c_size_t, vp]
    lib.fsem_graph_create.argtypes = [C.POINTER(vp), vp, vp, C.POINTER(Batch), C.c_int, fp, i32p, fp, fp, i32p, i32p, C.c_int]
    lib.fsem_graph_launch.argtypes = [vp, vp]
    lib.fsem_graph_nodes.argtypes = [vp]
    lib.fsem_graph_slices.argtypes = [vp]
    lib.fsem_graph_destroy.argtypes = [vp]
    lib.fsem_stoi_mask_margin.argtypes = [vp, i64, i64, i32p, vp, fp, vp]
    lib.fsem_resampler_create.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.POINTER(C.c_float)]
    lib.fsem_resampler_destroy.argtypes = [vp]
    lib.fsem_resampled_len.argtypes = [vp, i64]
    lib.fsem_resampled_len.restype = C.c_int64
    lib.fsem_resample_f32.argtypes = [vp, C.POINTER(Batch), fp, i64, i32p, vp]
    lib.fsem_lsd_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_float)]
    lib.fsem_lsd_destroy.argtypes = [vp]
    lib.fsem_lsd_workspace_bytes.argtypes = [vp, i64, i64]
    lib.fsem_lsd_workspace_bytes.restype = C.c_size_t
    lib.fsem_lsd_score_f32.argtypes = [vp, C.POINTER(Batch), fp, vp, C.c_size_t, vp]
    lib.fsem_sdr_workspace_bytes.argtypes = [i64, i64]
    lib.fsem_sdr_workspace_bytes.restype = C.c_size_t
    lib.fsem_sdr_score_f32.argtypes = [C.POINTER(Batch), fp, vp, C.c_size_t, vp]
    lib.fsem_stoi_debug_taps.argtypes = [vp, i64, i64, vp, vp, fp, fp, C.POINTER(i64), vp]
    lib.fsem_ingest_f32.argtypes = [vp, C.c_int, i64, i64, i64, fp, i64, vp]
    lib.fsem_score_host.argtypes = [vp, vp, vp, vp, C.c_int, i32p, i64, i64, i64, fp, i32p, fp, fp, i32p, i32p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("fsem_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != FSEM_OK:
        msg = load().fsem_last_error()
        raise FsemError(code, msg.decode("utf-8", "replace") if msg else "unknown error")


def launch_count() -> int:
    return int(load().fsem_launch_count())


def profile_enable(on: bool) -> None:
    load().fsem_profile_enable(1 if on else 0)


def profile_reset() -> None:
    load().fsem_profile_reset()


def profile_read() -> dict:
    """{kernel name: (total device ms, launches)} accumulated since the last reset."""
    lib, out = load(), {}
    for i in range(16):
        name, ms, cnt = C.c_char_p(), C.c_double(), C.c_int64()
        if lib.fsem_profile_read(i, C.byref(name), C.byref(ms), C.byref(cnt)) != FSEM_OK:
            break
        out[name.value.decode()] = (ms.value, cnt.value)
    return out
