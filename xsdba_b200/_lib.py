"""ctypes binding of the C ABI declared in include/xsdba_b200.h.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded, importing callers get
a loud RuntimeError telling them to build it (``python -m xsdba_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxsdba_b200.so")

_lock = threading.Lock()
_lib = None

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
vp = C.c_void_p
i64 = C.c_int64
i32 = C.c_int32

KIND = {"+": 43, "*": 42}
INTERP = {"nearest": 0, "linear": 1, "cubic": 2}
EXTRAP = {"constant": 0, "nan": 1}

# name -> (restype, argtypes).  Data pointers are passed as void* (device or host addresses).
SIGNATURES = {
    "xsdba_version": (C.c_int, []),
    "xsdba_status_string": (C.c_char_p, [C.c_int]),
    "xsdba_launch_count": (i64, []),
    "xsdba_grouping_create": (C.c_int, [C.POINTER(vp), c_i32p, i64, i32, i32]),
    "xsdba_grouping_destroy": (C.c_int, [vp]),
    "xsdba_grouping_max_segment": (i64, [vp]),
    "xsdba_grouping_n_groups": (i32, [vp]),
}
for _t in ("f32", "f64"):
    SIGNATURES[f"xsdba_qm_train_{_t}"] = (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, i32, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_qm_train_jitter_{_t}"] = (
        C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, i32, c_f64p, C.c_uint64, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_jitter_{_t}"] = (C.c_int, [vp, i64, c_f64p, C.c_uint64, vp, vp])
    SIGNATURES[f"xsdba_group_quantile_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, i32, vp, vp])
    SIGNATURES[f"xsdba_qm_adjust_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, i32, vp, vp])
    SIGNATURES[f"xsdba_qdm_adjust_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp])
    SIGNATURES[f"xsdba_poly_trend_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, i32, i32, vp, vp, vp])
    SIGNATURES[f"xsdba_loess_trend_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, i32, C.c_double, i32, i32, vp, vp, vp])
    SIGNATURES[f"xsdba_loess_trend_w_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, i32, C.c_double, i32, i32, i32, i32, vp, vp, vp])
    SIGNATURES[f"xsdba_dqm_adjust_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp])
    SIGNATURES[f"xsdba_rank_lookup_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp])
    SIGNATURES[f"xsdba_rotate_{_t}"] = (C.c_int, [vp, i64, i32, c_f32p, vp, vp])
    SIGNATURES[f"xsdba_rotate_unfused_{_t}"] = (C.c_int, [vp, i64, i32, c_f32p, vp, vp])
    SIGNATURES[f"xsdba_standardize_{_t}"] = (C.c_int, [vp, i64, i64, i64, i64, i32, i64, vp, vp])
    SIGNATURES[f"xsdba_reorder_{_t}"] = (C.c_int, [vp, vp, i64, i64, i64, vp, vp, vp])
    SIGNATURES[f"xsdba_group_vecquantile_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_map_cdf_{_t}"] = (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, vp, vp])
    SIGNATURES[f"xsdba_qm_train_adapt_{_t}"] = (
        C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, c_f64p, C.c_double, C.c_uint64, vp, vp, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_dqm_train_adapt_{_t}"] = (
        C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, c_f64p, C.c_double, C.c_uint64, vp, vp, vp, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_adapt_freq_apply_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, C.c_double, vp, vp, vp, C.c_uint64, vp, vp])
    SIGNATURES[f"xsdba_tail_mask_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, i32, C.c_double, vp, vp])
    SIGNATURES[f"xsdba_qdm_adjust_linear_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp])
    SIGNATURES[f"xsdba_escore_{_t}"] = (C.c_int, [vp, vp, i64, i64, i64, i64, i64, i32, i64, i64, i32, vp, vp])
    SIGNATURES[f"xsdba_group_rank_{_t}"] = (C.c_int, [vp, i64, i64, i64, vp, i32, vp, vp])
SIGNATURES["xsdba_qm_train_q64_f32"] = (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, vp, vp, vp])
SIGNATURES["xsdba_npdft_step_f32"] = (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32, i32, i32, vp, vp])
SIGNATURES["xsdba_debug_copy_rows_f32"] = (C.c_int, [vp, i64, i64, vp, vp, i32, vp])
SIGNATURES["xsdba_qm_train_adjust_host_f32"] = (
    C.c_int, [vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, i64])
SIGNATURES["xsdba_qm_train_adjust_host_workspace_bytes"] = (i64, [i64, vp, vp, i32, i32, i32, i64])
for _t in ("f32", "f64"):
    SIGNATURES[f"xsdba_qm_train_adjust_host_ws_{_t}"] = (
        C.c_int, [vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, i32, c_f64p, vp, vp, vp, vp, i64, vp, i64])


class XsdbaB200Error(RuntimeError):
    pass


def load():
    """Load libxsdba_b200.so (once) and declare every entry point."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise XsdbaB200Error(
                f"{LIB_PATH} is missing: xsdba_b200 has no CPU fallback. Build it with "
                "`python -m xsdba_b200.build` (needs nvcc, targets sm_100a).")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise XsdbaB200Error(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def check(status: int, what: str = ""):
    """Turn a status code into the exceptions the reference raises (ValueError / NotImplementedError)."""
    if status == 0:
        return
    msg = load().xsdba_status_string(status).decode()
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    if status < 0:
        raise ValueError(f"{what}: {msg}")
    raise XsdbaB200Error(f"{what}: CUDA error {status}: {msg}")
