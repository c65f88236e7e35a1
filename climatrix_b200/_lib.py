"""ctypes binding of libclimatrix_b200.so -- the only door to the CUDA kernels.

There is no CPU fallback: if the library is missing or no CUDA device is
present, the calls raise.
"""

from __future__ import annotations

import ctypes
import os
import re

from .build import LIB_PATH

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "climatrix_b200.h")

c_void_p, c_int, c_i64, c_double, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t

MAX_PEERS = 8


class Options(ctypes.Structure):
    """``cmx_options`` of include/climatrix_b200.h: everything optional about a call, passed explicitly."""

    _fields_ = [("search", c_int), ("n_peers", c_int), ("peer_out", c_void_p * MAX_PEERS), ("m_total", c_i64),
                ("m_offset", c_i64), ("params_on_device", c_int)]


SEARCH_AUTO, SEARCH_BRUTE, SEARCH_BINNED = 0, 1, 2
_OPT = ctypes.POINTER(Options)

# name -> (restype, argtypes)   [pointers are passed as integers / c_void_p]
_P = c_void_p
SIGNATURES = {
    "cmx_version": (c_int, []),
    "cmx_strerror": (ctypes.c_char_p, [c_int]),
    "cmx_last_cuda_error": (ctypes.c_char_p, []),
    "cmx_knn_workspace_bytes": (c_size_t, [c_int]),
    "cmx_knn_workspace_bytes_for": (c_size_t, [c_int, c_i64, c_i64, c_i64, c_int]),
    "cmx_knn": (c_int, [_P, _P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_idw_workspace_bytes": (c_size_t, [c_int, c_i64, c_i64]),
    "cmx_idw_f64": (c_int, [_P, _P, c_i64, _P, c_i64, c_int, _P, c_i64, _P, c_i64, c_int, c_int, c_int, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_idw_f32": (c_int, [_P, _P, c_i64, _P, c_i64, c_int, _P, c_i64, _P, c_i64, c_int, c_int, c_int, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_idw_apply_f64": (c_int, [_P, _P, c_i64, c_int, _P, c_i64, c_i64, c_int, c_int, c_int, c_double, _P, _P, _OPT, _P]),
    "cmx_idw_apply_f32": (c_int, [_P, _P, c_i64, c_int, _P, c_i64, c_i64, c_int, c_int, c_int, c_double, _P, _P, _OPT, _P]),
    "cmx_score_workspace_bytes": (c_size_t, [c_i64]),
    "cmx_idw_score_f64": (c_int, [_P, _P, c_i64, c_int, _P, c_i64, c_int, c_int, c_double, _P, _P, _P, _P, c_size_t, _P]),
    "cmx_idw_score_f32": (c_int, [_P, _P, c_i64, c_int, _P, c_i64, c_int, c_int, c_double, _P, _P, _P, _P, c_size_t, _P]),
    "cmx_score_f64": (c_int, [_P, _P, c_i64, _P, _P, c_size_t, _P]),
    "cmx_score_f32": (c_int, [_P, _P, c_i64, _P, _P, c_size_t, _P]),
    "cmx_variogram_minmax": (c_int, [_P, _P, c_i64, c_int, c_int, c_int, _P, _P]),
    "cmx_variogram_workspace_bytes": (c_size_t, [c_int]),
    "cmx_variogram_bins": (c_int, [_P, _P, _P, c_i64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "cmx_variogram_fit_host": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_double), c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
    "cmx_variogram_edges": (c_int, [_P, c_int, _P, _P]),
    "cmx_variogram_fit": (c_int, [c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cmx_ok_workspace_bytes": (c_size_t, [c_int, c_i64]),
    "cmx_ok_global_workspace_bytes": (c_size_t, [c_i64]),
    "cmx_ok_global_f64": (c_int, [_P, _P, _P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_double), c_int, c_int, c_double, c_double, _P, ctypes.POINTER(ctypes.c_int32), _P, c_size_t, _P]),
    "cmx_ok_global_f32": (c_int, [_P, _P, _P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_double), c_int, c_int, c_double, c_double, _P, ctypes.POINTER(ctypes.c_int32), _P, c_size_t, _P]),
    "cmx_ok_mw_f64": (c_int, [_P, _P, _P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_ok_mw_f32": (c_int, [_P, _P, _P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_timing_enable": (None, [c_int]),
    "cmx_timing_reset": (None, []),
    "cmx_timing_read": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_i64)]),
    "cmx_fp32_peak_tflops": (c_double, []),
    "cmx_fp64_peak_tflops": (c_double, []),
    "cmx_knn3": (c_int, [_P, _P, _P, c_i64, _P, _P, _P, c_i64, c_int, _P, _P, _P]),
    "cmx_ok_solve_workspace_bytes": (c_size_t, [c_i64]),
    "cmx_ok_solve_f64": (c_int, [_P, _P, _P, c_i64, _P, _P, c_i64, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_ok_solve_f32": (c_int, [_P, _P, _P, c_i64, _P, _P, c_i64, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P, _P, _P, _P, c_size_t, _OPT, _P]),
    "cmx_knn_plan": (c_int, [c_i64, c_i64, c_i64, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "cmx_copy_rows_to_host": (c_int, [_P, c_size_t, _P, c_size_t, c_size_t, c_size_t, _P]),
    "cmx_launch_count": (c_i64, []),
    "cmx_reset_launch_count": (None, []),
}

OK = 0
ERR_BAD_ARG, ERR_K_GT_N, ERR_K_TOO_LARGE, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED, ERR_TOO_MANY = -1, -2, -3, -4, -5, -6, -7
MAX_K = 128
LAYOUT_BATCH_MAJOR, LAYOUT_SAMPLE_MAJOR = 0, 1
METRIC_EUCLIDEAN, METRIC_GEOGRAPHIC = 0, 1
VARIOGRAM_MODELS = {"linear": 0, "power": 1, "gaussian": 2, "spherical": 3, "exponential": 4}

_lib = None


class ExtensionMissingError(ImportError):
    """libclimatrix_b200.so has not been built (run ``python -m climatrix_b200.build``)."""


def header_symbols() -> list[str]:
    """Every function name declared in include/climatrix_b200.h."""
    with open(_HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cmx_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissingError(
            f"{LIB_PATH} not found. climatrix_b200 has no CPU fallback: build the CUDA library with "
            "`python -m climatrix_b200.build` (needs nvcc, targets sm_100a)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == OK:
        return
    lib = load()
    msg = lib.cmx_strerror(rc).decode()
    if rc == ERR_K_GT_N:
        # the reference fails the same way: cKDTree pads with index N and values[idxs] raises (idw.py:167)
        raise IndexError(f"{what}: {msg}")
    if rc == ERR_CUDA:
        raise RuntimeError(f"{what}: {msg}: {lib.cmx_last_cuda_error().decode()}")
    if rc in (ERR_UNSUPPORTED,):
        raise NotImplementedError(f"{what}: {msg}")
    raise ValueError(f"{what}: {msg}")
