"""ctypes binding of ``libsnapb200.so`` (the C ABI in ``include/snapb200.h``).

The library is built in-tree by ``snapatac2_b200.build``; there is no Python
or CPU fallback -- if the shared object is missing or no B200 is present the
calls fail loudly.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libsnapb200.so"

# every symbol include/snapb200.h declares (tests check the export list)
SYMBOLS = [
    "snapb200_last_error", "snapb200_version", "snapb200_create", "snapb200_destroy",
    "snapb200_comm_unique_id", "snapb200_comm_init", "snapb200_load_csr",
    "snapb200_load_begin", "snapb200_load_append", "snapb200_load_end", "snapb200_set_geometry",
    "snapb200_set_defer_value_scan", "snapb200_values_verdict", "snapb200_load_values",
    "snapb200_select_features", "snapb200_generate", "snapb200_shape", "snapb200_export_csr",
    "snapb200_set_feature_weights", "snapb200_prepare", "snapb200_view_norms",
    "snapb200_attach_view", "snapb200_view_frobenius", "snapb200_combine_views", "snapb200_get_vector",
    "snapb200_gather_rows",
    "snapb200_prepare_projection", "snapb200_project",
    "snapb200_operator_apply", "snapb200_operator_time", "snapb200_eigsh", "snapb200_get_stats",
    "snapb200_get_stream", "snapb200_set_spmm_mode", "snapb200_set_block",
    "snapb200_dense_selftest", "snapb200_ortho_selftest", "snapb200_delta_selftest_host", "snapb200_knn", "snapb200_knn_limits", "snapb200_sym_eig",
]


class Stats(C.Structure):
    _fields_ = [
        ("ms_load", C.c_double), ("ms_transpose", C.c_double), ("ms_prepare", C.c_double),
        ("ms_eigsh", C.c_double), ("ms_spmm", C.c_double), ("ms_ortho", C.c_double),
        ("ms_comm", C.c_double), ("ms_host", C.c_double), ("max_residual", C.c_double),
        ("n_ops", C.c_int64), ("n_restarts", C.c_int64), ("basis_cols", C.c_int64),
        ("block", C.c_int64), ("nnz_local", C.c_int64), ("kernel_launches", C.c_int64),
        ("ms_format", C.c_double), ("spmm_tiled", C.c_int64),
        ("ms_prepare_wall", C.c_double),
        ("ms_pool", C.c_double), ("pool_mallocs", C.c_int64),
        ("converged", C.c_int64), ("n_spec_ops", C.c_int64), ("ms_d2h", C.c_double),
        ("bytes_h2d", C.c_int64), ("host_threads", C.c_int64), ("fused_allreduce", C.c_int64), ("bytes_h2d_indices", C.c_int64),
        ("ms_knn", C.c_double), ("ms_knn_wall", C.c_double), ("knn_mma", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m snapatac2_b200.build` "
            "(snapatac2_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    lib.snapb200_last_error.restype = C.c_char_p
    lib.snapb200_last_error.argtypes = []
    lib.snapb200_version.restype = i32
    sigs = {
        "snapb200_create": [i32, C.POINTER(vp)],
        "snapb200_destroy": [vp],
        "snapb200_comm_unique_id": [C.c_char_p],
        "snapb200_comm_init": [vp, i32, i32, C.c_char_p],
        "snapb200_load_csr": [vp, i64, i64, i64, i64, vp, i32, vp, i32, vp, i32, i32],
        "snapb200_load_begin": [vp, i64, i64, i64],
        "snapb200_load_append": [vp, i64, vp, i32, vp, i32, vp, i32],
        "snapb200_load_end": [vp, i64, i64],
        "snapb200_set_geometry": [vp, i64, i64],
        "snapb200_set_defer_value_scan": [vp, i32],
        "snapb200_values_verdict": [vp, C.POINTER(i32)],
        "snapb200_load_values": [vp, vp, i32],
        "snapb200_select_features": [vp, vp, i64],
        "snapb200_generate": [vp, i64, i64, i64, i64, i32, i32, C.c_uint64, vp, vp, vp, vp],
        "snapb200_shape": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
        "snapb200_export_csr": [vp, vp, vp, vp],
        "snapb200_set_feature_weights": [vp, vp, i64],
        "snapb200_prepare": [vp, vp, vp],
        "snapb200_view_norms": [vp, vp, vp],
        "snapb200_attach_view": [vp, vp],
        "snapb200_view_frobenius": [vp, vp, i64, C.POINTER(dbl)],
        "snapb200_combine_views": [vp, vp, vp, i32, vp],
        "snapb200_get_vector": [vp, i32, vp],
        "snapb200_gather_rows": [vp, vp, i64, vp, i64, i64],
        "snapb200_prepare_projection": [vp, vp, vp],
        "snapb200_project": [vp, i32, vp, i32, vp],
        "snapb200_operator_apply": [vp, vp, vp, i32],
        "snapb200_operator_time": [vp, i32, i32, i32, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)],
        "snapb200_eigsh": [vp, i32, i64, dbl, i32, i32, i32, vp, vp, i32],
        "snapb200_get_stats": [vp, C.POINTER(Stats)],
        "snapb200_get_stream": [vp, C.POINTER(vp)],
        "snapb200_set_spmm_mode": [vp, i32],
        "snapb200_set_block": [vp, i32],
        "snapb200_dense_selftest": [vp, i64, i32, i32, C.POINTER(dbl)],
        "snapb200_ortho_selftest": [vp, i64, i32, i32, C.POINTER(dbl)],
        "snapb200_delta_selftest_host": [vp, i32, i64, C.POINTER(i64)],
        "snapb200_knn": [vp, i64, i32, vp, i32, i64, i64, i32, vp, vp],
        "snapb200_knn_limits": [C.POINTER(i32), C.POINTER(i32)],
        "snapb200_sym_eig": [i32, vp, vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = i32
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().snapb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libsnapb200: {msg}")


def ptr(a):
    """Raw data pointer of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())   # torch tensor


_VALUE_KINDS = {
    np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.uint32): 3,
    np.dtype(np.int32): 4, np.dtype(np.int64): 5, np.dtype(np.uint64): 6,
    np.dtype(np.uint8): 7, np.dtype(np.bool_): 7, np.dtype(np.int8): 8,
    np.dtype(np.uint16): 9, np.dtype(np.int16): 10,
}


def value_kind(dtype) -> int | None:
    return _VALUE_KINDS.get(np.dtype(dtype))
