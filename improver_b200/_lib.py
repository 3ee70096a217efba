"""ctypes binding of libimprover_b200.so (the C ABI in include/improver_b200.h).

There is deliberately no fallback: if the shared object is missing or a CUDA
device is not available the import / call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("IMPROVER_B200_LIB", os.path.join(_HERE, "lib", "libimprover_b200.so"))

OPS = {">": 0, "gt": 0, "GT": 0, ">=": 1, "ge": 1, "GE": 1,
       "<": 2, "lt": 2, "LT": 2, "<=": 3, "le": 3, "LE": 3, "==": 4, "eq": 4, "EQ": 4}

SUM_ONLY, WRAP_X, WEIGHTED, EXACT_SUMS, LAND_SEA = 1, 2, 4, 8, 16


class NbhThreshold(C.Structure):
    _fields_ = [("value", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
                ("op", C.c_int32), ("_pad", C.c_int32)]


class NbhDesc(C.Structure):
    _fields_ = [
        ("n_planes", C.c_int64), ("n_members", C.c_int32), ("ny", C.c_int32),
        ("nx", C.c_int32), ("cells", C.c_int32),
        ("plane_stride", C.c_int64), ("member_stride", C.c_int64),
        ("n_thresholds", C.c_int32), ("flags", C.c_int32),
        ("thresholds", C.POINTER(NbhThreshold)),
        ("mask2d", C.c_void_p), ("mask_nd", C.c_void_p),
        ("plane_cells", C.POINTER(C.c_int32)),
        ("y_offset", C.c_int32), ("ny_global", C.c_int32), ("halo_top", C.c_int32), ("halo_bottom", C.c_int32),
        ("ext_stats", C.c_void_p),
    ]


class NbhSliceStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("y0min", "y0max", "x0min", "x0max", "y1min", "y1max", "x1min", "x1max",
                                         "vmin_key", "vmax_key")]


class EngineError(RuntimeError):
    """Raised when the native library reports an error."""


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise ImportError(
            f"{LIBPATH} is missing: build it with `python -m improver_b200._build` "
            "(or __graft_entry__.build()); improver_b200 has no CPU fallback.")
    lib = C.CDLL(LIBPATH)
    lib.nbh_version.restype = C.c_int
    lib.nbh_last_error.restype = C.c_char_p
    lib.nbh_device_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.nbh_workspace_bytes.restype = C.c_size_t
    lib.nbh_workspace_bytes.argtypes = [C.POINTER(NbhDesc)]
    for name in ("nbh_square_f32", "nbh_circular_f32"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(NbhDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_size_t, C.c_void_p]
    lib.nbh_slice_stats_f32.restype = C.c_int
    lib.nbh_slice_stats_f32.argtypes = [C.POINTER(NbhDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nbh_halo_pack.restype = C.c_int
    lib.nbh_halo_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.nbh_halo_unpack.restype = C.c_int
    lib.nbh_halo_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p]
    lib.nbh_halo_exchange_workspace_bytes.restype = C.c_size_t
    lib.nbh_halo_exchange_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    lib.nbh_halo_exchange.restype = C.c_int
    lib.nbh_halo_exchange.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.nbh_collapse_mean_f32.restype = C.c_int
    lib.nbh_collapse_mean_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.nbh_boxsum_f64.restype = C.c_int
    lib.nbh_boxsum_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]
    lib.nbh_linear_map_f32.restype = C.c_int
    lib.nbh_linear_map_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p]
    lib.nbh_zone_collapse_f32.restype = C.c_int
    lib.nbh_zone_collapse_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                          C.c_void_p]
    lib.nbh_percentiles_f32.restype = C.c_int
    lib.nbh_percentiles_f32.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int32,
                                        C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.nbh_percentiles_workspace_bytes.restype = C.c_size_t
    lib.nbh_percentiles_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32]
    lib.nbh_threshold_f32.restype = C.c_int
    lib.nbh_threshold_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(NbhThreshold), C.c_int32, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p]
    lib.nbh_threshold_vicinity_f32.restype = C.c_int
    lib.nbh_threshold_vicinity_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(NbhThreshold), C.c_int32, C.POINTER(C.c_int32),
                                               C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p]
    lib.nbh_vicinity_f32.restype = C.c_int
    lib.nbh_vicinity_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                     C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.nbh_recursive_filter_f32.restype = C.c_int
    lib.nbh_recursive_filter_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_size_t, C.c_void_p]
    lib.nbh_recursive_filter_workspace_bytes.restype = C.c_size_t
    lib.nbh_recursive_filter_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.nbh_recurse_f32.restype = C.c_int
    lib.nbh_recurse_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p]
    lib.nbh_profile_enable.argtypes = [C.c_int]
    lib.nbh_profile_last_ms.argtypes = [C.POINTER(C.c_float)]
    lib.nbh_profile_last_kernel.restype = C.c_char_p
    lib.nbh_profile_last_kernel.argtypes = []
    if lib.nbh_version() != 2:
        raise ImportError("libimprover_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().nbh_last_error().decode("utf-8", "replace")
        raise EngineError(msg or f"native call failed with code {rc}")


EXPORTED = ["nbh_version", "nbh_last_error", "nbh_device_info", "nbh_workspace_bytes", "nbh_slice_stats_f32",
            "nbh_collapse_mean_f32", "nbh_zone_collapse_f32", "nbh_linear_map_f32", "nbh_boxsum_f64", "nbh_halo_pack", "nbh_halo_unpack", "nbh_halo_exchange", "nbh_halo_exchange_workspace_bytes",
            "nbh_square_f32", "nbh_circular_f32", "nbh_percentiles_f32",
            "nbh_percentiles_workspace_bytes", "nbh_threshold_f32", "nbh_threshold_vicinity_f32", "nbh_vicinity_f32",
            "nbh_recursive_filter_f32", "nbh_recursive_filter_workspace_bytes", "nbh_recurse_f32",
            "nbh_profile_enable", "nbh_profile_last_ms", "nbh_profile_last_kernel"]
