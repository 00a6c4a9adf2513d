"""ctypes loader for libninterp_cuda.so (the C ABI in include/ninterp_cuda.h).

There is no CPU fallback: if the shared library is missing this raises, and every compute entry
point of the library fails with NINT_E_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
# NINT_LIB_VARIANT=checked selects the bounds-checked build (see build.py)
_VARIANT = __import__("os").environ.get("NINT_LIB_VARIANT", "")
SO = PKG / ("libninterp_cuda" + ("_" + _VARIANT if _VARIANT else "") + ".so")

# enums of include/ninterp_cuda.h
F32, F64 = 0, 1
KIND_FIXED, KIND_ND = 0, 1
LINEAR, NEAREST, LEFT_NEAREST, RIGHT_NEAREST = 0, 1, 2, 3
EXTRAPOLATE_ENABLE, EXTRAPOLATE_FILL, EXTRAPOLATE_CLAMP, EXTRAPOLATE_WRAP, EXTRAPOLATE_ERROR = 0, 1, 2, 3, 4
LAYOUT_SOA, LAYOUT_AOS = 0, 1
OK = 0
E_EXTRAPOLATE_SELECTION, E_EMPTY_GRID, E_MONOTONICITY, E_INCOMPATIBLE_SHAPES, E_OTHER = -1, -2, -3, -4, -5
E_POINT_LENGTH, E_EXTRAPOLATE, E_PANIC, E_CUDA, E_INVALID_ARGUMENT, E_UNSUPPORTED = -6, -7, -8, -9, -10, -11
STATUS_OK, STATUS_EXTRAPOLATE_ERROR, STATUS_PANIC = 0, 1, 0xFF
MAX_NDIM = 8

# every symbol include/ninterp_cuda.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "nint_create", "nint_destroy", "nint_ndim", "nint_validate", "nint_set_extrapolate", "nint_set_strategy",
    "nint_interpolate_batch", "nint_interpolate_batch_aos", "nint_interpolate_batch_host", "nint_interpolate_batch_host_multi", "nint_interpolate",
    "nint_host_alloc", "nint_host_free", "nint_last_error_message", "nint_last_error_dim", "nint_version",
    "nint_launch_count", "nint_selftest_divide", "nint_last_query_order", "nint_last_incoherent_route",
]


class BatchInfo(C.Structure):
    _fields_ = [("n_error", C.c_uint64), ("n_panic", C.c_uint64), ("first_error", C.c_uint64),
                ("first_panic", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not SO.exists():
        raise RuntimeError(
            f"{SO} is missing: build it with `python -m ninterp_b200.build` (nvcc, sm_100a). "
            "ninterp_b200 has no CPU fallback.")
    L = C.CDLL(str(SO))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.nint_create.argtypes = [i32, i32, i32, i32, C.POINTER(sz), C.POINTER(vp), i32, C.POINTER(sz), vp,
                              C.POINTER(C.c_ssize_t), i32, i32, vp, C.POINTER(vp)]
    L.nint_create.restype = i32
    L.nint_destroy.argtypes = [vp]
    L.nint_destroy.restype = None
    L.nint_ndim.argtypes = [vp]
    L.nint_validate.argtypes = [vp]
    L.nint_set_extrapolate.argtypes = [vp, i32, vp]
    L.nint_set_strategy.argtypes = [vp, i32]
    L.nint_interpolate_batch.argtypes = [vp, C.POINTER(vp), sz, vp, vp, vp, vp]
    L.nint_interpolate_batch_aos.argtypes = [vp, vp, sz, vp, vp, vp, vp]
    L.nint_interpolate_batch_host.argtypes = [vp, i32, vp, sz, vp, vp, C.POINTER(BatchInfo)]
    L.nint_interpolate.argtypes = [vp, vp, sz, vp]
    L.nint_interpolate_batch_host_multi.argtypes = [C.POINTER(vp), i32, i32, vp, sz, vp, vp, C.POINTER(BatchInfo)]
    L.nint_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.nint_host_free.argtypes = [vp]
    L.nint_last_error_message.restype = C.c_char_p
    L.nint_last_error_dim.restype = i32
    L.nint_version.restype = C.c_char_p
    L.nint_launch_count.restype = C.c_uint64
    L.nint_last_query_order.argtypes = [vp]
    L.nint_last_incoherent_route.argtypes = [vp]
    L.nint_selftest_divide.argtypes = [i32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def last_error() -> tuple[str, int]:
    L = lib()
    return L.nint_last_error_message().decode(), int(L.nint_last_error_dim())
