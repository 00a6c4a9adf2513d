"""ctypes binding of the CPU oracle (oracle/_build/libninterp_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package `ninterp_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
SO = ORACLE_DIR / "_build" / "libninterp_oracle.so"

LINEAR, NEAREST, LEFT_NEAREST, RIGHT_NEAREST = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4
KIND_FIXED, KIND_ND = 0, 1
ST_OK, ST_EXTRAPOLATE_ERROR, ST_PANIC = 0, 1, 0xFF

_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with its own Makefile (g++, -ffp-contract=off)."""
    srcs = [ORACLE_DIR / "oracle_capi.cpp", ORACLE_DIR / "ninterp_oracle.hpp", ORACLE_DIR / "Makefile"]
    stale = (not SO.exists()) or any(s.stat().st_mtime > SO.stat().st_mtime for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.run(["make", "-C", str(ORACLE_DIR)] + (["-B"] if force else []), check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(SO))
        _lib.nio_wrap_f64.restype = C.c_double
        _lib.nio_wrap_f64.argtypes = [C.c_double] * 3
        _lib.nio_wrap_f32.restype = C.c_float
        _lib.nio_wrap_f32.argtypes = [C.c_float] * 3
        _lib.nio_find_nearest_index_f64.restype = C.c_longlong
        _lib.nio_find_nearest_index_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
        _lib.nio_find_nearest_index_f32.restype = C.c_longlong
        _lib.nio_find_nearest_index_f32.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        _lib.nio_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().nio_max_threads())


def _ptr_array(arrs):
    return (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])


def _size_array(vals):
    return (C.c_size_t * max(len(vals), 1))(*[int(v) for v in vals])


class Oracle:
    """One interpolator (grid + values + strategy + extrapolate) evaluated by the CPU oracle."""

    def __init__(self, axes, values, strategy, extrap, fill=0.0, kind=KIND_FIXED, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.axes = [np.ascontiguousarray(a, dtype=self.dtype) for a in axes]
        self.values = np.ascontiguousarray(values, dtype=self.dtype)
        self.strategy, self.extrap, self.fill, self.kind = strategy, extrap, fill, kind
        self.ndim = len(self.axes)
        self.shape = list(self.values.shape)

    @property
    def _sfx(self):
        return "f64" if self.dtype == np.float64 else "f32"

    def validate(self):
        """(code, dim) of `InterpXD::new`'s validation: data.validate() then check_extrapolate()."""
        dim = C.c_int(-1)
        fn = getattr(lib(), f"nio_validate_{self._sfx}")
        rc = fn(self.kind, self.ndim, _size_array([len(a) for a in self.axes]), len(self.shape),
                _size_array(self.shape), _ptr_array(self.axes), self.strategy, self.extrap, C.byref(dim))
        return rc, dim.value

    def eval(self, coords, nthreads=1, want_status=True):
        coords = [np.ascontiguousarray(c, dtype=self.dtype) for c in coords]
        n = len(coords[0]) if coords else 0
        out = np.empty(n, dtype=self.dtype)
        status = np.empty(n, dtype=np.uint8)
        fn = getattr(lib(), f"nio_eval_{self._sfx}")
        fill = C.c_double(self.fill) if self.dtype == np.float64 else C.c_float(self.fill)
        fn(self.kind, self.ndim, _size_array([len(a) for a in self.axes]), _size_array(self.shape),
           _ptr_array(self.axes), C.c_void_p(self.values.ctypes.data), self.strategy, self.extrap, fill,
           _ptr_array(coords), C.c_size_t(n), C.c_void_p(out.ctypes.data),
           C.c_void_p(status.ctypes.data), int(nthreads))
        return (out, status) if want_status else out

    def eval_points(self, points, nthreads=1):
        """points: (n, ndim) array-like (AoS)."""
        pts = np.asarray(points, dtype=self.dtype).reshape(len(points), self.ndim)
        coords = [np.ascontiguousarray(pts[:, k]) for k in range(self.ndim)]
        return self.eval(coords, nthreads)

    def one(self, point):
        out, st = self.eval_points([list(point)])
        return out[0], int(st[0])


def wrap(x, lo, hi, dtype=np.float64):
    if np.dtype(dtype) == np.float64:
        return float(lib().nio_wrap_f64(x, lo, hi))
    return float(lib().nio_wrap_f32(x, lo, hi))


def find_nearest_index(arr, target, dtype=np.float64):
    a = np.ascontiguousarray(arr, dtype=dtype)
    if np.dtype(dtype) == np.float64:
        return int(lib().nio_find_nearest_index_f64(a.ctypes.data, len(a), float(target)))
    return int(lib().nio_find_nearest_index_f32(a.ctypes.data, len(a), float(target)))


def check_extrapolate(axis_lens, strategy, current, requested):
    dim = C.c_int(-1)
    rc = lib().nio_check_extrapolate(len(axis_lens), _size_array(axis_lens), strategy, current, requested,
                                     C.byref(dim))
    return rc, dim.value
