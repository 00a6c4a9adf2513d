"""One host batch sharded over several GPUs of the box in a single process.

`ReplicatedInterpolator` builds the same interpolator on every requested device (grid replicated) and evaluates
host batches through `nint_interpolate_batch_host_multi`: contiguous index ranges, one host thread and one PCIe
link per GPU, no collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .error import InterpolateError, check
from .interpolator import _check_host_out


class ReplicatedInterpolator:
    def __init__(self, factory, devices):
        """factory(device) -> an Interp1D/2D/3D/ND built on that device; devices: iterable of CUDA ordinals."""
        self.replicas = [factory(int(d)) for d in devices]
        if not self.replicas:
            raise ValueError("need at least one device")
        self.dtype = self.replicas[0].dtype

    def ndim(self) -> int:
        return self.replicas[0].ndim()

    def close(self):
        for r in self.replicas:
            r.close()

    def interpolate_batch_host(self, points: np.ndarray, out=None, return_status: bool = False):
        """points: (n, ndim) array of host points (pinned memory reaches PCIe speed on every link)."""
        nd = self.ndim()
        if points.ndim != 2 or points.shape[1] != nd:
            raise InterpolateError("PointLength", f"supplied point slice should have length {nd} for {nd}-D interpolation")
        pts = np.ascontiguousarray(points, dtype=self.dtype)
        n = pts.shape[0]
        if out is None:
            out = np.empty(n, dtype=self.dtype)
        _check_host_out(out, self.dtype, n)
        status = np.empty(n, dtype=np.uint8) if return_status else None
        info = L.BatchInfo()
        hs = (C.c_void_p * len(self.replicas))(*[r._h for r in self.replicas])
        rc = L.lib().nint_interpolate_batch_host_multi(
            hs, len(self.replicas), L.LAYOUT_AOS, C.c_void_p(pts.ctypes.data), n, C.c_void_p(out.ctypes.data),
            C.c_void_p(status.ctypes.data) if status is not None else None, C.byref(info))
        if return_status and rc in (L.E_EXTRAPOLATE, L.E_PANIC):
            rc = L.OK
        check(rc)
        if return_status:
            return out, status, dict(n_error=info.n_error, n_panic=info.n_panic, first_error=info.first_error,
                                     first_panic=info.first_panic)
        return out
