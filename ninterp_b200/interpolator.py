"""Host-side mirror of the reference's interpolator API over the C ABI.

Same names, argument meaning and error behaviour as the Rust crate (paths relative to the
reference's src/): `Interp1D/2D/3D/ND::new` (interpolator/{one,two,three,n}/mod.rs), the
`Interpolator` trait (`ndim`, `validate`, `interpolate`, `set_extrapolate`; interpolator/mod.rs:25-34),
`set_strategy`, plus the batch entry points this engine adds.  All arithmetic runs in the CUDA
kernels behind include/ninterp_cuda.h; nothing here computes an interpolated value.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib as L
from . import strategy as _strategy
from .error import EngineError, InterpolateError, check
from .extrapolate import Extrapolate

_NP = {L.F32: np.float32, L.F64: np.float64}


def _dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return L.F64
    if dt == np.float32:
        return L.F32
    raise EngineError(f"unsupported element type {dt}: the engine evaluates f32 and f64")


def _default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class InterpData:
    """`InterpData<D, N>` / `InterpDataND<D>` (interpolator/data.rs:31-44, n/mod.rs:26-35)."""

    def __init__(self, grid, values):
        self.grid = grid
        self.values = values


class _InterpBase:
    """Shared implementation of the `Interpolator<T>` trait for the GPU-backed interpolators."""

    _kind = L.KIND_FIXED
    _N: int | None = None

    def __init__(self, grid: Sequence, values, strategy=_strategy.Linear, extrapolate: Extrapolate = Extrapolate.Error,
                 dtype=None, device: int | None = None):
        values = np.asarray(values)
        if dtype is None:
            dtype = values.dtype if values.dtype in (np.float32, np.float64) else np.float64
        self._code = _dtype_code(dtype)
        self.dtype = np.dtype(_NP[self._code])
        grid = [np.ascontiguousarray(g, dtype=self.dtype).reshape(-1) for g in grid]
        values = np.asarray(values, dtype=self.dtype)  # keeps strides: views are uploaded through value_strides
        self.data = InterpData(grid, values)
        self.strategy = strategy
        self.extrapolate = extrapolate
        self.device = _default_device() if device is None else int(device)
        self._h = C.c_void_p()
        ngrid = len(grid)
        axis_len = (C.c_size_t * max(ngrid, 1))(*[len(g) for g in grid])
        axes = (C.c_void_p * max(ngrid, 1))(*[g.ctypes.data if len(g) else None for g in grid])
        nvd = values.ndim
        shape = (C.c_size_t * max(nvd, 1))(*values.shape)
        item = self.dtype.itemsize
        strides = (C.c_ssize_t * max(nvd, 1))(*[s // item for s in values.strides])
        fill = self._fill_arg(extrapolate)
        rc = L.lib().nint_create(self.device, self._kind, self._code, ngrid, axis_len, axes, nvd, shape,
                                 C.c_void_p(values.ctypes.data) if values.size else None, strides,
                                 strategy.code, extrapolate.code, fill, C.byref(self._h))
        check(rc)

    # -- plumbing ------------------------------------------------------------------------------
    def _fill_arg(self, extrapolate):
        if extrapolate.code != L.EXTRAPOLATE_FILL:
            return None
        self._fill_buf = np.array([extrapolate.value], dtype=self.dtype)
        return C.c_void_p(self._fill_buf.ctypes.data)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.lib().nint_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Interpolator<T> -----------------------------------------------------------------------
    def ndim(self) -> int:
        """`Interpolator::ndim`."""
        return int(L.lib().nint_ndim(self._h))

    def validate(self) -> None:
        """`Interpolator::validate` (three/mod.rs:186-191)."""
        check(L.lib().nint_validate(self._h))

    def interpolate(self, point: Sequence[float]) -> float:
        """`Interpolator::interpolate(&self, point: &[T]) -> Result<T, InterpolateError>`."""
        pt = np.ascontiguousarray(point, dtype=self.dtype).reshape(-1)
        out = np.empty(1, dtype=self.dtype)
        rc = L.lib().nint_interpolate(self._h, C.c_void_p(pt.ctypes.data) if pt.size else None, pt.size,
                                      C.c_void_p(out.ctypes.data))
        check(rc)
        return out[0]

    def set_extrapolate(self, extrapolate: Extrapolate) -> None:
        """`Interpolator::set_extrapolate` (three/mod.rs:240-244)."""
        check(L.lib().nint_set_extrapolate(self._h, extrapolate.code, self._fill_arg(extrapolate)))
        self.extrapolate = extrapolate

    def set_strategy(self, strategy) -> None:
        """`InterpXD::set_strategy` for the strategy enums (three/mod.rs:259-271)."""
        rc = L.lib().nint_set_strategy(self._h, strategy.code)
        # the reference assigns before checking (three/mod.rs:259-271), so the new strategy stays in place when the
        # check fails; a strategy the handle never took (not built in for this interpolator) leaves both unchanged
        if rc not in (L.E_UNSUPPORTED, L.E_INVALID_ARGUMENT):
            self.strategy = strategy
        check(rc)

    def last_query_order(self) -> int:
        """Diagnostics: route of the most recent batch (0 no order probe ran, 1 coherent: single pass on the
        cell-packed table, 2 incoherent: slab lists, slab passes or binned tiles).  Synchronises the device."""
        return int(L.lib().nint_last_query_order(self._h))

    def last_incoherent_route(self) -> int:
        """Diagnostics: what the most recent batch enqueued for an incoherent verdict (0 nothing, 1 slab passes on
        the plain table, 2 binned tiles staged in shared memory by TMA, 3 slab lists on the cell-packed table)."""
        return int(L.lib().nint_last_incoherent_route(self._h))

    # -- batch entry points (new) ----------------------------------------------------------------
    def interpolate_batch(self, coords, out=None, status=None, info=None, stream=None):
        """Evaluate points resident on this interpolator's GPU (torch CUDA tensors).

        coords: a list of `ndim` 1-D tensors (struct of arrays) or one (n, ndim) tensor (array of
        points).  Returns `out` (and leaves per-point status bytes in `status` if given).
        Asynchronous on `stream` (default: torch's current stream).  `info`, if given, is a
        4-element int64/uint64 CUDA tensor receiving (n_error, n_panic, first_error, first_panic).
        """
        import torch

        tdt = torch.float64 if self._code == L.F64 else torch.float32
        nd = self.ndim()
        aos = isinstance(coords, torch.Tensor)
        if aos:
            if coords.dim() != 2 or coords.shape[1] != nd:
                raise InterpolateError(
                    "PointLength", f"supplied point slice should have length {nd} for {nd}-D interpolation")
            tensors = [coords]
            n = coords.shape[0]
        else:
            tensors = list(coords)
            if len(tensors) != nd:
                raise InterpolateError(
                    "PointLength", f"supplied point slice should have length {nd} for {nd}-D interpolation")
            n = tensors[0].numel() if tensors else (out.numel() if out is not None else 0)
        for t in tensors:
            if not t.is_cuda or t.device.index != self.device or t.dtype != tdt or not t.is_contiguous():
                raise EngineError(f"coordinates must be contiguous {tdt} tensors on cuda:{self.device}")
            if not aos and t.numel() != n:
                raise EngineError("coordinate columns differ in length")
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty(n, dtype=tdt, device=dev)
        # caller-supplied result buffers are handed to the C ABI as raw pointers: check them here
        for name, t, want_dt, want_n in (("out", out, (tdt,), n), ("status", status, (torch.uint8,), n),
                                         ("info", info, (torch.int64, torch.uint64), 4)):
            if t is None:
                continue
            if (not isinstance(t, torch.Tensor) or not t.is_cuda or t.device.index != self.device or t.dtype not in want_dt
                    or not t.is_contiguous() or t.numel() < want_n):
                raise EngineError(f"`{name}` must be a contiguous {want_dt[0]} tensor of at least {want_n} elements on cuda:{self.device}")
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        st_ptr = C.c_void_p(status.data_ptr()) if status is not None else None
        info_ptr = C.c_void_p(info.data_ptr()) if info is not None else None
        lib = L.lib()
        if aos:
            rc = lib.nint_interpolate_batch_aos(self._h, C.c_void_p(coords.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                st_ptr, info_ptr, C.c_void_p(stream))
        else:
            cols = (C.c_void_p * max(nd, 1))(*[t.data_ptr() for t in tensors])
            rc = lib.nint_interpolate_batch(self._h, cols, n, C.c_void_p(out.data_ptr()), st_ptr, info_ptr,
                                            C.c_void_p(stream))
        check(rc)
        return out

    def interpolate_batch_host(self, points, out=None, return_status: bool = False):
        """Evaluate host points; the C library streams them through the GPU in overlapped chunks.

        points: (n, ndim) array (a Rust `&[[T; N]]`) or a list of `ndim` columns.  With
        `return_status=False` this behaves like mapping `interpolate` over the batch and collecting
        a `Result<Vec<T>, InterpolateError>`: the first failing point raises.  With
        `return_status=True` it returns (values, status bytes, info dict) and never raises for
        per-point failures.
        """
        nd = self.ndim()
        lib = L.lib()
        keep = []
        if isinstance(points, np.ndarray) and points.ndim == 2:
            layout = L.LAYOUT_AOS
            if points.shape[1] != nd:
                raise InterpolateError(
                    "PointLength", f"supplied point slice should have length {nd} for {nd}-D interpolation")
            pts = np.ascontiguousarray(points, dtype=self.dtype)
            keep.append(pts)
            n = pts.shape[0]
            data = C.c_void_p(pts.ctypes.data)
        else:
            layout = L.LAYOUT_SOA
            cols = [np.ascontiguousarray(c, dtype=self.dtype).reshape(-1) for c in points]
            if len(cols) != nd:
                raise InterpolateError(
                    "PointLength", f"supplied point slice should have length {nd} for {nd}-D interpolation")
            keep.extend(cols)
            n = len(cols[0]) if cols else 0
            arr = (C.c_void_p * max(nd, 1))(*[c.ctypes.data for c in cols])
            keep.append(arr)
            data = C.cast(arr, C.c_void_p)
        if out is None:
            out = np.empty(n, dtype=self.dtype)
        _check_host_out(out, self.dtype, n)
        status = np.empty(n, dtype=np.uint8) if return_status else None
        info = L.BatchInfo()
        rc = lib.nint_interpolate_batch_host(self._h, layout, data, n, C.c_void_p(out.ctypes.data),
                                             C.c_void_p(status.ctypes.data) if status is not None else None,
                                             C.byref(info))
        if return_status and rc in (L.E_EXTRAPOLATE, L.E_PANIC):
            rc = L.OK
        check(rc)
        if return_status:
            return out, status, dict(n_error=info.n_error, n_panic=info.n_panic, first_error=info.first_error,
                                     first_panic=info.first_panic)
        return out


def _check_host_out(out, dtype, n):
    """A caller-supplied host result buffer is written through a raw pointer: dtype, layout and length must fit."""
    if (not isinstance(out, np.ndarray) or out.dtype != np.dtype(dtype) or not out.flags.c_contiguous
            or not out.flags.writeable or out.size < n):
        raise EngineError(f"`out` must be a writeable C-contiguous {np.dtype(dtype)} array of at least {n} elements")


class Interp1D(_InterpBase):
    """`Interp1D::new(x, f_x, strategy, extrapolate)` (interpolator/one/mod.rs:110-124).

    Strategies: Linear, Nearest, LeftNearest, RightNearest; `Extrapolate.Enable` needs Linear.
    """

    _N = 1

    def __init__(self, x, f_x, strategy=_strategy.Linear, extrapolate=Extrapolate.Error, **kw):
        super().__init__([x], f_x, strategy, extrapolate, **kw)


class Interp2D(_InterpBase):
    """`Interp2D::new(x, y, f_xy, strategy, extrapolate)` (interpolator/two/mod.rs:118-133)."""

    _N = 2

    def __init__(self, x, y, f_xy, strategy=_strategy.Linear, extrapolate=Extrapolate.Error, **kw):
        super().__init__([x, y], f_xy, strategy, extrapolate, **kw)


class Interp3D(_InterpBase):
    """`Interp3D::new(x, y, z, f_xyz, strategy, extrapolate)` (interpolator/three/mod.rs:129-145)."""

    _N = 3

    def __init__(self, x, y, z, f_xyz, strategy=_strategy.Linear, extrapolate=Extrapolate.Error, **kw):
        super().__init__([x, y, z], f_xyz, strategy, extrapolate, **kw)


class InterpND(_InterpBase):
    """`InterpND::new(grid, values, strategy, extrapolate)` (interpolator/n/mod.rs:225-239)."""

    _kind = L.KIND_ND

    def __init__(self, grid, values, strategy=_strategy.Linear, extrapolate=Extrapolate.Error, **kw):
        super().__init__(list(grid), values, strategy, extrapolate, **kw)


class Interp0D:
    """`Interp0D<T>(pub T)` (interpolator/zero/mod.rs): a constant; stays on the host."""

    def __init__(self, value):
        self.value = value

    def ndim(self) -> int:
        return 0

    def validate(self) -> None:
        return None

    def interpolate(self, point=()):
        if len(point) != 0:  # zero/mod.rs:52-54
            raise InterpolateError("PointLength", "supplied point slice should have length 0 for 0-D interpolation")
        return self.value

    def set_extrapolate(self, extrapolate) -> None:
        return None


class InterpolatorEnum:
    """Constructors of `InterpolatorEnum` (interpolator/enums.rs:117-191): pure dispatch."""

    new_0d = staticmethod(Interp0D)
    new_1d = staticmethod(Interp1D)
    new_2d = staticmethod(Interp2D)
    new_3d = staticmethod(Interp3D)
    new_nd = staticmethod(InterpND)
