"""Error types mirroring src/error.rs of the reference crate."""
from __future__ import annotations

from . import _lib as L


class ValidateError(Exception):
    """ninterp::error::ValidateError (src/error.rs:9-27); `.variant` names the enum variant."""

    def __init__(self, variant: str, message: str, dim: int | None = None):
        super().__init__(message)
        self.variant, self.dim = variant, dim


class InterpolateError(Exception):
    """ninterp::error::InterpolateError (src/error.rs:38-45)."""

    def __init__(self, variant: str, message: str):
        super().__init__(message)
        self.variant = variant


class ReferencePanic(Exception):
    """The reference crate panics on this input (index out of bounds); the engine reports it instead."""


class EngineError(RuntimeError):
    """CUDA failure, bad argument or unsupported configuration (no reference counterpart)."""


_VALIDATE = {
    L.E_EXTRAPOLATE_SELECTION: "ExtrapolateSelection",
    L.E_EMPTY_GRID: "EmptyGrid",
    L.E_MONOTONICITY: "Monotonicity",
    L.E_INCOMPATIBLE_SHAPES: "IncompatibleShapes",
    L.E_OTHER: "Other",
}


def check(rc: int) -> None:
    if rc == L.OK:
        return
    msg, dim = L.last_error()
    if rc in _VALIDATE:
        raise ValidateError(_VALIDATE[rc], msg, dim if dim >= 0 else None)
    if rc == L.E_POINT_LENGTH:
        raise InterpolateError("PointLength", msg)
    if rc == L.E_EXTRAPOLATE:
        raise InterpolateError("ExtrapolateError", msg)
    if rc == L.E_PANIC:
        raise ReferencePanic(msg)
    raise EngineError(f"ninterp_cuda error {rc}: {msg}")
