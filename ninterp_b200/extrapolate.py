"""`ninterp::Extrapolate<T>` (src/interpolator/mod.rs:57-72)."""
from __future__ import annotations

from . import _lib as L


class Extrapolate:
    """Enable | Fill(value) | Clamp | Wrap | Error (the default)."""

    __slots__ = ("code", "value")
    _names = {L.EXTRAPOLATE_ENABLE: "Enable", L.EXTRAPOLATE_FILL: "Fill", L.EXTRAPOLATE_CLAMP: "Clamp",
              L.EXTRAPOLATE_WRAP: "Wrap", L.EXTRAPOLATE_ERROR: "Error"}

    def __init__(self, code: int, value: float | None = None):
        self.code, self.value = code, value

    @staticmethod
    def Fill(value: float) -> "Extrapolate":
        return Extrapolate(L.EXTRAPOLATE_FILL, float(value))

    def __repr__(self):
        return f"Fill({self.value})" if self.code == L.EXTRAPOLATE_FILL else self._names[self.code]

    def __eq__(self, other):
        if not isinstance(other, Extrapolate) or other.code != self.code:
            return False
        return self.code != L.EXTRAPOLATE_FILL or other.value == self.value  # NaN != NaN, as in Rust


Extrapolate.Enable = Extrapolate(L.EXTRAPOLATE_ENABLE)
Extrapolate.Clamp = Extrapolate(L.EXTRAPOLATE_CLAMP)
Extrapolate.Wrap = Extrapolate(L.EXTRAPOLATE_WRAP)
Extrapolate.Error = Extrapolate(L.EXTRAPOLATE_ERROR)
