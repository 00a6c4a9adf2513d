"""`ninterp::strategy` unit structs (src/strategy/mod.rs:14,25,33,41) as Python singletons."""
from __future__ import annotations

from . import _lib as L


class _Strategy:
    code: int
    name: str

    def __repr__(self):
        return self.name

    def __eq__(self, other):
        return isinstance(other, _Strategy) and other.code == self.code

    def __hash__(self):
        return hash(self.code)

    def allow_extrapolate(self) -> bool:
        """`StrategyXD::allow_extrapolate`: true for Linear only (one/strategies.rs:32,61,84,107)."""
        return self.code == L.LINEAR


def _make(name, code):
    return type(name, (_Strategy,), {"code": code, "name": name})()


Linear = _make("Linear", L.LINEAR)
Nearest = _make("Nearest", L.NEAREST)
LeftNearest = _make("LeftNearest", L.LEFT_NEAREST)
RightNearest = _make("RightNearest", L.RIGHT_NEAREST)
