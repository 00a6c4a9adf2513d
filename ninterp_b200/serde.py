"""Loader for the reference's serde JSON form: host-side parsing only.

The crate serialises `Interp{1,2,3}D` / `InterpND` as
    {"data": {"grid": [<array>, ...], "values": <array>}, "strategy": "Linear", "extrapolate": "Error" | {"Fill": x}}
with ndarray's array form {"v": 1, "dim": [..], "data": [flat, C order]} (ndarray's serde feature), the unit-struct
strategies as their names (src/strategy/mod.rs:10-14) and `Extrapolate` as an externally tagged enum
(src/interpolator/mod.rs:57-72, default Error when absent: one/mod.rs:59-60).  `Interp0D` is a bare number
(zero/mod.rs:11).  The hard-coded and N-D encodings of the same interpolator are identical (enums.rs:362-373), so
`kind` picks which one to build; by default grids of 1-3 axes become Interp1D/2D/3D like the untagged
`InterpolatorEnum` does.
"""
from __future__ import annotations

import json

import numpy as np

from . import strategy as S
from .extrapolate import Extrapolate
from .interpolator import Interp0D, Interp1D, Interp2D, Interp3D, InterpND

_STRATEGIES = {"Linear": S.Linear, "Nearest": S.Nearest, "LeftNearest": S.LeftNearest, "RightNearest": S.RightNearest}


def _array(obj, dtype):
    if isinstance(obj, dict):
        if obj.get("v", 1) != 1:
            raise ValueError(f"unsupported ndarray serde version {obj.get('v')}")
        return np.asarray(obj["data"], dtype=dtype).reshape(obj["dim"])
    return np.asarray(obj, dtype=dtype)


def _extrapolate(obj):
    if obj is None or obj == "Error":
        return Extrapolate.Error
    if isinstance(obj, str):
        return {"Enable": Extrapolate.Enable, "Clamp": Extrapolate.Clamp, "Wrap": Extrapolate.Wrap}[obj]
    if isinstance(obj, dict) and list(obj) == ["Fill"]:
        v = obj["Fill"]
        return Extrapolate.Fill(float("nan") if v is None else v)  # serde_json writes NaN as null
    raise ValueError(f"unknown Extrapolate encoding {obj!r}")


def from_json(text, kind: str = "auto", dtype=np.float64, device=None):
    """Build a GPU-backed interpolator from the reference's JSON.  kind: "auto" | "fixed" | "nd"."""
    doc = json.loads(text) if isinstance(text, (str, bytes)) else text
    if isinstance(doc, (int, float)):
        return Interp0D(doc)
    grid = [_array(g, dtype).reshape(-1) for g in doc["data"]["grid"]]
    values = _array(doc["data"]["values"], dtype)
    strat = doc.get("strategy", "Linear")
    if isinstance(strat, dict):  # Strategy*Enum is externally tagged around a unit struct
        strat = next(iter(strat))
    strategy = _STRATEGIES[strat]
    extrap = _extrapolate(doc.get("extrapolate"))
    kw = dict(dtype=dtype)
    if device is not None:
        kw["device"] = device
    if kind == "nd" or (kind == "auto" and not 1 <= len(grid) <= 3) or (kind == "auto" and values.ndim != len(grid)):
        return InterpND(grid, values, strategy, extrap, **kw)
    return [Interp1D, Interp2D, Interp3D][len(grid) - 1](*grid, values, strategy, extrap, **kw)


def to_json(interp) -> str:
    """The same form back (host copies of grid and values; nothing is read from the device)."""
    if isinstance(interp, Interp0D):
        return json.dumps(interp.value)

    def arr(a):
        a = np.asarray(a)
        return {"v": 1, "dim": list(a.shape), "data": a.reshape(-1).tolist()}

    e = interp.extrapolate
    # serde_json has no NaN: a NaN fill value is written as null
    ext = {"Fill": None if e.value != e.value else e.value} if repr(e).startswith("Fill") else repr(e)
    return json.dumps({"data": {"grid": [arr(g) for g in interp.data.grid], "values": arr(interp.data.values)},
                       "strategy": repr(interp.strategy), "extrapolate": ext})
