#!/usr/bin/env python
"""Writes the serde JSON fixtures under tests/golden/.

The reference crate cannot be compiled in this image (no Rust toolchain), so these are not crate output: they are
the strings `serde_json::to_string` produces for the interpolators the reference's own serde tests build, derived
from the serde attributes in the source:
  * arrays: ndarray's serde form {"v":1,"dim":[...],"data":[flat C order]} (ndarray 0.17, feature "serde")
  * strategies: unit structs serialise as their name (src/strategy/mod.rs:10-14, test :49-66); the Strategy*Enum
    wrappers are #[serde(untagged)] (src/strategy/enums/one.rs:6-7) and serialise identically (one.rs:96-113)
  * Extrapolate: externally tagged enum (src/interpolator/mod.rs:57-72): "Error", "Clamp", ... or {"Fill": x};
    serde_json writes a NaN fill as null
  * InterpolatorEnum: #[serde(untagged)] (src/interpolator/enums.rs:66-67): the encoding of the variant it holds;
    enums.rs:362-373 asserts that Interp2D, InterpND and both enum wrappings of the same data give one string
Run from the repository root: python tests/golden/make_serde_fixtures.py
"""
import json
from pathlib import Path

HERE = Path(__file__).resolve().parent


def arr(data, dim):
    return {"v": 1, "dim": dim, "data": data}


def main():
    # src/interpolator/enums.rs:338-360 (test_serde): x, y, f_xy, Nearest, Error
    enums_2d = {"data": {"grid": [arr([0.05, 0.10, 0.15], [3]), arr([0.10, 0.20, 0.30], [3])],
                         "values": arr([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0], [3, 3])},
                "strategy": "Nearest", "extrapolate": "Error"}
    # src/interpolator/three/tests.rs:213-229 (test_serde)
    three = {"data": {"grid": [arr([0.0, 1.0], [2])] * 3, "values": arr([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0], [2, 2, 2])},
             "strategy": "Nearest", "extrapolate": "Error"}
    # src/interpolator/one/tests.rs:134-147 (Fill(NaN)): serde_json has no NaN, the fill value becomes null
    one_fill_nan = {"data": {"grid": [arr([0.0, 1.0, 2.0, 3.0, 4.0], [5])], "values": arr([0.2, 0.4, 0.6, 0.8, 1.0], [5])},
                    "strategy": "Linear", "extrapolate": {"Fill": None}}
    # a field the crate marks #[serde(default)] may be absent (one/mod.rs:62): extrapolate defaults to Error
    one_default = {"data": one_fill_nan["data"], "strategy": "LeftNearest"}
    for name, doc in (("serde_enums_rs_2d_nearest.json", enums_2d), ("serde_three_tests_nearest.json", three),
                      ("serde_one_fill_nan.json", one_fill_nan), ("serde_one_default_extrapolate.json", one_default)):
        (HERE / name).write_text(json.dumps(doc, separators=(",", ":")) + "\n")
    (HERE / "serde_interp0d.json").write_text("0.5\n")  # Interp0D<T>(pub T): a bare number (zero/mod.rs:11)


if __name__ == "__main__":
    main()
