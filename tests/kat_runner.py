"""Shared KAT driver: runs tests/kat_vectors.py against any evaluator.

`evaluate(case, points) -> (values ndarray, status ndarray)`; TEST INFRASTRUCTURE ONLY.
"""
import math

import numpy as np

import kat_vectors as K


def run_case(case, evaluate):
    axes = case["axes"]
    values = np.asarray(case["values"], dtype=np.float64)
    for chk in case["checks"]:
        kind = chk[0]
        where = f"{case['name']} ({case['ref']}) {chk}"
        if kind == "nodes":
            pts, idx = K.node_points(axes)
            got, st = evaluate(case, pts)
            want = values[tuple(idx.T)]
            assert (st == 0).all(), where
            assert got.tobytes() == want.tobytes(), where
        elif kind == "exact":
            got, st = evaluate(case, [chk[1]])
            assert st[0] == 0, where
            assert np.float64(got[0]).tobytes() == np.float64(chk[2]).tobytes(), f"{where}: got {got[0]!r}"
        elif kind == "approx":
            got, st = evaluate(case, [chk[1]])
            assert st[0] == 0, where
            assert abs(float(got[0]) - chk[2]) <= 1e-6, f"{where}: got {got[0]!r}"
        elif kind == "nan":
            got, st = evaluate(case, [chk[1]])
            assert st[0] == 0 and math.isnan(got[0]), where
        elif kind == "error":
            got, st = evaluate(case, [chk[1]])
            assert st[0] == 1, where
        elif kind == "same":
            got, st = evaluate(case, [chk[1], chk[2]])
            assert (st == 0).all(), where
            assert got[0].tobytes() == got[1].tobytes(), f"{where}: {got[0]!r} vs {got[1]!r}"
        else:
            raise AssertionError(kind)
