"""Cross-checks the two independently written CPU models (oracle/ C++ vs tests/pyref.py).

The C++ oracle follows the Rust line by line; pyref uses the closed forms the CUDA kernels use.
They must agree bit-for-bit, statuses included, on the reference KATs and on adversarial inputs.
"""
import itertools
import math

import numpy as np
import pytest

import kat_vectors as K
import oracle_binding as O
import pyref as P
from kat_runner import run_case


def _pyref_eval(case, points):
    fill = case["fill"] if case["fill"] is not None else 0.0
    vals = np.asarray(case["values"], dtype=np.float64)
    out = np.empty(len(points))
    st = np.empty(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        out[i], st[i] = P.interpolate(case["kind"], case["axes"], vals, case["strategy"], case["extrap"], fill, p)
    return out, st


@pytest.mark.parametrize("case", K.KATS, ids=[c["name"] for c in K.KATS])
def test_reference_kat(case):
    run_case(case, _pyref_eval)


def _special_points(axes, rng, n_random, dtype):
    """Adversarial coordinates per axis: nodes, neighbours of nodes, far outside, +-0, inf, NaN, denormals."""
    per_axis = []
    for g in axes:
        g = np.asarray(g, dtype=dtype)
        span = float(g[-1] - g[0]) if len(g) > 1 else 1.0
        c = list(g) + list(np.nextafter(g, dtype(np.inf))) + list(np.nextafter(g, dtype(-np.inf)))
        c += [g[0] - dtype(0.3 * span), g[-1] + dtype(0.7 * span), g[0] - dtype(5 * span), g[-1] + dtype(5 * span)]
        c += [dtype(0.0), dtype(-0.0), dtype(np.inf), dtype(-np.inf), dtype(np.nan),
              np.finfo(dtype).tiny, -np.finfo(dtype).tiny, np.finfo(dtype).max, dtype(5e-324)]
        c += list(rng.uniform(float(g[0]) - 0.5 * span, float(g[-1]) + 0.5 * span, n_random).astype(dtype))
        per_axis.append(np.array(c, dtype=dtype))
    return per_axis


def _axes(rng, lens, dtype, dup=False):
    axes = []
    for L in lens:
        steps = rng.uniform(0.5, 1.5, L).astype(dtype)
        g = np.cumsum(steps).astype(dtype) + dtype(rng.uniform(-3, 3))
        if dup and L >= 3:
            g[L // 2] = g[L // 2 - 1]
            g[-1] = g[-2]
        axes.append(g)
    return axes


CONFIGS = []
for dtype in (np.float64, np.float32):
    for kind, lens in ((P.FIXED, (7,)), (P.FIXED, (5, 4)), (P.FIXED, (4, 3, 5)), (P.ND, (6,)), (P.ND, (4, 3)),
                       (P.ND, (3, 4, 3)), (P.ND, (3, 2, 3, 2)), (P.FIXED, (1,)), (P.FIXED, (1, 3)), (P.ND, (1, 3)),
                       (P.ND, (2, 1, 2))):
        strategies = [P.LINEAR, P.NEAREST] + ([P.LEFT_NEAREST, P.RIGHT_NEAREST] if kind == P.FIXED and len(lens) == 1 else [])
        for strategy in strategies:
            for extrap in (P.ENABLE, P.FILL, P.CLAMP, P.WRAP, P.ERROR):
                if extrap == P.ENABLE and (strategy != P.LINEAR or min(lens) < 2):
                    continue
                CONFIGS.append((dtype, kind, lens, strategy, extrap))


@pytest.mark.parametrize("dup", [False, True], ids=["distinct", "duplicates"])
def test_oracle_equals_pyref(dup):
    rng = np.random.default_rng(20261019)
    total = 0
    for dtype, kind, lens, strategy, extrap in CONFIGS:
        axes = _axes(rng, lens, dtype, dup=dup)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        per_axis = _special_points(axes, rng, 6, dtype)
        # random combinations of the per-axis candidates
        m = 160
        pts = np.stack([rng.choice(c, m) for c in per_axis], axis=1).astype(dtype)
        o = O.Oracle(axes, values, strategy, extrap, fill=-9.5, kind=kind, dtype=dtype)
        got, st = o.eval_points(pts)
        for i in range(m):
            v, s = P.interpolate(kind, axes, values, strategy, extrap, -9.5, pts[i], dtype=dtype)
            where = (np.dtype(dtype).name, kind, lens, strategy, extrap, pts[i].tolist())
            assert s == st[i], where
            a, b = np.asarray(v, dtype=dtype), got[i]
            assert a.tobytes() == b.tobytes() or (np.isnan(a) and np.isnan(b)), (where, v, got[i])
        total += m
    assert total > 10000


def test_closed_form_bracket_matches_binary_search():
    # SURVEY.md §8a closed form vs strategy/traits.rs:10-33 on random sorted axes (with duplicates)
    rng = np.random.default_rng(7)
    for _ in range(400):
        L = int(rng.integers(2, 40))
        g = np.sort(rng.integers(0, 25, L).astype(np.float64))
        for p in list(g) + list(g + 0.5) + [g[0] - 1, g[-1] + 1, math.nan]:
            assert O.find_nearest_index(g, p) == P.bracket_raw(g, np.float64(p)), (g, p)
