"""Pins the CPU oracle against every golden vector the reference's tests hold for this path."""
import numpy as np
import pytest

import kat_vectors as K
import oracle_binding as O
from kat_runner import run_case


def _oracle_eval(case, points):
    fill = case["fill"] if case["fill"] is not None else 0.0
    o = O.Oracle(case["axes"], case["values"], case["strategy"], case["extrap"], fill=fill, kind=case["kind"])
    rc, _ = o.validate()
    assert rc == 0, case["name"]
    return o.eval_points(points)


@pytest.mark.parametrize("case", K.KATS, ids=[c["name"] for c in K.KATS])
def test_reference_kat(case):
    run_case(case, _oracle_eval)


@pytest.mark.parametrize("c", K.ND_EQUALS_FIXED, ids=[c["name"] for c in K.ND_EQUALS_FIXED])
def test_nd_equals_fixed_bitwise(c):
    fixed = O.Oracle(c["axes"], c["values"], c["strategy"], c["extrap"], kind=O.KIND_FIXED)
    nd = O.Oracle(c["axes"], c["values"], c["strategy"], c["extrap"], kind=O.KIND_ND)
    a, sa = fixed.eval_points(c["points"])
    b, sb = nd.eval_points(c["points"])
    assert (sa == 0).all() and (sb == 0).all()
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("x,lo,hi,want", K.WRAP_KATS)
def test_wrap(x, lo, hi, want):
    # src/lib.rs:89-105
    assert np.float64(O.wrap(x, lo, hi)).tobytes() == np.float64(want).tobytes()


@pytest.mark.parametrize("v", K.VALIDATE_KATS, ids=[v["name"] for v in K.VALIDATE_KATS])
def test_validate(v):
    o = O.Oracle(v["axes"], v["values"], v["strategy"], v["extrap"], kind=v["kind"])
    rc, _ = o.validate()
    assert rc == v["code"]


def test_validate_rules():
    # data.rs:76 empty grid; :80 `<=` so duplicates pass, decreasing/NaN fails; :84 shape mismatch
    assert O.Oracle([[]], np.zeros(0), O.LINEAR, O.ERROR).validate() == (-2, 0)
    assert O.Oracle([[0., 1., 1., 2.]], np.zeros(4), O.LINEAR, O.ERROR).validate() == (0, -1)
    assert O.Oracle([[0., 2., 1.]], np.zeros(3), O.LINEAR, O.ERROR).validate() == (-3, 0)
    assert O.Oracle([[0., float("nan"), 1.]], np.zeros(3), O.LINEAR, O.ERROR).validate() == (-3, 0)
    assert O.Oracle([[0., 1.], [0., 1., 2.]], np.zeros((2, 2)), O.LINEAR, O.ERROR).validate() == (-4, 1)
    # interpolator/mod.rs:97-104: Enable needs >= 2 nodes per axis
    assert O.Oracle([[0., 1.], [3.]], np.zeros((2, 1)), O.LINEAR, O.ENABLE).validate() == (-5, 1)
    # ... but set_extrapolate(Enable) from another policy skips that check (mod.rs:97 reads self.extrapolate)
    assert O.check_extrapolate([2, 1], O.LINEAR, O.ERROR, O.ENABLE) == (0, -1)
    assert O.check_extrapolate([2, 1], O.LINEAR, O.ENABLE, O.CLAMP) == (-5, 1)
    assert O.check_extrapolate([2, 2], O.NEAREST, O.ERROR, O.ENABLE) == (-1, -1)


def test_find_nearest_index_quirks():
    # strategy/traits.rs:10-33
    g = [0., 1., 2., 3.]
    assert O.find_nearest_index(g, 0.0) == 0
    assert O.find_nearest_index(g, 0.5) == 0
    assert O.find_nearest_index(g, 1.0) == 0       # exact node hit -> cell to the left
    assert O.find_nearest_index(g, 2.0) == 1
    assert O.find_nearest_index(g, 3.0) == 2       # == last -> len-2
    assert O.find_nearest_index(g, 3.5) == 3       # beyond last -> len-1 (sic)
    assert O.find_nearest_index(g, -1.0) == 0
    assert O.find_nearest_index(g, float("nan")) == 3
    assert O.find_nearest_index([5.], 5.) == -1    # len-2 underflow -> panic
    assert O.find_nearest_index([0., 1., 2., 2.], 2.) == 2   # duplicates at the end: `== last` wins


def test_degenerate_inputs():
    nan, inf = float("nan"), float("inf")
    lin = lambda e, **kw: O.Oracle([X, X, X], V, O.LINEAR, e, **kw)
    X = [0., 1., 2.]
    V = np.arange(27.).reshape(3, 3, 3)
    assert lin(O.FILL, fill=-7.).one([nan, 0., 0.]) == (-7., 0)
    assert lin(O.ERROR).one([nan, 0., 0.])[1] == 1
    for e in (O.ENABLE, O.CLAMP, O.WRAP):
        assert lin(e).one([0.5, nan, 0.5])[1] == 0xFF
    v, st = lin(O.ENABLE).one([inf, 0.5, 0.5])
    assert st == 0 and (np.isinf(v) or np.isnan(v))
    assert lin(O.CLAMP).one([inf, -inf, 0.])[1] == 0
    assert lin(O.WRAP).one([inf, 0.5, 0.5])[1] == 0xFF
    # length-1 axis in a hard-coded 2-D interpolator panics at interpolate time
    assert O.Oracle([[0., 1.], [3.]], np.zeros((2, 1)), O.LINEAR, O.ERROR).one([0.5, 3.])[1] == 0xFF
    # 1-D and N-D survive through the exact-node path
    assert O.Oracle([[3.]], [9.], O.LINEAR, O.ERROR).one([3.]) == (9., 0)
    assert O.Oracle([[0., 1.], [3.]], np.array([[1.], [2.]]), O.LINEAR, O.ERROR, kind=O.KIND_ND).one([0.5, 3.]) == (1.5, 0)
    # LeftNearest + NaN under Clamp: find_nearest_index -> len-1, values[len-1] is in range: no panic
    assert O.Oracle([X], [1., 2., 3.], O.LEFT_NEAREST, O.CLAMP).one([nan]) == (3., 0)
    assert O.Oracle([X], [1., 2., 3.], O.RIGHT_NEAREST, O.CLAMP).one([nan])[1] == 0xFF
