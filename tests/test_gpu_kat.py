"""The reference's golden vectors through the CUDA path (C ABI), scalar and batch entry points."""
import numpy as np
import pytest

import kat_vectors as K
from kat_runner import run_case

pytestmark = pytest.mark.gpu


def _interp(case):
    from gpu_util import make_pair

    fill = case["fill"] if case["fill"] is not None else 0.0
    g, _ = make_pair(case["kind"], case["axes"], case["values"], case["strategy"], case["extrap"], fill=fill)
    return g


def _batch_eval(case, points):
    from gpu_util import gpu_eval

    g = _interp(case)
    pts = np.asarray(points, dtype=np.float64).reshape(len(points), -1)
    out, st, _ = gpu_eval(g, [pts[:, k] for k in range(pts.shape[1])])
    return out, st


def _scalar_eval(case, points):
    """`interp.interpolate(&[..])` one point at a time, Result mapped back to a status."""
    import ninterp_b200 as nb

    g = _interp(case)
    out = np.empty(len(points))
    st = np.zeros(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        try:
            out[i] = g.interpolate(list(p))
        except nb.InterpolateError as e:
            assert e.variant == "ExtrapolateError"
            out[i], st[i] = np.nan, 1
        except nb.ReferencePanic:
            out[i], st[i] = np.nan, 0xFF
    return out, st


@pytest.mark.parametrize("case", K.KATS, ids=[c["name"] for c in K.KATS])
def test_reference_kat_batch(case):
    run_case(case, _batch_eval)


@pytest.mark.parametrize("case", K.KATS, ids=[c["name"] for c in K.KATS])
def test_reference_kat_scalar(case):
    run_case(case, _scalar_eval)


@pytest.mark.parametrize("c", K.ND_EQUALS_FIXED, ids=[c["name"] for c in K.ND_EQUALS_FIXED])
def test_nd_equals_fixed_bitwise(c):
    from gpu_util import gpu_eval, make_pair

    fixed, _ = make_pair(0, c["axes"], c["values"], c["strategy"], c["extrap"])
    nd, _ = make_pair(1, c["axes"], c["values"], c["strategy"], c["extrap"])
    pts = np.asarray(c["points"])
    cols = [pts[:, k] for k in range(pts.shape[1])]
    a, sa, _ = gpu_eval(fixed, cols)
    b, sb, _ = gpu_eval(nd, cols)
    assert (sa == 0).all() and (sb == 0).all()
    assert a.tobytes() == b.tobytes()


def test_point_length():
    # one/tests.rs:12-15: interpolate(&[]) on a 1-D interpolator is PointLength
    import ninterp_b200 as nb

    g = nb.Interp1D(K.X5, K.F5, nb.strategy.Linear, nb.Extrapolate.Error)
    with pytest.raises(nb.InterpolateError) as e:
        g.interpolate([])
    assert e.value.variant == "PointLength"
    assert g.interpolate([1.0]) == 0.4
    g3 = nb.Interp3D(K.X3, K.Y3, K.Z3, K.F333)
    with pytest.raises(nb.InterpolateError) as e:
        g3.interpolate([0.1, 0.2])
    assert e.value.variant == "PointLength"


def test_set_strategy_and_extrapolate():
    # strategy/enums/mod.rs:57-79 (set_strategy keeps data), three/mod.rs:240-244
    import ninterp_b200 as nb

    g = nb.Interp1D(K.X5, K.F5, nb.strategy.Linear, nb.Extrapolate.Error)
    assert g.interpolate([3.75]) == 0.95
    g.set_strategy(nb.strategy.Nearest)
    assert g.interpolate([3.25]) == 0.8 and g.interpolate([3.5]) == 1.0
    g.set_strategy(nb.strategy.LeftNearest)
    assert g.interpolate([3.75]) == 0.8
    with pytest.raises(nb.ValidateError) as e:
        g.set_extrapolate(nb.Extrapolate.Enable)
    assert e.value.variant == "ExtrapolateSelection"
    g.set_strategy(nb.strategy.Linear)
    g.set_extrapolate(nb.Extrapolate.Enable)
    assert g.interpolate([5.0]) == 1.2
    g.set_extrapolate(nb.Extrapolate.Fill(-3.0))
    assert g.interpolate([5.0]) == -3.0
    g.validate()


def test_zero_d():
    # n/tests.rs:362-369 and benches/benchmark.rs:24-33: InterpND over one value with an empty grid
    import ninterp_b200 as nb
    import torch

    g = nb.InterpND([[]], [0.5])
    assert g.ndim() == 0
    assert g.interpolate([]) == 0.5
    out = torch.empty(1000, dtype=torch.float64, device="cuda")
    g.interpolate_batch([], out=out)
    assert (out.cpu().numpy() == 0.5).all()
    assert nb.Interp0D(0.5).interpolate([]) == 0.5
