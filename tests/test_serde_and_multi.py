"""serde JSON form of the reference (host-side parsing) and the multi-GPU host batch."""
import json

import numpy as np
import pytest

import kat_vectors as K
import oracle_binding as O

# what `serde_json::to_string(&interp)` produces for three/tests.rs:213-229's interpolator (ndarray serde form)
JSON_3D = json.dumps({
    "data": {"grid": [{"v": 1, "dim": [2], "data": [0.0, 1.0]}] * 3,
             "values": {"v": 1, "dim": [2, 2, 2], "data": [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]}},
    "strategy": "Nearest", "extrapolate": "Error"})


def test_json_parsing_needs_no_gpu():
    from ninterp_b200 import serde
    from ninterp_b200.extrapolate import Extrapolate

    assert serde._extrapolate({"Fill": 2.5}) == Extrapolate.Fill(2.5)
    assert serde._extrapolate(None) == Extrapolate.Error          # #[serde(default)] one/mod.rs:59-60
    assert repr(serde._extrapolate("Wrap")) == "Wrap"
    a = serde._array({"v": 1, "dim": [2, 3], "data": [1, 2, 3, 4, 5, 6]}, np.float64)
    assert a.shape == (2, 3) and a[1, 0] == 4.0
    assert serde.from_json("0.5").interpolate([]) == 0.5          # Interp0D is a bare number (zero/mod.rs:11)


GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"


def test_golden_fixtures_are_reproducible_and_parse():
    # tests/golden/*.json are written by tests/golden/make_serde_fixtures.py from the reference's serde attributes and
    # serde tests (the crate itself cannot run here); regenerate and compare, then parse without a GPU
    import subprocess
    import sys

    before = {f.name: f.read_text() for f in GOLDEN.glob("*.json")}
    subprocess.run([sys.executable, str(GOLDEN / "make_serde_fixtures.py")], check=True)
    assert before == {f.name: f.read_text() for f in GOLDEN.glob("*.json")} and len(before) == 5
    from ninterp_b200 import serde

    doc = json.loads(before["serde_enums_rs_2d_nearest.json"])
    assert doc["strategy"] == "Nearest" and doc["extrapolate"] == "Error"      # untagged Strategy2DEnum == bare name
    assert serde._array(doc["data"]["values"], np.float64).shape == (3, 3)
    assert np.isnan(serde._extrapolate(json.loads(before["serde_one_fill_nan.json"])["extrapolate"]).value)
    assert "extrapolate" not in json.loads(before["serde_one_default_extrapolate.json"])
    assert serde.from_json(before["serde_interp0d.json"]).interpolate([]) == 0.5


@pytest.mark.gpu
def test_golden_serde_fixtures_evaluate_like_the_reference():
    import ninterp_b200 as nb

    # enums.rs:362-373: Interp2D, InterpND and both InterpolatorEnum wrappings serialise to one string, so one fixture
    # loads as each of them; values from two/tests.rs:79-93 (Nearest, incl. the "float imprecision" case)
    text = (GOLDEN / "serde_enums_rs_2d_nearest.json").read_text()
    for kind, cls in (("auto", nb.Interp2D), ("fixed", nb.Interp2D), ("nd", nb.InterpND)):
        g = nb.from_json(text, kind=kind)
        assert isinstance(g, cls) and g.ndim() == 2
        assert g.interpolate([0.05, 0.12]) == 0.0 and g.interpolate([0.07, 0.15 + 0.0001]) == 1.0
        assert g.interpolate([0.08, 0.21]) == 4.0 and g.interpolate([0.14, 0.29]) == 8.0
        with pytest.raises(nb.InterpolateError):
            g.interpolate([0.2, 0.2])
        assert json.loads(nb.to_json(g)) == json.loads(text)
    g = nb.from_json((GOLDEN / "serde_three_tests_nearest.json").read_text())
    assert isinstance(g, nb.Interp3D) and g.interpolate([0.75, 0.25, 0.75]) == 5.0   # three/tests.rs:124
    g = nb.from_json((GOLDEN / "serde_one_fill_nan.json").read_text())              # one/tests.rs:134-147
    assert g.interpolate([1.5]) == 0.5 and g.interpolate([2.0]) == 0.6
    assert np.isnan(g.interpolate([-1.0])) and np.isnan(g.interpolate([5.0]))
    g = nb.from_json((GOLDEN / "serde_one_default_extrapolate.json").read_text())
    assert g.interpolate([1.5]) == 0.4                                              # LeftNearest
    with pytest.raises(nb.InterpolateError):
        g.interpolate([5.0])                                                        # default Extrapolate::Error


@pytest.mark.gpu
def test_from_json_round_trip_and_values():
    import ninterp_b200 as nb

    g = nb.from_json(JSON_3D)
    assert isinstance(g, nb.Interp3D) and g.ndim() == 3
    assert g.interpolate([0.75, 0.25, 0.75]) == 5.0               # three/tests.rs:124
    nd = nb.from_json(JSON_3D, kind="nd")                         # same encoding, N-D interpolator (enums.rs:362-373)
    assert isinstance(nd, nb.InterpND) and nd.interpolate([0.75, 0.25, 0.75]) == 5.0
    again = nb.from_json(nb.to_json(g))
    assert again.interpolate([0.25, 0.75, 0.25]) == 2.0
    fill = nb.from_json(json.dumps({"data": {"grid": [K.X5], "values": K.F5}, "strategy": "Linear", "extrapolate": {"Fill": -1.0}}))
    assert fill.interpolate([9.0]) == -1.0 and fill.interpolate([3.75]) == 0.95


@pytest.mark.gpu
def test_multi_gpu_host_batch_matches_oracle():
    import torch

    import ninterp_b200 as nb
    from gpu_util import assert_same, expected_info

    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]        # on one GPU: three replicas still exercise the split
    rng = np.random.default_rng(31)
    axes = [np.cumsum(rng.uniform(0.5, 1.5, L)) for L in (20, 30, 25)]
    values = rng.uniform(-1, 1, (20, 30, 25))
    rep = nb.ReplicatedInterpolator(lambda d: nb.Interp3D(*axes, values, nb.strategy.Linear, nb.Extrapolate.Error, device=d), devices)
    n = 3_000_001
    cols = [rng.uniform(a[0], a[-1], n) for a in axes]
    cols[0][rng.permutation(n)[:500]] = axes[0][-1] + 2.0
    cols[2][rng.permutation(n)[:5]] = np.nan
    pts = np.stack(cols, axis=1)
    got, st, info = rep.interpolate_batch_host(pts, return_status=True)
    want, st_want = O.Oracle(axes, values, O.LINEAR, O.ERROR).eval(cols, nthreads=O.max_threads())
    assert_same(got, st, want, st_want, "multi-GPU host batch")
    assert info == expected_info(st_want)
    with pytest.raises(nb.InterpolateError):
        rep.interpolate_batch_host(pts)
    rep.close()
