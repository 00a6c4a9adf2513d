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
