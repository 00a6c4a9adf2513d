"""BASELINE.json's five configurations through the CUDA path: oracle parity at sizes the oracle finishes
in seconds, and size-independent properties at larger sizes."""
import numpy as np
import pytest

import oracle_binding as O

pytestmark = pytest.mark.gpu

LIN, NEA, LEFT, RIGHT = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4


def _run(kind, axes, values, strategy, extrap, cols, dtype=np.float64, fill=0.0):
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    g, o = make_pair(kind, axes, values, strategy, extrap, fill=fill, dtype=dtype)
    got, st, info = gpu_eval(g, cols)
    want, st_want = o.eval(cols, nthreads=O.max_threads())
    assert_same(got, st, want, st_want)
    assert info == expected_info(st_want)
    return got, st


def test_config1_interp1d_linear_error():
    # "Interp1D Linear f64, 1,000-node non-uniform grid, 1e6 query points, Extrapolate::Error"
    rng = np.random.default_rng(1234567890)
    g = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, 999))])
    v = rng.uniform(0, 1, 1000)
    q = rng.uniform(g[0], g[-1], 1_000_000)
    _, st = _run(0, [g], v, LIN, ERROR, [q])
    assert (st == 0).all()
    q[::100] = g[-1] + rng.uniform(0.1, 5, q[::100].size)   # 1 % out of range: the status path
    _, st = _run(0, [g], v, LIN, ERROR, [q])
    assert (st == 1).sum() == q[::100].size
    # the literal benches/benchmark.rs:37-50 shape: 100 uniform nodes 0..99, 1 000 points in [0, 99)
    g2 = np.arange(100.0)
    _run(0, [g2], rng.uniform(0, 1, 100), LIN, ERROR, [rng.uniform(0, 1, 1000) * 99.0])


def test_config2_interp2d_linear_nearest_clamp():
    # "Interp2D Linear + Nearest f64, 2048x2048 grid, 1e8 query points, Extrapolate::Clamp" at 2e6 points
    rng = np.random.default_rng(2)
    v = rng.uniform(0, 1, (2048, 2048))
    for axes in ([np.arange(2048.0)] * 2, [np.cumsum(rng.uniform(0.5, 1.5, 2048)) for _ in range(2)]):
        cols = [rng.uniform(a[0] - 0.05 * (a[-1] - a[0]), a[-1] + 0.05 * (a[-1] - a[0]), 2_000_000) for a in axes]
        for s in (LIN, NEA):
            _run(0, axes, v, s, CLAMP, cols)


def test_config3_interp3d_linear_256cubed():
    # "Interp3D Linear f64, 256^3 uniform grid" at 4e6 points, uniform and non-uniform axes
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 1, (256, 256, 256))
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    for axes, route in (([np.arange(256) * (1.0 / 255)] * 3, 3), ([np.cumsum(rng.uniform(0.5, 1.5, 256)) for _ in range(3)], 1)):
        # 6e6 points: above the batch size from which a table of this size is routed by the order probe
        cols = [rng.uniform(a[0], a[-1], 6_000_000) for a in axes]
        g, o = make_pair(0, axes, v, LIN, ERROR, dtype=np.float64)
        want, st_want = o.eval(cols, nthreads=O.max_threads())
        got, st, info = gpu_eval(g, cols)
        assert_same(got, st, want, st_want, "random order")
        assert info == expected_info(st_want) and (st == 0).all()
        assert _lib.lib().nint_last_query_order(g._h) == 2  # incoherent ...
        # ... slab lists on the cell-packed table (uniform axes) or slab passes over the plain table (non-uniform)
        assert _lib.lib().nint_last_incoherent_route(g._h) == route
        # cell-sorted order must give the same values as the unsorted batch, through the single pass
        order = np.lexsort([np.searchsorted(axes[2], cols[2]), np.searchsorted(axes[1], cols[1]),
                            np.searchsorted(axes[0], cols[0])])
        got, st, info = gpu_eval(g, [c[order] for c in cols])
        assert_same(got, st, want[order], st_want[order], "cell order")
        assert _lib.lib().nint_last_query_order(g._h) == 1


def test_config3_wrap_routes():
    # the same table under Extrapolate::Wrap (coordinates are re-mapped across axes, so the slab passes do not apply):
    # uniform axes take the slab lists (points are classified by their wrapped axis-0 coordinate); non-uniform axes are
    # bucketed by grid tile and evaluated from TMA-staged tiles in shared memory
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    rng = np.random.default_rng(33)
    v = rng.uniform(0, 1, (256, 256, 256))
    for axes, route in (([np.arange(256) * (1.0 / 255)] * 3, 3), ([np.cumsum(rng.uniform(0.5, 1.5, 256)) for _ in range(3)], 2)):
        span = [a[-1] - a[0] for a in axes]
        cols = [rng.uniform(a[0] - 0.3 * w, a[-1] + 0.3 * w, 6_000_000) for a, w in zip(axes, span)]
        g, o = make_pair(0, axes, v, LIN, WRAP, dtype=np.float64)
        want, st_want = o.eval(cols, nthreads=O.max_threads())
        got, st, info = gpu_eval(g, cols)
        assert_same(got, st, want, st_want, "wrap, random order")
        assert info == expected_info(st_want)
        assert _lib.lib().nint_last_query_order(g._h) == 2 and _lib.lib().nint_last_incoherent_route(g._h) == route


def test_config4_interpnd_f32_n5_wrap_fill():
    # "InterpND Linear f32 N=5, 32^5 non-uniform grid, Extrapolate::Wrap and Fill" at 1e6 points
    rng = np.random.default_rng(4)
    axes = [np.cumsum(rng.uniform(0.5, 1.5, 32)).astype(np.float32) for _ in range(5)]
    v = rng.uniform(0, 1, (32,) * 5).astype(np.float32)
    cols = [rng.uniform(a[0] - 0.25 * (a[-1] - a[0]), a[-1] + 0.25 * (a[-1] - a[0]), 1_000_000).astype(np.float32) for a in axes]
    _run(1, axes, v, LIN, WRAP, cols, dtype=np.float32)
    _run(1, axes, v, LIN, FILL, cols, dtype=np.float32, fill=-1.0)
    _run(1, axes, v, NEA, WRAP, cols, dtype=np.float32)
    # at 6e6 points the default route of an incoherent batch under Wrap is the slab lists with five columns (nint_part.cuh)
    from gpu_util import assert_same, gpu_eval, make_pair

    from ninterp_b200 import _lib

    cols = [rng.uniform(a[0] - 0.25 * (a[-1] - a[0]), a[-1] + 0.25 * (a[-1] - a[0]), 6_000_000).astype(np.float32) for a in axes]
    g, o = make_pair(1, axes, v, LIN, WRAP, dtype=np.float32)
    got, st, _ = gpu_eval(g, cols)
    assert _lib.lib().nint_last_query_order(g._h) == 2 and _lib.lib().nint_last_incoherent_route(g._h) == 3
    assert_same(got, st, *o.eval(cols, nthreads=O.max_threads()), "config 4 under Wrap, slab lists")


def test_config5_enable_mixed_and_1d_index_strategies():
    # "Interp3D Linear f64 with Extrapolate::Enable and out-of-range/NaN-mixed queries plus
    #  LeftNearest/RightNearest, bit-exact/ulp sweep"
    rng = np.random.default_rng(5)
    axes = [np.cumsum(rng.uniform(0.5, 1.5, 64)) for _ in range(3)]
    v = rng.uniform(-1, 1, (64, 64, 64))
    n = 2_000_000
    cols = []
    for a in axes:
        span = a[-1] - a[0]
        c = rng.uniform(a[0] - 0.5 * span, a[-1] + 0.5 * span, n)
        k = rng.permutation(n)[: n // 5]
        kind = rng.integers(0, 8, k.size)
        node = a[rng.integers(0, 64, k.size)]
        sp = np.select([kind == 0, kind == 1, kind == 2, kind == 3, kind == 4, kind == 5, kind == 6],
                       [node, np.nextafter(node, np.inf), np.nextafter(node, -np.inf), 0.0, -0.0, np.inf, -np.inf], np.nan)
        sp = np.where(kind == 7, np.where(rng.random(k.size) < 0.5, np.nan, 5e-324), sp)
        c[k] = sp
        cols.append(c)
    got, st = _run(0, axes, v, LIN, ENABLE, cols)
    assert (st == 0xFF).sum() > 0 and (st == 0).sum() > n // 2
    g1 = np.cumsum(rng.uniform(0.5, 1.5, 1000))
    v1 = rng.uniform(-1, 1, 1000)
    q = cols[0][:1_000_000] * (g1[-1] / axes[0][-1])
    q[::7] = g1[rng.integers(0, 1000, q[::7].size)]
    for s in (LEFT, RIGHT, NEA):
        for e in (CLAMP, WRAP, ERROR, FILL):
            _run(0, [g1], v1, s, e, [q], fill=7.0)


def test_large_batch_properties():
    # size-independent checks at 5e7 points (no oracle): Nearest returns table entries only; Linear at
    # grid nodes reproduces the table; Clamp equals evaluation at clamped coordinates; linearity in the table.
    import ninterp_b200 as nb
    import torch

    rng = np.random.default_rng(6)
    L = 128
    axes = [np.arange(L) * (1.0 / (L - 1))] * 3
    v = rng.uniform(0, 1, (L, L, L))
    n = 50_000_000
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(6)
    cols = [torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 1.2 - 0.1 for _ in range(3)]
    lin = nb.Interp3D(*axes, v, nb.strategy.Linear, nb.Extrapolate.Clamp)
    out = lin.interpolate_batch(cols)
    clamped = [c.clamp(0.0, 1.0) for c in cols]
    out_c = lin.interpolate_batch(clamped)
    assert torch.equal(out, out_c)
    assert float(out.min()) >= v.min() and float(out.max()) <= v.max()
    # linearity: interp(a*V1 + V2) is within rounding of a*interp(V1) + interp(V2)
    v2 = rng.uniform(0, 1, (L, L, L))
    lin2 = nb.Interp3D(*axes, v2, nb.strategy.Linear, nb.Extrapolate.Clamp)
    lin3 = nb.Interp3D(*axes, 2.0 * v + v2, nb.strategy.Linear, nb.Extrapolate.Clamp)
    comb = lin3.interpolate_batch(cols)
    ref = 2.0 * out + lin2.interpolate_batch(cols)
    assert float((comb - ref).abs().max()) < 1e-13
    # nearest: every output is a table entry (checksum of membership via sorted unique lookup)
    near = nb.Interp3D(*axes, v, nb.strategy.Nearest, nb.Extrapolate.Clamp)
    o = near.interpolate_batch(cols)
    table = torch.from_numpy(np.sort(v.reshape(-1))).to(dev)
    pos = torch.searchsorted(table, o).clamp(max=table.numel() - 1)
    assert torch.equal(table[pos], o)
    # nodes: Linear at exact grid nodes returns the table bit-for-bit (three/tests.rs:22-31)
    idx = torch.randint(0, L, (3, 10_000_000), device=dev)
    ax_t = torch.from_numpy(axes[0]).to(dev)
    node_cols = [ax_t[idx[k]].contiguous() for k in range(3)]
    vt = torch.from_numpy(v).to(dev)
    assert torch.equal(lin.interpolate_batch(node_cols), vt[idx[0], idx[1], idx[2]])


def test_batch_larger_than_one_launch():
    # the kernels index points with 32 bits and the host splits batches beyond 2^30 points into several launches:
    # values, statuses and the batch-wide indices in nint_batch_info must be unaffected by the split
    import ninterp_b200 as nb
    import torch

    dev = torch.device("cuda")
    n = (1 << 30) + 12345
    g = np.arange(64, dtype=np.float32)
    v = (np.arange(64, dtype=np.float32) * 3.0).astype(np.float32)
    interp = nb.Interp1D(g, v, nb.strategy.Linear, nb.Extrapolate.Error, dtype=np.float32)
    x = torch.empty(n, dtype=torch.float32, device=dev)
    x.uniform_(0.0, 63.0)
    bad = [5, (1 << 30) - 1, (1 << 30) + 7, n - 1]
    x[bad] = 100.0
    out = torch.empty(n, dtype=torch.float32, device=dev)
    status = torch.empty(n, dtype=torch.uint8, device=dev)
    info = torch.zeros(4, dtype=torch.int64, device=dev)
    interp.interpolate_batch([x], out=out, status=status, info=info)
    torch.cuda.synchronize()
    i = info.cpu().numpy().view(np.uint64)
    assert int(i[0]) == 4 and int(i[1]) == 0 and int(i[2]) == 5
    assert int(status.sum()) == 4 and status[bad].tolist() == [1, 1, 1, 1]
    ok = status == 0
    # f(x) = 3x on a uniform grid: exact up to rounding of the blend
    assert float((out[ok] - 3.0 * x[ok]).abs().max()) < 1e-3
    # second half (after the split) against the oracle on a sample
    sl = slice((1 << 30) - 50000, (1 << 30) + 50000)
    want, st = O.Oracle([g], v, LIN, ERROR, dtype=np.float32).eval([x[sl].cpu().numpy()])
    from gpu_util import assert_same

    assert_same(out[sl].cpu().numpy(), status[sl].cpu().numpy(), want, st, "split launch")
