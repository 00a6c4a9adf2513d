"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact values and statuses.

Bar (BASELINE.json north_star): bit-exact for Nearest/LeftNearest/RightNearest; <=4 ulp (f64) /
<=2 ulp (f32) for Linear.  The kernels execute the reference's operation sequence without FMA
contraction, so Linear is held to bit-exact here as well (0 ulp).
"""
import numpy as np
import pytest

import oracle_binding as O

pytestmark = pytest.mark.gpu

LIN, NEA, LEFT, RIGHT = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4
FIXED, ND = 0, 1


def _axes(rng, lens, dtype, mode="nonuniform"):
    axes = []
    for L in lens:
        if mode == "uniform":
            g = (np.arange(L) * 0.37 - 1.3).astype(dtype)
        elif mode == "analytic":  # nodes are exactly k * h in `dtype`: the bin kernel computes them instead of loading them
            g = np.arange(L, dtype=dtype) * dtype(0.37)
        elif mode == "log":
            g = np.sort(np.unique(np.logspace(-6, 3, L).astype(dtype)))
            if len(g) < L:
                g = np.concatenate([g, g[-1] + 1 + np.arange(L - len(g))]).astype(dtype)
        elif mode == "dup":
            g = np.cumsum(rng.uniform(0.5, 1.5, L)).astype(dtype)
            if L >= 4:
                g[L // 2] = g[L // 2 - 1]
                g[-1] = g[-2]
                g[1] = g[0]
        elif mode == "offset":  # huge offset, tiny span: bucket edges are below one ulp of the nodes
            base = dtype(1e15 if dtype == np.float64 else 1e7)
            g = np.unique((base + np.cumsum(rng.uniform(1, 3, L))).astype(dtype))
            while len(g) < L:
                g = np.append(g, np.nextafter(g[-1], dtype(np.inf)) + dtype(4))
            g = g[:L].astype(dtype)
        elif mode == "wide":  # spans the whole finite range: the bucket width overflows
            g = np.sort(np.concatenate([[-np.finfo(dtype).max], rng.uniform(-1e3, 1e3, L - 2), [np.finfo(dtype).max]])).astype(dtype)
        else:
            g = (np.cumsum(rng.uniform(0.5, 1.5, L)) + rng.uniform(-3, 3)).astype(dtype)
        axes.append(np.ascontiguousarray(g))
    return axes


def _queries(rng, axes, n, dtype, frac_special=0.25):
    """Random queries over [-0.3 span, 1.3 span] mixed with nodes, neighbours of nodes, +-0, inf, NaN, denormals."""
    cols = []
    for g in axes:
        lo, hi = float(g[0]), float(g[-1])
        if not np.isfinite(hi - lo):
            lo, hi = -2e3, 2e3
        span = max(hi - lo, 1e-3)
        c = rng.uniform(lo - 0.3 * span, hi + 0.3 * span, n).astype(dtype)
        k = int(n * frac_special)
        pick = rng.integers(0, len(g), k)
        kind = rng.integers(0, 9, k)
        special = g[pick].astype(dtype)
        special = np.where(kind == 1, np.nextafter(g[pick], dtype(np.inf)), special)
        special = np.where(kind == 2, np.nextafter(g[pick], dtype(-np.inf)), special)
        special = np.where(kind == 3, dtype(np.nan), special)
        special = np.where(kind == 4, dtype(np.inf) * rng.choice([-1, 1], k).astype(dtype), special)
        special = np.where(kind == 5, dtype(0.0) * rng.choice([-1, 1], k).astype(dtype), special)
        special = np.where(kind == 6, np.finfo(dtype).tiny * dtype(0.5), special)
        special = np.where(kind == 7, np.finfo(dtype).max * rng.choice([-1, 1], k).astype(dtype), special)
        c[rng.permutation(n)[:k]] = special.astype(dtype)
        cols.append(c)
    return cols


def _check(kind, lens, strategy, extrap, dtype, mode, n, seed, aos=False, frac_special=0.25):
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    rng = np.random.default_rng(seed)
    axes = _axes(rng, lens, dtype, mode)
    values = rng.uniform(-1, 1, lens).astype(dtype)
    g, o = make_pair(kind, axes, values, strategy, extrap, fill=-9.5, dtype=dtype)
    cols = _queries(rng, axes, n, dtype, frac_special)
    got, st, info = gpu_eval(g, cols, aos=aos)
    want, st_want = o.eval(cols, nthreads=O.max_threads())
    ctx = f"kind={kind} lens={lens} strat={strategy} extrap={extrap} {np.dtype(dtype).name} {mode}"
    assert_same(got, st, want, st_want, ctx)
    assert info == expected_info(st_want), ctx


FIXED_CASES = [(FIXED, (33,)), (FIXED, (17, 9)), (FIXED, (9, 7, 11))]
ND_CASES = [(ND, (13,)), (ND, (7, 5)), (ND, (5, 4, 6)), (ND, (4, 3, 4, 3)), (ND, (3, 4, 3, 2, 3)), (ND, (2, 3, 2, 3, 2, 2)),
            (ND, (2, 3, 2, 2, 3, 2, 2)), (ND, (2, 2, 3, 2, 2, 2, 3, 2))]


def _combos(cases):
    out = []
    for kind, lens in cases:
        strategies = [LIN, NEA] + ([LEFT, RIGHT] if kind == FIXED and len(lens) == 1 else [])
        for s in strategies:
            for e in (ENABLE, FILL, CLAMP, WRAP, ERROR):
                if e == ENABLE and s != LIN:
                    continue
                out.append((kind, lens, s, e))
    return out


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind,lens,strategy,extrap", _combos(FIXED_CASES + ND_CASES))
def test_all_policies_adversarial(kind, lens, strategy, extrap, dtype):
    _check(kind, lens, strategy, extrap, dtype, "nonuniform", 20000, seed=hash((kind, lens, strategy, extrap)) % 2**31)


@pytest.mark.parametrize("mode", ["uniform", "log", "dup", "offset", "wide"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_grid_shapes_stress_bracket(mode, dtype):
    # the bucket table must never change the index the reference's binary search would return
    for kind, lens, strategy in [(FIXED, (200,), LIN), (FIXED, (200,), LEFT), (FIXED, (64, 48), NEA), (FIXED, (20, 30, 25), LIN),
                                 (ND, (12, 10, 9, 8), LIN)]:
        for extrap in (ENABLE, CLAMP, WRAP, ERROR):
            if extrap == ENABLE and strategy != LIN:
                continue
            _check(kind, lens, strategy, extrap, dtype, mode, 30000, seed=11)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_length_one_and_two_axes(dtype):
    for kind, lens in [(FIXED, (1,)), (FIXED, (2,)), (FIXED, (1, 5)), (FIXED, (4, 1, 3)), (ND, (1, 4)), (ND, (3, 1, 2)),
                       (ND, (2, 2, 1, 2))]:
        strategies = [LIN, NEA] + ([LEFT, RIGHT] if kind == FIXED and len(lens) == 1 else [])
        for s in strategies:
            for e in (FILL, CLAMP, WRAP, ERROR) + ((ENABLE,) if s == LIN and min(lens) >= 2 else ()):
                _check(kind, lens, s, e, dtype, "nonuniform", 4000, seed=5, frac_special=0.5)


def test_nd_single_value_needs_empty_grid():
    # n/mod.rs:71-83, n/tests.rs:370-381: an InterpND over one value is 0-D and rejects a non-empty grid
    import ninterp_b200 as nb

    for grid, vals in (([[1.0]], [0.0]), ([[1.0], [2.0]], [[0.0]])):
        with pytest.raises(nb.ValidateError) as e:
            nb.InterpND(grid, vals)
        assert e.value.variant == "Other"


def test_aos_points_equal_soa_columns():
    for kind, lens, s in [(FIXED, (40,), LIN), (FIXED, (12, 9), LIN), (FIXED, (9, 8, 7), LIN), (ND, (4, 5, 3, 4, 3), NEA)]:
        _check(kind, lens, s, CLAMP, np.float64, "nonuniform", 50000, seed=3, aos=True)
        _check(kind, lens, s, ERROR, np.float32, "nonuniform", 50000, seed=4, aos=True)


def test_axes_too_large_for_shared_memory():
    # 60k nodes * 8 B * 2 axes > the shared-memory budget: the global-memory variant must agree as well
    _check(FIXED, (60000,), LIN, ERROR, np.float64, "nonuniform", 200000, seed=8, frac_special=0.05)
    _check(FIXED, (60000,), RIGHT, WRAP, np.float64, "nonuniform", 200000, seed=8, frac_special=0.05)
    _check(FIXED, (30000, 20), LIN, ENABLE, np.float64, "nonuniform", 200000, seed=9, frac_special=0.05)
    _check(ND, (30000, 8), NEA, CLAMP, np.float64, "nonuniform", 200000, seed=10, frac_special=0.05)


def test_strided_value_views():
    # "view data": InterpData::view (data.rs:92-97) — non-contiguous values are honoured at upload
    import ninterp_b200 as nb
    from gpu_util import assert_same, gpu_eval

    rng = np.random.default_rng(12)
    big = rng.uniform(-1, 1, (14, 18, 10))
    view = big[::2, 1::3, ::-1]                  # strides (2, 3, -1) elements
    view_t = np.transpose(big, (2, 0, 1))[2:9]   # permuted axes
    for v in (view, view_t):
        axes = [np.cumsum(rng.uniform(0.5, 1.5, L)) for L in v.shape]
        g = nb.Interp3D(*axes, v, nb.strategy.Linear, nb.Extrapolate.Clamp)
        o = O.Oracle(axes, np.ascontiguousarray(v), LIN, CLAMP)
        cols = _queries(rng, axes, 20000, np.float64, 0.1)
        got, st, _ = gpu_eval(g, cols)
        want, st_want = o.eval(cols)
        assert_same(got, st, want, st_want, "strided view")


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_host_pipeline_multi_chunk(layout):
    # nint_interpolate_batch_host: > 2 chunks of 4 Mi points, status + info, ragged last chunk
    import ninterp_b200 as nb

    rng = np.random.default_rng(21)
    axes = [np.cumsum(rng.uniform(0.5, 1.5, L)) for L in (50, 40, 30)]
    values = rng.uniform(-1, 1, (50, 40, 30))
    g = nb.Interp3D(*axes, values, nb.strategy.Linear, nb.Extrapolate.Error)
    o = O.Oracle(axes, values, LIN, ERROR)
    n = (4 << 20) * 2 + 12345
    cols = [rng.uniform(a[0], a[-1], n) for a in axes]
    bad = rng.permutation(n)[:1000]
    cols[1][bad] = axes[1][-1] + 1.0
    cols[2][bad[:10]] = np.nan
    pts = cols if layout == "soa" else np.stack(cols, axis=1)
    got, st, info = g.interpolate_batch_host(pts, return_status=True)
    want, st_want = o.eval(cols, nthreads=O.max_threads())
    from gpu_util import assert_same, expected_info

    assert_same(got, st, want, st_want, "host pipeline")
    assert info == expected_info(st_want)
    with pytest.raises(nb.InterpolateError):
        g.interpolate_batch_host(pts)
    # a clean batch returns values only
    ok = [rng.uniform(a[0], a[-1], 100000) for a in axes]
    vals = g.interpolate_batch_host(ok if layout == "soa" else np.stack(ok, axis=1))
    assert_same(vals, np.zeros(100000, np.uint8), *o.eval(ok), "host pipeline clean")


def test_empty_batch():
    import ninterp_b200 as nb
    import torch

    g = nb.Interp2D([0., 1.], [0., 1.], [[0., 1.], [2., 3.]])
    out = g.interpolate_batch([torch.empty(0, dtype=torch.float64, device="cuda")] * 2)
    assert out.numel() == 0
    assert g.interpolate_batch_host(np.empty((0, 2))).size == 0


@pytest.mark.parametrize("dtype_code,name", [(1, "f64"), (0, "f32")])
def test_division_fast_path_is_ieee_exact(dtype_code, name):
    # div_cell(): q = a*r corrected by two FMA residual steps must equal IEEE a / w bit-for-bit
    import ctypes as C

    from ninterp_b200 import _lib

    lib = _lib.lib()
    bad, fast = C.c_uint64(), C.c_uint64()
    n = 2_000_000_000
    for seed in (1, 2):
        assert lib.nint_selftest_divide(dtype_code, n, seed, C.byref(bad), C.byref(fast)) == 0
        assert bad.value == 0, f"{name}: {bad.value} of {n} quotients differ from IEEE division"
        assert fast.value > n // 4, f"{name}: fast path barely exercised ({fast.value})"


def test_more_than_eight_dimensions_is_rejected():
    import ninterp_b200 as nb

    with pytest.raises(nb.EngineError):
        nb.InterpND([[0., 1.]] * 9, np.zeros((2,) * 9))


def test_plain_table_path_matches_packed(monkeypatch):
    # NINT_PACK=0 keeps Linear on the plain C-order table (paired row loads) instead of the cell-packed copy
    from gpu_util import assert_same, gpu_eval, make_pair

    rng = np.random.default_rng(77)
    for kind, lens, dtype in [(FIXED, (40,), np.float64), (FIXED, (21, 18), np.float32), (FIXED, (15, 16, 17), np.float64),
                              (ND, (6, 5, 7, 4), np.float32)]:
        axes = _axes(rng, lens, dtype, "uniform" if len(lens) == 3 else "nonuniform")
        values = rng.uniform(-1, 1, lens).astype(dtype)
        cols = _queries(rng, axes, 60000, dtype, 0.1)
        for extrap in (ENABLE, CLAMP, WRAP):
            monkeypatch.setenv("NINT_PACK", "0")
            g0, o = make_pair(kind, axes, values, LIN, extrap, dtype=dtype)
            monkeypatch.setenv("NINT_PACK", "1")
            g1, _ = make_pair(kind, axes, values, LIN, extrap, dtype=dtype)
            a, sa, _ = gpu_eval(g0, cols)
            b, sb, _ = gpu_eval(g1, cols)
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            assert_same(a, sa, want, st_want, f"plain {lens} {extrap}")
            assert_same(b, sb, want, st_want, f"packed {lens} {extrap}")


def test_slab_passes_and_order_probe(monkeypatch):
    # Linear on tables beyond L2 lets an on-device probe choose between one pass on the cell-packed table
    # (coherent queries) and slab passes on the plain table (incoherent ones).  Force the mode on small tables.
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    monkeypatch.setenv("NINT_SLAB_MIN_BYTES", "0")
    monkeypatch.setenv("NINT_SLAB_MIN_POINTS", "0")
    monkeypatch.setenv("NINT_COMPACT", "0")  # 4-D Fill/Error batches would otherwise take the in-range compaction pass
    rng = np.random.default_rng(99)
    n = 300_001
    for kind, lens, dtype, mode, passes in [(FIXED, (40, 9, 7), np.float64, "uniform", 3), (FIXED, (33, 12), np.float32, "nonuniform", 2),
                                            (FIXED, (64,), np.float64, "nonuniform", 5), (ND, (17, 5, 4, 3), np.float32, "nonuniform", 4),
                                            (ND, (20, 6, 5), np.float64, "uniform", 1)]:
        monkeypatch.setenv("NINT_SLABS", str(passes))
        # the probe counts distinct cells (shift 0) per window of 4096 consecutive points: a random window touches
        # nearly every cell of these small grids, a window sorted along axis 0 one or two axis-0 layers
        monkeypatch.setenv("NINT_PROBE_SHIFT", "0")
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", str(int(np.prod([l - 1 for l in lens])) // 3))
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        for extrap in (ERROR, CLAMP, ENABLE, FILL):
            g, o = make_pair(kind, axes, values, LIN, extrap, fill=-2.5, dtype=dtype)
            cols = _queries(rng, axes, n, dtype, 0.05)
            # every axis-0 node (the pass boundaries are nodes), their neighbours, and the non-finite values
            a0 = np.asarray(axes[0], dtype=dtype)
            special = np.concatenate([a0, np.nextafter(a0, dtype(np.inf)), np.nextafter(a0, dtype(-np.inf)),
                                      np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
            where = rng.choice(n, special.size, replace=False)
            cols[0][where] = special
            # incoherent order
            got, st, info = gpu_eval(g, cols)
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            assert_same(got, st, want, st_want, f"slab random {lens} {extrap}")
            assert info == expected_info(st_want)
            assert _lib.lib().nint_last_query_order(g._h) == 2
            assert _lib.lib().nint_last_incoherent_route(g._h) == 1
            # coherent order: sorted along axis 0
            order = np.argsort(cols[0], kind="stable")
            cols_s = [c[order] for c in cols]
            got, st, info = gpu_eval(g, cols_s)
            assert_same(got, st, want[order], st_want[order], f"slab sorted {lens} {extrap}")
            assert info == expected_info(st_want[order])
            assert _lib.lib().nint_last_query_order(g._h) == 1
            # array-of-points input takes the same route
            got, st, _ = gpu_eval(g, cols, aos=True)
            assert_same(got, st, want, st_want, f"slab aos {lens} {extrap}")
    # Wrap re-maps every axis when one is out of range, so it stays on the single-pass kernels
    monkeypatch.setenv("NINT_SLABS", "2")
    g, o = make_pair(FIXED, _axes(rng, (30, 8), np.float64), rng.uniform(-1, 1, (30, 8)), LIN, WRAP)
    cols = _queries(rng, g.data.grid, 50_000, np.float64, 0.05)
    got, st, _ = gpu_eval(g, cols)
    assert_same(got, st, *o.eval(cols), "wrap stays single pass")
    assert _lib.lib().nint_last_query_order(g._h) == 0


def test_pipelined_single_pass(monkeypatch):
    # the TMA-pipelined form of the single pass (nint_pipe.cuh): coordinates arrive in shared memory by cp.async.bulk,
    # whole chunks of 512 points; the ragged tail and unaligned / array-of-points inputs go to interp_kernel
    import torch

    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    monkeypatch.setenv("NINT_PIPE", "1")
    rng = np.random.default_rng(616)
    n = 1_100_003
    for kind, lens, dtype, mode in [(FIXED, (40, 33, 26), np.float64, "nonuniform"), (FIXED, (57, 31), np.float64, "uniform"),
                                    (FIXED, (19, 22, 35), np.float32, "uniform"), (FIXED, (64, 48), np.float32, "nonuniform")]:
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        cols = _queries(rng, axes, n, dtype, 0.1)
        for strat, extrap in [(LIN, e) for e in (ERROR, CLAMP, ENABLE, FILL, WRAP)] + [(NEA, e) for e in (ERROR, CLAMP, WRAP)]:
            g, o = make_pair(kind, axes, values, strat, extrap, fill=0.5, dtype=dtype)
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            got, st, info = gpu_eval(g, cols)
            assert_same(got, st, want, st_want, f"pipe {lens} {dtype.__name__} {strat} {extrap}")
            assert info == expected_info(st_want)
        # columns that start 8 bytes off a 16-byte boundary: not eligible, same results
        g, o = make_pair(kind, axes, values, LIN, CLAMP, dtype=dtype)
        dev = torch.device("cuda", g.device)
        shifted = [torch.from_numpy(np.concatenate([c[:1], c])).to(dev)[1:] for c in cols]
        out = g.interpolate_batch(shifted)
        torch.cuda.synchronize()
        want, _ = o.eval(cols, nthreads=O.max_threads())
        assert np.array_equal(out.cpu().numpy().view(np.uint8), want.view(np.uint8))


def test_inrange_compaction_pass():
    # InterpND with four or more axes under Fill / Error: the scan finishes the out-of-range points itself and the
    # in-range ones are compacted into full warps before their 2^N corners are gathered (slab_kernel<INR>)
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    rng = np.random.default_rng(808)
    for lens, dtype, mode in [((6, 5, 7, 4), np.float64, "nonuniform"), ((5, 4, 6, 3, 4), np.float32, "nonuniform"),
                              ((4, 3, 3, 4, 3, 3), np.float64, "uniform"), ((9, 3, 1, 4), np.float32, "nonuniform")]:
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        for n in (70_001, 300_000):
            cols = _queries(rng, axes, n, dtype, 0.2)   # +-30 % of the span: most points are outside on some axis
            for extrap in (FILL, ERROR):
                g, o = make_pair(ND, axes, values, LIN, extrap, fill=7.25, dtype=dtype)
                want, st_want = o.eval(cols, nthreads=O.max_threads())
                got, st, info = gpu_eval(g, cols)
                assert_same(got, st, want, st_want, f"compaction {lens} {extrap} n={n}")
                assert info == expected_info(st_want)
                got, st, _ = gpu_eval(g, cols, aos=True)
                assert_same(got, st, want, st_want, f"compaction aos {lens} {extrap}")


def test_analytic_uniform_axes(monkeypatch):
    # axes whose nodes are exactly k * h (arange * h) may have their nodes computed instead of loaded, with IEEE division
    # for the weights (NINT_ANALYTIC=1 on the single pass; the bin kernel of the tile route always does): same bits
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    monkeypatch.setenv("NINT_ANALYTIC", "1")
    rng = np.random.default_rng(515)
    for kind, lens, dtype in [(FIXED, (40, 33, 26), np.float64), (FIXED, (57, 31), np.float64), (FIXED, (19, 22, 35), np.float32),
                              (FIXED, (64, 48), np.float32)]:
        axes = _axes(rng, lens, dtype, "analytic")
        values = rng.uniform(-1, 1, lens).astype(dtype)
        cols = _queries(rng, axes, 200_000, dtype, 0.1)
        for strat, extrap in [(LIN, e) for e in (ERROR, CLAMP, ENABLE, FILL, WRAP)] + [(NEA, e) for e in (ERROR, CLAMP, FILL, WRAP)]:
            g, o = make_pair(kind, axes, values, strat, extrap, fill=0.5, dtype=dtype)
            got, st, info = gpu_eval(g, cols)
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            assert_same(got, st, want, st_want, f"analytic {lens} {dtype.__name__} {strat} {extrap}")
            assert info == expected_info(st_want)
    # the same with an axis pack too large for shared memory
    for lens, dtype in [((6000, 5), np.float64), ((3000, 7, 3), np.float64), ((9000, 4), np.float32)]:
        axes = _axes(rng, lens, dtype, "analytic")
        values = rng.uniform(-1, 1, lens).astype(dtype)
        cols = _queries(rng, axes, 200_000, dtype, 0.1)
        for strat, extrap in [(LIN, CLAMP), (LIN, ERROR), (NEA, CLAMP), (NEA, WRAP)]:
            g, o = make_pair(FIXED, axes, values, strat, extrap, dtype=dtype)
            got, st, info = gpu_eval(g, cols)
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            assert_same(got, st, want, st_want, f"analytic, pack in global memory {lens} {dtype.__name__} {strat} {extrap}")
            assert info == expected_info(st_want)


def test_binned_tile_route(monkeypatch):
    # Incoherent batches on 3-D Linear tables beyond L2 are bucketed by grid tile and evaluated from TMA-staged tiles
    # in shared memory (nint_tile.cuh).  Forced here on small tables, with small superchunks and segments so that a
    # batch spans several of both, and with queries piled into one tile so that lists overflow.
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    monkeypatch.setenv("NINT_SLAB_MIN_BYTES", "0")
    monkeypatch.setenv("NINT_SLAB_MIN_POINTS", "0")
    monkeypatch.setenv("NINT_ROUTE", "tile")
    monkeypatch.setenv("NINT_PROBE_SHIFT", "0")
    monkeypatch.setenv("NINT_COMPACT", "0")  # 4-D / 5-D Fill and Error batches would otherwise take the in-range compaction pass
    rng = np.random.default_rng(4242)
    n = 300_001
    for kind, lens, dtype, mode, sc_shift, seg_shift in [(FIXED, (40, 9, 8), np.float64, "uniform", 16, 17),
                                                         (FIXED, (37, 35, 70), np.float64, "nonuniform", 15, 17),
                                                         (ND, (21, 40, 36), np.float32, "nonuniform", 17, 28),
                                                         (ND, (50, 18, 34), np.float64, "uniform", 23, 28),
                                                         (FIXED, (18, 67, 72), np.float32, "uniform", 14, 16),
                                                         (FIXED, (45, 33, 66), np.float64, "analytic", 15, 17),
                                                         (ND, (20, 35, 40), np.float32, "analytic", 16, 28),
                                                         (ND, (10, 9, 12, 8), np.float64, "nonuniform", 15, 17),
                                                         (ND, (6, 5, 7, 4, 8), np.float32, "nonuniform", 16, 18),
                                                         (ND, (5, 4, 6, 3, 6), np.float64, "uniform", 14, 28)]:
        monkeypatch.setenv("NINT_TILE_SC_SHIFT", str(sc_shift))
        monkeypatch.setenv("NINT_TILE_SEG_SHIFT", str(seg_shift))
        # distinct cells per window of 4096 consecutive points: ~4000 (or every cell) in random order, the cells of a
        # short run in cell order
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", str(min(int(np.prod([l - 1 for l in lens])) // 3, 2000)))
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        for extrap in (ERROR, CLAMP, ENABLE, FILL, WRAP):
            g, o = make_pair(kind, axes, values, LIN, extrap, fill=-2.5, dtype=dtype)
            cols = _queries(rng, axes, n, dtype, 0.05)
            for d in range(len(lens)):  # nodes (tile boundaries are nodes), their neighbours, non-finite values
                a = np.asarray(axes[d], dtype=dtype)
                special = np.concatenate([a, np.nextafter(a, dtype(np.inf)), np.nextafter(a, dtype(-np.inf)),
                                          np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
                where = rng.choice(n, special.size, replace=False)
                cols[d][where] = special
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            got, st, info = gpu_eval(g, cols)
            assert_same(got, st, want, st_want, f"tiles random {lens} {extrap}")
            assert info == expected_info(st_want)
            assert _lib.lib().nint_last_query_order(g._h) == 2
            assert _lib.lib().nint_last_incoherent_route(g._h) == 2
            got, st, _ = gpu_eval(g, cols, aos=True)
            assert_same(got, st, want, st_want, f"tiles aos {lens} {extrap}")
            # coherent order (sorted by cell): the probe sends it to the single pass
            order = np.lexsort([np.searchsorted(axes[2], cols[2]), np.searchsorted(axes[1], cols[1]),
                                np.searchsorted(axes[0], cols[0])])
            got, st, info = gpu_eval(g, [c[order] for c in cols])
            assert_same(got, st, want[order], st_want[order], f"tiles sorted {lens} {extrap}")
            assert _lib.lib().nint_last_query_order(g._h) == 1
        # all queries in one corner of the grid, probe forced to "incoherent": the lists of that tile overflow and the
        # bin kernel finishes the surplus itself
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", "0")
        g, o = make_pair(kind, axes, values, LIN, ERROR, dtype=dtype)
        cols = [rng.uniform(a[0], a[min(3, len(a) - 1)], n).astype(dtype) for a in axes]
        got, st, info = gpu_eval(g, cols)
        want, st_want = o.eval(cols, nthreads=O.max_threads())
        assert_same(got, st, want, st_want, f"tiles overflow {lens}")
        assert _lib.lib().nint_last_incoherent_route(g._h) == 2
    # a last axis whose byte length is not a multiple of 16 cannot be described to TMA: slab passes take over
    monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", "0")
    g, o = make_pair(FIXED, _axes(rng, (30, 8, 7), np.float64), rng.uniform(-1, 1, (30, 8, 7)), LIN, ERROR)
    cols = _queries(rng, g.data.grid, 50_000, np.float64, 0.05)
    got, st, _ = gpu_eval(g, cols)
    assert_same(got, st, *o.eval(cols), "odd last axis")
    assert _lib.lib().nint_last_incoherent_route(g._h) == 1


def test_slab_list_route(monkeypatch):
    # Incoherent batches on 3-D Linear tables far beyond L2 are split once, chunk by chunk, into lists by slab of the
    # cell-packed table, evaluated list by list and merged back into query order (nint_part.cuh).  Forced here on small
    # tables with a tiny slab size (many lists) and small segments, ragged last chunks included.
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    monkeypatch.setenv("NINT_SLAB_MIN_BYTES", "0")
    monkeypatch.setenv("NINT_SLAB_MIN_POINTS", "0")
    monkeypatch.setenv("NINT_ROUTE", "part")
    monkeypatch.setenv("NINT_PROBE_SHIFT", "0")
    rng = np.random.default_rng(777)
    for kind, lens, dtype, mode, slab_mb, seg_shift, n in [(FIXED, (40, 9, 8), np.float64, "uniform", 0.008, 14, 300_001),
                                                           (FIXED, (37, 35, 70), np.float64, "nonuniform", 0.3, 17, 300_001),
                                                           (ND, (21, 40, 36), np.float32, "nonuniform", 0.1, 28, 300_001),
                                                           (ND, (50, 18, 34), np.float64, "uniform", 0.3, 16, 123_457),
                                                           (FIXED, (18, 67, 72), np.float32, "uniform", 0.2, 28, 4_097),
                                                           (FIXED, (45, 33, 66), np.float64, "analytic", 0.8, 17, 300_001),
                                                           (FIXED, (3, 33, 66), np.float64, "uniform", 0.1, 15, 100_000)]:
        monkeypatch.setenv("NINT_PART_SLAB_MB", str(slab_mb))
        monkeypatch.setenv("NINT_PART_SEG_SHIFT", str(seg_shift))
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", str(min(int(np.prod([l - 1 for l in lens])) // 3, 2000)))
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        for extrap in (ERROR, CLAMP, ENABLE, FILL, WRAP):
            g, o = make_pair(kind, axes, values, LIN, extrap, fill=-2.5, dtype=dtype)
            cols = _queries(rng, axes, n, dtype, 0.05)
            for d in range(len(lens)):  # nodes (list boundaries are nodes), their neighbours, non-finite values
                a = np.asarray(axes[d], dtype=dtype)
                special = np.concatenate([a, np.nextafter(a, dtype(np.inf)), np.nextafter(a, dtype(-np.inf)),
                                          np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
                where = rng.choice(n, special.size, replace=False)
                cols[d][where] = special
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            got, st, info = gpu_eval(g, cols)
            assert_same(got, st, want, st_want, f"slab lists random {lens} {extrap}")
            assert info == expected_info(st_want)
            assert _lib.lib().nint_last_query_order(g._h) == 2
            assert _lib.lib().nint_last_incoherent_route(g._h) == 3
            got, st, _ = gpu_eval(g, cols, aos=True)
            assert_same(got, st, want, st_want, f"slab lists aos {lens} {extrap}")
            # coherent order (sorted by cell): the probe sends it to the single pass
            order = np.lexsort([np.searchsorted(axes[2], cols[2]), np.searchsorted(axes[1], cols[1]),
                                np.searchsorted(axes[0], cols[0])])
            got, st, info = gpu_eval(g, [c[order] for c in cols])
            assert_same(got, st, want[order], st_want[order], f"slab lists sorted {lens} {extrap}")
            if n >= 100_000:  # (fewer points than cells: consecutive queries do not share cells in any order)
                assert _lib.lib().nint_last_query_order(g._h) == 1
        # every query in the first cells of axis 0, probe forced to "incoherent": one list holds whole chunks
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", "0")
        g, o = make_pair(kind, axes, values, LIN, ERROR, dtype=dtype)
        cols = [rng.uniform(a[0], a[min(1, len(a) - 1)], n).astype(dtype) for a in axes]
        got, st, info = gpu_eval(g, cols)
        want, st_want = o.eval(cols, nthreads=O.max_threads())
        assert_same(got, st, want, st_want, f"slab lists, one list {lens}")
        assert _lib.lib().nint_last_incoherent_route(g._h) == 3


def test_slab_list_route_nd(monkeypatch):
    # The same route for InterpND with four and five axes (default there under Extrapolate::Wrap on tables beyond L2):
    # N columns, lists by slab of the flattened (axis 0, axis 1) cell index -- up to 128 of them --, cells found from
    # the axis records.  Forced on small tables; NINT_COMPACT=0 sends Fill and Error through it as well.
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    from ninterp_b200 import _lib

    monkeypatch.setenv("NINT_SLAB_MIN_BYTES", "0")
    monkeypatch.setenv("NINT_SLAB_MIN_POINTS", "0")
    monkeypatch.setenv("NINT_ROUTE", "part")
    monkeypatch.setenv("NINT_COMPACT", "0")
    monkeypatch.setenv("NINT_PROBE_SHIFT", "0")
    rng = np.random.default_rng(778)
    for lens, dtype, mode, slab_mb, seg_shift, n in [((7, 6, 9, 8), np.float64, "nonuniform", 0.02, 15, 200_001),
                                                     ((6, 5, 7, 6, 8), np.float32, "nonuniform", 0.05, 28, 200_001),
                                                     ((13, 12, 4, 4, 4), np.float32, "uniform", 0.007, 16, 150_000),
                                                     ((5, 6, 5, 6, 5), np.float64, "uniform", 0.1, 28, 4_097)]:
        monkeypatch.setenv("NINT_PART_SLAB_MB", str(slab_mb))
        monkeypatch.setenv("NINT_PART_SEG_SHIFT", str(seg_shift))
        monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", str(min(int(np.prod([l - 1 for l in lens])) // 3, 2000)))
        axes = _axes(rng, lens, dtype, mode)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        for extrap in (WRAP, CLAMP, ENABLE, ERROR, FILL):
            g, o = make_pair(ND, axes, values, LIN, extrap, fill=-2.5, dtype=dtype)
            cols = _queries(rng, axes, n, dtype, 0.3 if extrap == WRAP else 0.05)
            for d in range(len(lens)):
                a = np.asarray(axes[d], dtype=dtype)
                special = np.concatenate([a, np.nextafter(a, dtype(np.inf)), np.nextafter(a, dtype(-np.inf)),
                                          np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
                where = rng.choice(n, special.size, replace=False)
                cols[d][where] = special
            want, st_want = o.eval(cols, nthreads=O.max_threads())
            got, st, info = gpu_eval(g, cols)
            assert_same(got, st, want, st_want, f"slab lists N-D random {lens} {extrap}")
            assert info == expected_info(st_want)
            assert _lib.lib().nint_last_query_order(g._h) == 2
            assert _lib.lib().nint_last_incoherent_route(g._h) == 3
            got, st, _ = gpu_eval(g, cols, aos=True)
            assert_same(got, st, want, st_want, f"slab lists N-D aos {lens} {extrap}")
            order = np.lexsort([np.searchsorted(axes[d], cols[d]) for d in reversed(range(len(lens)))])
            got, st, info = gpu_eval(g, [c[order] for c in cols])
            assert_same(got, st, want[order], st_want[order], f"slab lists N-D sorted {lens} {extrap}")


def test_concurrent_batches_on_one_handle(monkeypatch):
    # include/ninterp_cuda.h: a handle is immutable during batch calls, so several host threads may evaluate
    # batches on their own streams at once -- on the single-pass route and on the probed one (order flag ring)
    import threading

    import torch

    from gpu_util import assert_same, make_pair

    monkeypatch.setenv("NINT_SLAB_MIN_BYTES", "0")
    monkeypatch.setenv("NINT_SLAB_MIN_POINTS", "0")
    rng = np.random.default_rng(2024)
    dev = torch.device("cuda", 0)
    monkeypatch.setenv("NINT_PROBE_SHIFT", "0")
    monkeypatch.setenv("NINT_PROBE_MAX_DISTINCT", "600")
    for passes in ("0", "3", "tile", "part"):
        if passes == "tile":  # binned route: every call draws its lists from the handle's pool in stream order
            monkeypatch.delenv("NINT_SLABS")
            monkeypatch.setenv("NINT_ROUTE", "tile")
            monkeypatch.setenv("NINT_TILE_SC_SHIFT", "16")
        elif passes == "part":  # slab lists: the same, and every call has its own work counter
            monkeypatch.setenv("NINT_ROUTE", "part")
            monkeypatch.setenv("NINT_PART_SLAB_MB", "0.008")
            monkeypatch.setenv("NINT_PART_SEG_SHIFT", "16")
        else:
            monkeypatch.setenv("NINT_SLABS", passes)
        axes = _axes(rng, (40, 9, 8), np.float64, "uniform")
        values = rng.uniform(-1, 1, (40, 9, 8))
        g, o = make_pair(FIXED, axes, values, LIN, ERROR, dtype=np.float64)
        jobs = []
        for t in range(6):
            cols = _queries(rng, axes, 150_000 + 1000 * t, np.float64, 0.05)
            if t % 2:  # every other batch is coherent: both kinds of launch are live at the same time
                order = np.argsort(cols[0], kind="stable")
                cols = [c[order] for c in cols]
            jobs.append((cols, o.eval(cols, nthreads=O.max_threads())))
        results = [None] * len(jobs)

        def work(t):
            s = torch.cuda.Stream(device=dev)
            cols = [torch.from_numpy(c).to(dev) for c in jobs[t][0]]
            torch.cuda.synchronize()
            outs = []
            for _ in range(5):
                out = torch.empty(cols[0].numel(), dtype=torch.float64, device=dev)
                st = torch.empty(cols[0].numel(), dtype=torch.uint8, device=dev)
                g.interpolate_batch(cols, out=out, status=st, stream=s.cuda_stream)
                outs.append((out, st))
            s.synchronize()
            results[t] = [(a.cpu().numpy(), b.cpu().numpy()) for a, b in outs]

        threads = [threading.Thread(target=work, args=(t,)) for t in range(len(jobs))]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        for t, (cols, (want, st_want)) in enumerate(jobs):
            assert results[t] is not None, f"thread {t} failed"
            for got, st in results[t]:
                assert_same(got, st, want, st_want, f"concurrent passes={passes} thread {t}")


def test_f32_axis_spanning_more_than_flt_max():
    # an f32 grid from -3e38 to 3e38: the span is finite in double but p - g[0] overflows in f32, so bucket and cell
    # formulas cannot be used; such an axis takes the general path (full binary search, IEEE division)
    from gpu_util import assert_same, expected_info, gpu_eval, make_pair

    g = np.array([-3e38, -1e38, 1e38, 2e38, 3e38], dtype=np.float32)
    v = np.array([1.0, -2.0, 4.0, 0.5, 3.0], dtype=np.float32)
    q = np.array([0.5e38, -2.9e38, 2.5e38, 1e38, -1e38, 3e38, -3e38, 0.0, 3.3e38, np.nan, -np.inf] + list(np.linspace(-3e38, 3e38, 4001)),
                 dtype=np.float32)
    for kind in (FIXED, ND):
        for strat in (LIN, NEA, LEFT, RIGHT) if kind == FIXED else (LIN, NEA):
            for extrap in (ERROR, CLAMP, FILL) + ((ENABLE,) if strat == LIN else ()):
                gi, o = make_pair(kind, [g], v, strat, extrap, fill=9.0, dtype=np.float32)
                got, st, info = gpu_eval(gi, [q])
                want, st_want = o.eval([q])
                assert_same(got, st, want, st_want, f"wide f32 axis kind={kind} strat={strat} extrap={extrap}")
                assert info == expected_info(st_want)
    # and as one axis of a 2-D interpolator next to an ordinary one
    g2 = np.linspace(0, 1, 7).astype(np.float32)
    v2 = np.random.default_rng(1).uniform(-1, 1, (5, 7)).astype(np.float32)
    cols = [q[:2000], np.random.default_rng(2).uniform(-0.1, 1.1, 2000).astype(np.float32)]
    for strat in (LIN, NEA):
        gi, o = make_pair(FIXED, [g, g2], v2, strat, CLAMP, dtype=np.float32)
        got, st, _ = gpu_eval(gi, cols)
        assert_same(got, st, *o.eval(cols), f"wide f32 axis in 2-D strat={strat}")


def test_result_buffers_are_checked_and_setters_keep_the_mirror_consistent():
    import torch

    import ninterp_b200 as nb

    g = nb.Interp2D([0., 1., 2.], [0., 1.], np.arange(6.).reshape(3, 2), nb.strategy.Linear, nb.Extrapolate.Error)
    dev = torch.device("cuda", g.device)
    cols = [torch.rand(100, dtype=torch.float64, device=dev) for _ in range(2)]
    for bad in (dict(out=torch.empty(100, dtype=torch.float32, device=dev)), dict(out=torch.empty(50, dtype=torch.float64, device=dev)),
                dict(out=torch.empty(100, dtype=torch.float64)), dict(status=torch.empty(100, dtype=torch.int32, device=dev)),
                dict(status=torch.empty(10, dtype=torch.uint8, device=dev)), dict(info=torch.zeros(4, dtype=torch.int64)),
                dict(out=torch.empty(200, dtype=torch.float64, device=dev)[::2])):
        with pytest.raises(nb.EngineError):
            g.interpolate_batch(cols, **bad)
    with pytest.raises(nb.EngineError):
        g.interpolate_batch_host(np.zeros((4, 2)), out=np.zeros(4, dtype=np.float32))
    # a strategy the interpolator cannot take leaves the mirror and the handle unchanged (LeftNearest is 1-D only)
    with pytest.raises(nb.EngineError):
        g.set_strategy(nb.strategy.LeftNearest)
    assert g.strategy == nb.strategy.Linear and g.interpolate([0.5, 0.5]) == 1.5
    # the reference assigns the strategy before checking it (two/mod.rs set_strategy): Nearest + Enable fails and stays
    g.set_extrapolate(nb.Extrapolate.Enable)
    with pytest.raises(nb.ValidateError):
        g.set_strategy(nb.strategy.Nearest)
    assert g.strategy == nb.strategy.Nearest
    # 0-D InterpND: set_extrapolate runs check_extrapolate too (n/mod.rs:342-346)
    z = nb.InterpND([[]], [3.25], nb.strategy.Nearest, nb.Extrapolate.Error)
    with pytest.raises(nb.ValidateError):
        z.set_extrapolate(nb.Extrapolate.Enable)
