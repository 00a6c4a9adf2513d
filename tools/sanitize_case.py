#!/usr/bin/env python
"""A reduced parity run for compute-sanitizer: every route and the loads that speculate.

usage: compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_case.py
Covers: single pass on the cell-packed and on the plain table (slack-pair load `ld_pair`, the speculative gathers of
points that fail verification), slab passes (warp-autonomous shared lists), binned tiles (TMA, mbarrier, atomics,
list overflow), length-1 and length-2 axes, an axis pack too large for shared memory, 1-D/2-D/N-D, f32 and f64,
AoS input, the host pipeline.  Every result is compared with the CPU oracle.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_binding as O  # noqa: E402
from gpu_util import assert_same, expected_info, gpu_eval, make_pair  # noqa: E402

FIXED, ND = 0, 1
LIN, NEA, LEFT, RIGHT = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4


def axes_of(rng, lens, dtype, uniform=False):
    if uniform == "analytic":  # nodes are exactly k * h: eligible for the kernels that compute them
        return [np.arange(L, dtype=dtype) * dtype(0.37) for L in lens]
    return [((np.arange(L) * 0.37 - 1.3) if uniform else (np.cumsum(rng.uniform(0.5, 1.5, L)) + rng.uniform(-3, 3))).astype(dtype) for L in lens]


def queries(rng, axes, n, dtype):
    cols = []
    for a in axes:
        lo, hi = float(a[0]), float(a[-1])
        span = max(hi - lo, 1.0)
        c = rng.uniform(lo - 0.05 * span, hi + 0.05 * span, n).astype(dtype)
        k = min(len(a), n // 8)
        c[rng.choice(n, k, replace=False)] = a[:k]                      # exact node hits
        c[rng.choice(n, 3, replace=False)] = np.array([np.nan, np.inf, -np.inf], dtype=dtype)
        cols.append(c)
    return cols


def run(label, kind, lens, dtype, strat, extrap, n, rng, uniform=False, env=None, aos=False):
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        axes = axes_of(rng, lens, dtype, uniform)
        values = rng.uniform(-1, 1, lens).astype(dtype)
        g, o = make_pair(kind, axes, values, strat, extrap, fill=-2.5, dtype=dtype)
        cols = queries(rng, axes, n, dtype)
        got, st, info = gpu_eval(g, cols, aos=aos)
        want, st_want = o.eval(cols, nthreads=O.max_threads())
        assert_same(got, st, want, st_want, label)
        assert info == expected_info(st_want), label
        if kind == FIXED and len(lens) == 3 and not aos:
            pts = np.stack(cols, axis=1)
            h, hs, _ = g.interpolate_batch_host(pts, return_status=True)  # host pipeline: three streams
            assert_same(h, hs, want, st_want, label + " host")
        g.close()
        print("ok", label, flush=True)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    rng = np.random.default_rng(5)
    n = 40_000
    forced = {"NINT_SLAB_MIN_BYTES": "0", "NINT_SLAB_MIN_POINTS": "0", "NINT_PROBE_MAX_DISTINCT": "0"}
    for extrap in (ERROR, CLAMP, ENABLE, WRAP):
        run(f"3d packed {extrap}", FIXED, (15, 16, 18), np.float64, LIN, extrap, n, rng)
        run(f"3d plain {extrap}", FIXED, (15, 16, 18), np.float64, LIN, extrap, n, rng, env={"NINT_PACK": "0"})
    run("3d nearest", FIXED, (9, 8, 7), np.float32, NEA, CLAMP, n, rng)
    run("slab passes", FIXED, (40, 9, 7), np.float64, LIN, ERROR, n, rng, uniform=True, env=dict(forced, NINT_SLABS="3"))
    run("slab passes nd f32", ND, (17, 5, 4, 3), np.float32, LIN, CLAMP, n, rng, env=dict(forced, NINT_SLABS="2"))
    tile = dict(forced, NINT_ROUTE="tile", NINT_TILE_SC_SHIFT="13", NINT_TILE_SEG_SHIFT="14")
    run("binned tiles", FIXED, (37, 35, 70), np.float64, LIN, ERROR, n, rng, env=tile)
    run("binned tiles wrap aos", FIXED, (40, 9, 8), np.float64, LIN, WRAP, n, rng, uniform="analytic", env=tile, aos=True)
    run("binned tiles f32 nd", ND, (21, 40, 36), np.float32, LIN, FILL, n, rng, env=tile)
    run("binned tiles 5-d f32 wrap", ND, (6, 5, 7, 4, 8), np.float32, LIN, WRAP, n, rng, env=tile)
    run("binned tiles 4-d f64", ND, (10, 9, 12, 8), np.float64, LIN, CLAMP, n, rng, env=tile)
    part = dict(forced, NINT_ROUTE="part", NINT_PART_SLAB_MB="0.05", NINT_PART_SEG_SHIFT="14")
    run("slab lists", FIXED, (37, 35, 70), np.float64, LIN, ERROR, n, rng, env=part)
    run("slab lists wrap aos", FIXED, (40, 9, 8), np.float64, LIN, WRAP, n, rng, uniform="analytic", env=dict(part, NINT_PART_SLAB_MB="0.008"), aos=True)
    run("slab lists f32 nd", ND, (21, 40, 36), np.float32, LIN, FILL, 50_001, rng, env=part)
    run("slab lists 5-d f32 wrap", ND, (6, 5, 7, 4, 8), np.float32, LIN, WRAP, 50_001, rng, env=dict(part, NINT_PART_SLAB_MB="0.02"))
    run("slab lists 4-d f64 aos", ND, (10, 9, 12, 8), np.float64, LIN, CLAMP, n, rng, env=part, aos=True)
    run("length-1 axes nd", ND, (1, 6, 1, 5), np.float64, LIN, CLAMP, n, rng)
    run("length-2 axes", FIXED, (2, 2, 2), np.float64, LIN, ENABLE, n, rng)
    run("1d left", FIXED, (33,), np.float64, LEFT, ERROR, n, rng)
    run("1d right", FIXED, (33,), np.float32, RIGHT, FILL, n, rng)
    run("2d linear", FIXED, (21, 18), np.float32, LIN, WRAP, n, rng)
    run("axis pack beyond shared memory", FIXED, (6000, 5), np.float64, LIN, CLAMP, n, rng)
    run("nd5 f32", ND, (6, 5, 7, 4, 5), np.float32, LIN, WRAP, n, rng)
    run("nd5 f32 in-range compaction", ND, (6, 5, 7, 4, 5), np.float32, LIN, FILL, 100_000, rng)
    run("nd4 f64 in-range compaction", ND, (6, 5, 7, 4), np.float64, LIN, ERROR, 100_000, rng)
    run("1d fused", FIXED, (300,), np.float64, LIN, CLAMP, n, rng)
    run("analytic", FIXED, (40, 33, 26), np.float64, LIN, CLAMP, n, rng, uniform="analytic", env={"NINT_ANALYTIC": "1"})
    print("sanitize_case: all cases bit-exact against the oracle")


if __name__ == "__main__":
    main()
