"""Pure-Python model of ninterp 0.8.1's per-point semantics, written independently of oracle/.

TEST INFRASTRUCTURE ONLY.  It is deliberately *not* a transliteration of the Rust: it uses the
closed forms that the CUDA kernels use (count of nodes below the point, explicit panic
classification) so that agreement between this file and oracle/ninterp_oracle.hpp (which follows
the Rust line by line) on random and adversarial inputs checks both the oracle and the closed forms
of SURVEY.md §8a.  Slow (Python loops): small cases only.

All arithmetic is done on numpy scalars of the requested dtype so every operation rounds once in T.
"""
from __future__ import annotations

import math

import numpy as np

LINEAR, NEAREST, LEFT_NEAREST, RIGHT_NEAREST = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4
FIXED, ND = 0, 1
ST_OK, ST_ERR, ST_PANIC = 0, 1, 0xFF


class _Panic(Exception):
    pass


def count_less(g, p):
    """#{i : g[i] < p} over the whole axis (NaN -> 0)."""
    return int(np.sum(g < p))


def bracket_raw(g, p):
    """find_nearest_index (strategy/traits.rs:10-33) in closed form."""
    L = len(g)
    if p == g[L - 1]:
        if L < 2:
            raise _Panic
        return L - 2
    if p != p:  # NaN: every `>=` is false, the search runs off to the right
        return L - 1
    c = count_less(g, p)
    if c >= L:  # beyond the last node: lower bound saturates at L-1 and is not decremented
        return L - 1
    return max(c, 1) - 1


def linear_lower(g, p):
    L = len(g)
    if p < g[0]:
        return 0
    if p > g[L - 1]:
        if L < 2:
            raise _Panic
        return L - 2
    return bracket_raw(g, p)


def _node(g, i):
    if i < 0 or i >= len(g):
        raise _Panic
    return g[i]


def first_equal(g, p):
    """Smallest i with g[i] == p, or None (sorted axis: it is count_less when that node equals p)."""
    c = count_less(g, p)
    if c < len(g) and g[c] == p:
        return c
    return None


def wrap(x, lo, hi):
    T = type(x)
    a = x - lo
    b = hi - lo
    with np.errstate(all="ignore"):
        r = T(np.fmod(a, b))
    if r < 0:
        r = r + abs(b)
    return lo + r


def clamp(x, lo, hi):
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


def _lerp_axis(T, g, p):
    l = linear_lower(g, p)
    gl, gu = _node(g, l), _node(g, l + 1)
    with np.errstate(all="ignore"):
        t = (p - gl) / (gu - gl)
    return l, t


def _nearest_axis(g, p):
    l = bracket_raw(g, p)
    gl, gu = _node(g, l), _node(g, l + 1)
    with np.errstate(all="ignore"):
        return l if (p - gl) < (gu - p) else l + 1


def _strategy(kind, axes, values, strategy, q, T):
    n = len(axes)
    one = T(1)
    if kind == FIXED and n == 1 or kind == ND:
        # exact-node short-cuts: one/strategies.rs:14 (1-D), n/strategies.rs:36-46 (N-D, per axis)
        fixed = [first_equal(axes[d], q[d]) for d in range(n)]
    else:
        fixed = [None] * n
    active = [d for d in range(n) if fixed[d] is None]
    if kind == ND:
        rem = 1
        for d in active:
            rem *= len(axes[d])
        if rem == 1:  # n/strategies.rs:47-50
            idx = tuple(fixed[d] if fixed[d] is not None else 0 for d in range(n))
            return values[idx]
    elif n == 1 and fixed[0] is not None:
        return values[fixed[0]]

    if strategy == LINEAR:
        lows, ts = {}, {}
        for d in active:
            lows[d], ts[d] = _lerp_axis(T, axes[d], q[d])
        for d in active:
            if lows[d] + 1 >= values.shape[d]:
                raise _Panic
        # gather the cube over active axes
        m = len(active)
        cube = np.empty((2,) * m, dtype=values.dtype)
        for c in range(1 << m):
            bits = [(c >> (m - 1 - k)) & 1 for k in range(m)]
            idx = [fixed[d] if fixed[d] is not None else 0 for d in range(n)]
            for k, d in enumerate(active):
                idx[d] = lows[d] + bits[k]
            cube[tuple(bits)] = values[tuple(idx)]
        with np.errstate(all="ignore"):
            for d in active:  # lowest axis first
                t = ts[d]
                s = one - t
                cube = cube[0] * s + cube[1] * t
        return T(cube)
    if strategy == NEAREST:
        idx = [fixed[d] if fixed[d] is not None else 0 for d in range(n)]
        for d in active:
            idx[d] = _nearest_axis(axes[d], q[d])
        for d in range(n):
            if idx[d] >= values.shape[d]:
                raise _Panic
        return values[tuple(idx)]
    # Left/RightNearest: 1-D only
    g = axes[0]
    l = bracket_raw(g, q[0])
    i = l if strategy == LEFT_NEAREST else l + 1
    if i >= len(values):
        raise _Panic
    return values[i]


def interpolate(kind, axes, values, strategy, extrap, fill, point, dtype=np.float64):
    """-> (value, status).  axes: list of 1-D arrays; values: N-D array; point: sequence."""
    T = np.dtype(dtype).type
    axes = [np.asarray(a, dtype=dtype) for a in axes]
    values = np.asarray(values, dtype=dtype)
    nan = T(math.nan)
    if kind == ND and values.size == 1:
        return values.reshape(-1)[0], ST_OK
    p = [T(v) for v in point]
    n = len(axes)
    oob = [not (axes[d][0] <= p[d] <= axes[d][-1]) for d in range(n)]
    q = p
    if any(oob):
        if extrap == FILL:
            return T(fill), ST_OK
        if extrap == ERROR:
            return nan, ST_ERR
        if extrap == CLAMP:
            q = [clamp(p[d], axes[d][0], axes[d][-1]) for d in range(n)]
        elif extrap == WRAP:
            with np.errstate(all="ignore"):
                q = [wrap(p[d], axes[d][0], axes[d][-1]) for d in range(n)]
    try:
        return T(_strategy(kind, axes, values, strategy, q, T)), ST_OK
    except _Panic:
        return nan, ST_PANIC
