"""Golden vectors transcribed from the reference's own tests, doctests and examples.

Every entry cites the reference file:line (relative to /root/reference/) that holds the assertion.
`exact` entries come from `assert_eq!` on f64 and are bit-exact targets; `approx` entries come from
`assert_approx_eq!` (abs diff <= 1e-6, src/lib.rs:66-75).  The same table drives
tests/test_oracle_kat.py (CPU oracle), tests/test_pyref.py (pure-Python model) and
tests/test_gpu_kat.py (CUDA path through the C ABI).

Check kinds:
  ("exact",  point, value)       result == value bit-for-bit, status OK
  ("approx", point, value)       |result - value| <= 1e-6, status OK
  ("nan",    point)              result is NaN with status OK (Extrapolate::Fill(NAN))
  ("error",  point)              InterpolateError::ExtrapolateError
  ("same",   point_a, point_b)   interpolate(a) == interpolate(b) bit-for-bit
  ("nodes",)                     interpolate(node) == values[node] for every grid node
"""
import numpy as np

LINEAR, NEAREST, LEFT_NEAREST, RIGHT_NEAREST = 0, 1, 2, 3
ENABLE, FILL, CLAMP, WRAP, ERROR = 0, 1, 2, 3, 4
FIXED, ND = 0, 1
NAN = float("nan")

X5 = [0., 1., 2., 3., 4.]
F5 = [0.2, 0.4, 0.6, 0.8, 1.0]

X3 = [0.05, 0.10, 0.15]
Y3 = [0.10, 0.20, 0.30]
Z3 = [0.20, 0.40, 0.60]
F33 = [[0., 1., 2.], [3., 4., 5.], [6., 7., 8.]]
F333 = [
    [[0., 1., 2.], [3., 4., 5.], [6., 7., 8.]],
    [[9., 10., 11.], [12., 13., 14.], [15., 16., 17.]],
    [[18., 19., 20.], [21., 22., 23.], [24., 25., 26.]],
]
U2 = [0., 1.]
F22 = [[0., 1.], [2., 3.]]
F222 = [[[0., 1.], [2., 3.]], [[4., 5.], [6., 7.]]]
OX, OY, OZ = [0.1, 1.1], [0.2, 1.2], [0.3, 1.3]

# doc examples three/mod.rs:98-121, n/mod.rs:191-217: f = 0.2x + 0.2y + 0.2z
DOC3_AXES = [[1., 2.], [1., 2., 3.], [1., 2., 3., 4.]]
DOC3_VALUES = [
    [[0.6, 0.8, 1.0, 1.2], [0.8, 1.0, 1.2, 1.4], [1.0, 1.2, 1.4, 1.6]],
    [[0.8, 1.0, 1.2, 1.4], [1.0, 1.2, 1.4, 1.6], [1.2, 1.4, 1.6, 1.8]],
]
# two/mod.rs:96-111: f = 0.2x + 0.4y
DOC2_AXES = [[0., 1., 2.], [0., 1., 2., 3.]]
DOC2_VALUES = [[0.0, 0.4, 0.8, 1.2], [0.2, 0.6, 1.0, 1.4], [0.4, 0.8, 1.2, 1.6]]

EXTRAP3_PTS = [  # three/tests.rs:56-78 (value), n/tests.rs:141-212 (N-D == 3-D)
    ([0.01, 0.06, 0.17], -8.55), ([0.02, 0.08, 0.19], -6.05),
    ([0.01, 0.06, 0.63], -6.25), ([0.02, 0.08, 0.65], -3.75),
    ([0.01, 0.33, 0.17], -0.45), ([0.02, 0.36, 0.19], 2.35),
    ([0.01, 0.33, 0.63], 1.85), ([0.02, 0.36, 0.65], 4.65),
    ([0.17, 0.06, 0.17], 20.25), ([0.19, 0.08, 0.19], 24.55),
    ([0.17, 0.06, 0.63], 22.55), ([0.19, 0.08, 0.65], 26.85),
    ([0.17, 0.33, 0.17], 28.35), ([0.19, 0.36, 0.19], 32.95),
    ([0.17, 0.33, 0.63], 30.65), ([0.19, 0.36, 0.65], 35.25),
]
EXTRAP2_PTS = [  # two/tests.rs:52-62 (value), n/tests.rs:72-107 (N-D == 2-D)
    ([0.0, 0.0], -4.), ([0.03, 0.04], -1.8), ([0.0, 0.32], -0.8), ([0.03, 0.36], 1.4),
    ([0.17, 0.0], 6.2), ([0.19, 0.04], 7.8), ([0.17, 0.32], 9.4), ([0.19, 0.36], 11.),
]
LINEAR3_APPROX = [  # three/tests.rs:32-37, n/tests.rs:35-40, strategy/enums/mod.rs:132-137,172-177
    ([0.05, 0.10, 0.3], 0.5), ([0.05, 0.15, 0.20], 1.5), ([0.05, 0.15, 0.3], 2.),
    ([0.075, 0.10, 0.20], 4.5), ([0.075, 0.10, 0.3], 5.), ([0.075, 0.15, 0.20], 6.),
]
NEAREST3_ENUM = [  # strategy/enums/mod.rs:140-147, 180-187
    ([0.06, 0.11, 0.22], 0.), ([0.06, 0.11, 0.31], 1.), ([0.06, 0.19, 0.22], 3.), ([0.06, 0.19, 0.31], 4.),
    ([0.09, 0.11, 0.22], 9.), ([0.09, 0.11, 0.31], 10.), ([0.09, 0.19, 0.22], 12.), ([0.09, 0.19, 0.31], 13.),
]
OCTANTS = [[a, b, c] for a in (0., 2.) for b in (0., 2.) for c in (0., 2.)]


def case(name, ref, ndim_kind, axes, values, strategy, extrap, checks, fill=None):
    return dict(name=name, ref=ref, kind=ndim_kind, axes=axes, values=values, strategy=strategy,
                extrap=extrap, fill=fill, checks=checks)


KATS = [
    # ----------------------------------------------------------------------------- 1-D
    case("1d_invalid_args_value", "src/interpolator/one/tests.rs:16", FIXED, [X5], F5, LINEAR, ERROR,
         [("exact", [1.0], 0.4)]),
    case("1d_linear", "src/interpolator/one/tests.rs:20-36", FIXED, [X5], F5, LINEAR, ERROR,
         [("nodes",), ("exact", [3.00], 0.8), ("exact", [3.75], 0.95), ("exact", [4.00], 1.0)]),
    case("1d_left_nearest", "src/interpolator/one/tests.rs:39-56", FIXED, [X5], F5, LEFT_NEAREST, ERROR,
         [("nodes",), ("exact", [3.00], 0.8), ("exact", [3.75], 0.8), ("exact", [4.00], 1.0)]),
    case("1d_right_nearest", "src/interpolator/one/tests.rs:59-76", FIXED, [X5], F5, RIGHT_NEAREST, ERROR,
         [("nodes",), ("exact", [3.00], 0.8), ("exact", [3.25], 1.0), ("exact", [4.00], 1.0)]),
    case("1d_nearest", "src/interpolator/one/tests.rs:79-98", FIXED, [X5], F5, NEAREST, ERROR,
         [("nodes",), ("exact", [3.00], 0.8), ("exact", [3.25], 0.8), ("exact", [3.50], 1.0),
          ("exact", [3.75], 1.0), ("exact", [4.00], 1.0)]),
    case("1d_extrapolate_error", "src/interpolator/one/tests.rs:114-131", FIXED, [X5], F5, LINEAR, ERROR,
         [("error", [-1.]), ("error", [5.])]),
    case("1d_extrapolate_fill", "src/interpolator/one/tests.rs:134-147", FIXED, [X5], F5, LINEAR, FILL,
         [("exact", [1.5], 0.5), ("exact", [2.], 0.6), ("nan", [-1.]), ("nan", [5.])], fill=NAN),
    case("1d_extrapolate_clamp", "src/interpolator/one/tests.rs:150-160", FIXED, [X5], F5, LINEAR, CLAMP,
         [("exact", [-1.], 0.2), ("exact", [5.], 1.0)]),
    case("1d_extrapolate_enable", "src/interpolator/one/tests.rs:163-174", FIXED, [X5], F5, LINEAR, ENABLE,
         [("exact", [-1.], 0.0), ("approx", [-0.75], 0.05), ("exact", [5.], 1.2)]),
    case("1d_doc", "src/interpolator/one/mod.rs:94-108", FIXED, [[0., 1., 2.]], [0.0, 0.4, 0.8], LINEAR, ENABLE,
         [("exact", [1.4], 0.56), ("exact", [3.6], 1.44)]),
    case("1d_enum_linear", "src/strategy/enums/mod.rs:57-67", FIXED, [X5], F5, LINEAR, ERROR,
         [("exact", [3.00], 0.8), ("exact", [3.75], 0.95), ("exact", [4.00], 1.0)]),
    case("1d_enum_nearest", "src/strategy/enums/mod.rs:69-74", FIXED, [X5], F5, NEAREST, ERROR,
         [("exact", [3.00], 0.8), ("exact", [3.25], 0.8), ("exact", [3.50], 1.0), ("exact", [3.75], 1.0),
          ("exact", [4.00], 1.0)]),
    case("1d_enum_left", "src/strategy/enums/mod.rs:76-79", FIXED, [X5], F5, LEFT_NEAREST, ERROR,
         [("exact", [3.00], 0.8), ("exact", [3.75], 0.8), ("exact", [4.00], 1.0)]),
    case("1d_example_dynamic_strategy_linear", "examples/dynamic_strategy.rs:16-24", FIXED,
         [[0., 1., 2.]], [0., 3., 6.], LINEAR, ERROR, [("exact", [1.75], 5.25)]),
    case("1d_example_dynamic_strategy_nearest", "examples/dynamic_strategy.rs:26-27", FIXED,
         [[0., 1., 2.]], [0., 3., 6.], NEAREST, ERROR, [("exact", [1.75], 6.)]),
    case("1d_example_dynamic_interpolator_nearest", "examples/dynamic_interpolator.rs:27-35", FIXED,
         [[0., 1., 2.]], [0., 4., 8.], NEAREST, ERROR, [("exact", [1.75], 8.)]),
    case("1d_example_uom", "examples/uom.rs:11-28", FIXED, [[0., 1.]], [250., 750.], LINEAR, ERROR,
         [("exact", [0.5], 500.)]),
    # ----------------------------------------------------------------------------- 2-D
    case("2d_linear", "src/interpolator/two/tests.rs:4-24", FIXED, [X3, Y3], F33, LINEAR, ERROR,
         [("nodes",), ("exact", [0.15, 0.20], 7.), ("exact", [0.075, 0.25], 3.)]),
    case("2d_linear_offset", "src/interpolator/two/tests.rs:27-37", FIXED, [U2, U2], F22, LINEAR, ERROR,
         [("approx", [0.25, 0.65], 1.15)]),
    case("2d_linear_extrapolation", "src/interpolator/two/tests.rs:40-63", FIXED, [X3, Y3], F33, LINEAR, ENABLE,
         [("approx", p, v) for p, v in EXTRAP2_PTS]),
    case("2d_nearest", "src/interpolator/two/tests.rs:66-94", FIXED, [X3, Y3], F33, NEAREST, ERROR,
         [("nodes",), ("exact", [0.05, 0.12], 0.), ("exact", [0.07, 0.15 + 0.0001], 1.),
          ("exact", [0.08, 0.21], 4.), ("exact", [0.11, 0.26], 5.), ("exact", [0.13, 0.12], 6.),
          ("exact", [0.14, 0.29], 8.)]),
    case("2d_extrapolate_error", "src/interpolator/two/tests.rs:110-126", FIXED, [OX, OY], F22, LINEAR, ERROR,
         [("error", [-1., -1.]), ("error", [2., 2.])]),
    case("2d_extrapolate_fill", "src/interpolator/two/tests.rs:129-145", FIXED, [OX, OY], F22, LINEAR, FILL,
         [("exact", [0.5, 0.5], 1.1), ("exact", [0.1, 1.2], 1.), ("nan", [0., 0.]), ("nan", [0., 2.]),
          ("nan", [2., 0.]), ("nan", [2., 2.])], fill=NAN),
    case("2d_dyn_strategy_linear", "src/interpolator/two/tests.rs:148-158", FIXED, [U2, U2], F22, LINEAR, ERROR,
         [("exact", [0.2, 0.], 0.4)]),
    case("2d_dyn_strategy_nearest", "src/interpolator/two/tests.rs:159-160", FIXED, [U2, U2], F22, NEAREST, ERROR,
         [("exact", [0.2, 0.], 0.)]),
    case("2d_extrapolate_clamp", "src/interpolator/two/tests.rs:163-175", FIXED, [OX, OY], F22, LINEAR, CLAMP,
         [("exact", [-1., -1.], 0.), ("exact", [2., 2.], 3.)]),
    case("2d_doc", "src/interpolator/two/mod.rs:96-117", FIXED, DOC2_AXES, DOC2_VALUES, LINEAR, CLAMP,
         [("exact", [1.5, 1.5], 0.9), ("same", [-1., 3.5], [0., 3.])]),
    case("2d_enum_doc_nearest", "src/interpolator/enums.rs:37-57", FIXED, [X3, Y3], F33, NEAREST, ERROR,
         [("exact", [0.08, 0.21], 4.), ("exact", [0.11, 0.26], 5.), ("exact", [0.13, 0.12], 6.),
          ("exact", [0.14, 0.29], 8.)]),
    case("2d_example_dynamic_interpolator", "examples/dynamic_interpolator.rs:16-24,56-66", FIXED,
         [[0., 1.], [0., 1., 2.]], [[2., 4., 6.], [4., 16., 32.]], LINEAR, ENABLE,
         [("exact", [1.5, -0.5], -3.5)]),
    # ----------------------------------------------------------------------------- 3-D
    case("3d_linear", "src/interpolator/three/tests.rs:4-38", FIXED, [X3, Y3, Z3], F333, LINEAR, ERROR,
         [("nodes",)] + [("approx", p, v) for p, v in LINEAR3_APPROX]),
    case("3d_linear_extrapolation", "src/interpolator/three/tests.rs:41-79", FIXED, [X3, Y3, Z3], F333, LINEAR,
         ENABLE, [("approx", p, v) for p, v in EXTRAP3_PTS]),
    case("3d_linear_offset", "src/interpolator/three/tests.rs:82-93", FIXED, [U2, U2, U2], F222, LINEAR, ERROR,
         [("approx", [0.25, 0.65, 0.9], 3.2)]),
    case("3d_nearest", "src/interpolator/three/tests.rs:96-127", FIXED, [U2, U2, U2], F222, NEAREST, ERROR,
         [("nodes",), ("exact", [0., 0., 0.], 0.), ("exact", [0.25, 0.25, 0.25], 0.),
          ("exact", [0.25, 0.75, 0.25], 2.), ("exact", [0., 1., 0.], 2.), ("exact", [0.75, 0.25, 0.75], 5.),
          ("exact", [0.75, 0.75, 0.75], 7.), ("exact", [1., 1., 1.], 7.)]),
    case("3d_extrapolate_error", "src/interpolator/three/tests.rs:144-162", FIXED, [OX, OY, OZ], F222, LINEAR, ERROR,
         [("error", [-1., -1., -1.]), ("error", [2., 2., 2.])]),
    case("3d_extrapolate_fill", "src/interpolator/three/tests.rs:165-185", FIXED, [OX, OY, OZ], F222, LINEAR, FILL,
         [("approx", [0.4, 0.4, 0.4], 1.7), ("approx", [0.8, 0.8, 0.8], 4.5)] + [("nan", p) for p in OCTANTS],
         fill=NAN),
    case("3d_extrapolate_clamp", "src/interpolator/three/tests.rs:188-200", FIXED, [OX, OY, OZ], F222, LINEAR, CLAMP,
         [("exact", [-1., -1., -1.], 0.), ("exact", [2., 2., 2.], 7.)]),
    case("3d_doc", "src/interpolator/three/mod.rs:98-127", FIXED, DOC3_AXES, DOC3_VALUES, LINEAR, ERROR,
         [("exact", [1.5, 1.5, 1.5], 0.9), ("error", [5.5, 5.5, 5.5])]),
    case("3d_enum_linear", "src/strategy/enums/mod.rs:112-137", FIXED, [X3, Y3, Z3], F333, LINEAR, ERROR,
         [("approx", p, v) for p, v in LINEAR3_APPROX]),
    case("3d_enum_nearest", "src/strategy/enums/mod.rs:139-147", FIXED, [X3, Y3, Z3], F333, NEAREST, ERROR,
         [("exact", p, v) for p, v in NEAREST3_ENUM]),
    case("3d_example_dynamic_interpolator_nearest", "examples/dynamic_interpolator.rs:38-47", FIXED,
         [U2, U2, U2], [[[0., 1.], [0.1, 1.1]], [[0.2, 1.2], [0.3, 1.3]]], NEAREST, ERROR,
         [("exact", [0.8, 0.7, 0.6], 1.3)]),
    # ----------------------------------------------------------------------------- N-D
    case("nd_linear", "src/interpolator/n/tests.rs:4-41", ND, [X3, Y3, Z3], F333, LINEAR, ERROR,
         [("nodes",)] + [("approx", p, v) for p, v in LINEAR3_APPROX]),
    case("nd_linear_offset", "src/interpolator/n/tests.rs:44-53", ND, [U2, U2, U2], F222, LINEAR, ERROR,
         [("approx", [0.25, 0.65, 0.9], 3.2)]),
    case("nd_linear_extrapolation_2d_values", "src/interpolator/n/tests.rs:56-107 + two/tests.rs:52-62", ND,
         [X3, Y3], F33, LINEAR, ENABLE, [("approx", p, v) for p, v in EXTRAP2_PTS]),
    case("nd_linear_extrapolate_3d_values", "src/interpolator/n/tests.rs:110-212 + three/tests.rs:56-78", ND,
         [X3, Y3, Z3], F333, LINEAR, ENABLE, [("approx", p, v) for p, v in EXTRAP3_PTS]),
    case("nd_nearest", "src/interpolator/n/tests.rs:215-242", ND, [U2, U2, U2], F222, NEAREST, ERROR,
         [("nodes",), ("exact", [0.25, 0.25, 0.25], 0.), ("exact", [0.25, 0.75, 0.25], 2.),
          ("exact", [0.75, 0.25, 0.75], 5.), ("exact", [0.75, 0.75, 0.75], 7.)]),
    case("nd_extrapolate_error", "src/interpolator/n/tests.rs:258-273", ND, [U2, U2, U2], F222, LINEAR, ERROR,
         [("error", [-1., -1., -1.]), ("error", [2., 2., 2.])]),
    case("nd_extrapolate_fill", "src/interpolator/n/tests.rs:276-292", ND, [OX, OY, OZ], F222, LINEAR, FILL,
         [("nan", p) for p in OCTANTS], fill=NAN),
    case("nd_extrapolate_clamp", "src/interpolator/n/tests.rs:295-319", ND, [OX, OY, OZ], F222, LINEAR, CLAMP,
         [("exact", [-1., -1., -1.], 0.), ("exact", [-1., 2., -1.], 2.), ("exact", [2., -1., 2.], 5.),
          ("exact", [2., 2., 2.], 7.)]),
    case("nd_extrapolate_wrap", "src/interpolator/n/tests.rs:322-346", ND, [U2, U2, U2], F222, LINEAR, WRAP,
         [("same", [-0.25, -0.2, -0.4], [0.75, 0.8, 0.6]), ("same", [-0.25, 2.1, -0.4], [0.75, 0.1, 0.6]),
          ("same", [-0.25, 2.1, 2.3], [0.75, 0.1, 0.3]), ("same", [2.5, 2.1, 2.3], [0.5, 0.1, 0.3])]),
    case("nd_doc", "src/interpolator/n/mod.rs:191-223", ND, DOC3_AXES, DOC3_VALUES, LINEAR, ERROR,
         [("exact", [1.5, 1.5, 1.5], 0.9), ("error", [5.5, 5.5, 5.5])]),
    case("nd_enum_linear", "src/strategy/enums/mod.rs:150-177", ND, [X3, Y3, Z3], F333, LINEAR, ERROR,
         [("approx", p, v) for p, v in LINEAR3_APPROX]),
    case("nd_enum_nearest", "src/strategy/enums/mod.rs:179-187", ND, [X3, Y3, Z3], F333, NEAREST, ERROR,
         [("exact", p, v) for p, v in NEAREST3_ENUM]),
]

# N-D must equal the hard-coded interpolators bit-for-bit (n/tests.rs:72-107 and :141-212)
ND_EQUALS_FIXED = [
    dict(name="nd_eq_2d", ref="src/interpolator/n/tests.rs:56-107", axes=[X3, Y3], values=F33,
         strategy=LINEAR, extrap=ENABLE, points=[p for p, _ in EXTRAP2_PTS]),
    dict(name="nd_eq_3d", ref="src/interpolator/n/tests.rs:110-212", axes=[X3, Y3, Z3], values=F333,
         strategy=LINEAR, extrap=ENABLE, points=[p for p, _ in EXTRAP3_PTS]),
]

# src/lib.rs:89-105 (integer cases evaluated in f64: same values, exactly representable)
WRAP_KATS = [
    (-3., -2., 5., 4.), (3., -2., 5., 3.), (6., -2., 5., -1.), (5., 0., 10., 5.), (11., 0., 10., 1.),
    (-3., 0., 10., 7.), (-11., 0., 10., 9.), (-0.1, -2., -1., -1.1), (-0., -2., -1., -2.0),
    (0.1, -2., -1., -1.9), (-0.5, -1., 1., -0.5), (0., -1., 1., 0.), (0.5, -1., 1., 0.5), (0.8, -1., 1., 0.8),
]

# Construction-time validation (ValidateError variants, src/error.rs:9-27).  Codes as in
# include/ninterp_cuda.h: -1 ExtrapolateSelection, -2 EmptyGrid, -3 Monotonicity,
# -4 IncompatibleShapes, -5 Other, 0 Ok.
VALIDATE_KATS = [
    dict(name="1d_nearest_enable", ref="src/interpolator/one/tests.rs:103-112", kind=FIXED, axes=[X5],
         values=F5, strategy=NEAREST, extrap=ENABLE, code=-1),
    dict(name="2d_nearest_enable", ref="src/interpolator/two/tests.rs:99-109", kind=FIXED, axes=[OX, OY],
         values=F22, strategy=NEAREST, extrap=ENABLE, code=-1),
    dict(name="3d_nearest_enable", ref="src/interpolator/three/tests.rs:132-143", kind=FIXED, axes=[OX, OY, OZ],
         values=F222, strategy=NEAREST, extrap=ENABLE, code=-1),
    dict(name="nd_nearest_enable", ref="src/interpolator/n/tests.rs:247-256", kind=ND, axes=[U2, U2, U2],
         values=F222, strategy=NEAREST, extrap=ENABLE, code=-1),
    dict(name="nd_mismatched_grid", ref="src/interpolator/n/tests.rs:350-361", kind=ND, axes=[U2, U2, U2],
         values=F22, strategy=LINEAR, extrap=ERROR, code=-5),
    dict(name="nd_0d_empty_grid_ok", ref="src/interpolator/n/tests.rs:362-369", kind=ND, axes=[[]],
         values=[0.], strategy=LINEAR, extrap=ERROR, code=0),
    dict(name="nd_0d_nonempty_grid", ref="src/interpolator/n/tests.rs:370-381", kind=ND, axes=[[1.]],
         values=[0.], strategy=LINEAR, extrap=ERROR, code=-5),
]


def node_points(axes):
    """All grid nodes (AoS) and their multi-indices, C-order."""
    idx = np.indices([len(a) for a in axes]).reshape(len(axes), -1).T
    pts = np.stack([np.asarray(axes[k], dtype=np.float64)[idx[:, k]] for k in range(len(axes))], axis=1)
    return pts, idx
