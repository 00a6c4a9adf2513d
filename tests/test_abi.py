"""The C-ABI library loads, exports every symbol include/ninterp_cuda.h declares, and applies the
reference's construction-time validation before touching a device (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import kat_vectors as K
import oracle_binding as O

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "ninterp_cuda.h").read_text()
    return re.findall(r"NINT_API\s+[\w\s\*]+?\b(nint_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    from ninterp_b200 import _lib

    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ninterp_cuda.h but not exported"
    assert sorted(names) == sorted(_lib.SYMBOLS)
    assert b"ninterp" in lib.nint_version()


def test_rust_sys_crate_declares_the_same_abi():
    # integration/rust/ninterp-cuda-sys is source only (no Rust toolchain in this image): at least keep it in step
    # with the header -- same functions in the same order, same argument counts, same enum values
    header = (ROOT / "include" / "ninterp_cuda.h").read_text()
    rust = (ROOT / "integration" / "rust" / "ninterp-cuda-sys" / "src" / "lib.rs").read_text()
    rust_fns = re.findall(r"pub fn (nint_\w+)\s*\(([^)]*)\)", rust, flags=re.S)
    assert [n for n, _ in rust_fns] == _declared()
    for name, rust_args in rust_fns:
        c_args = re.search(r"\b%s\s*\(([^)]*)\)" % name, header, flags=re.S).group(1)
        n_c = 0 if c_args.strip() in ("", "void") else c_args.count(",") + 1
        n_rust = len([a for a in rust_args.split(",") if a.strip()])
        assert n_c == n_rust, f"{name}: {n_c} arguments in C, {n_rust} in Rust"
    enum_values = dict(re.findall(r"\b(NINT_[A-Z0-9_]+)\s*=\s*(-?\d+)", header))
    rust_consts = dict(re.findall(r"pub const (NINT_[A-Z0-9_]+): c_int = (-?\d+);", rust))
    assert enum_values and set(enum_values) <= set(rust_consts)
    for k, v in enum_values.items():
        assert rust_consts[k] == v, k
    # INTEGRATION.md quotes the same extern block, and the wrapper module verbatim
    doc = (ROOT / "INTEGRATION.md").read_text()
    for name, _ in rust_fns:
        assert f"pub fn {name}(" in doc, f"{name} is missing from INTEGRATION.md"
    sec2 = doc[doc.index("## 2. The batch method"):doc.index("## 3. What maps to what")]
    quoted = "".join(re.findall(r"```rust\n(.*?)```", sec2, flags=re.S))
    wrapper = (ROOT / "integration" / "rust" / "ninterp" / "src" / "cuda.rs").read_text()
    assert wrapper.endswith(quoted) and len(quoted) > 2000


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_rust_wrapper_matches_the_header_and_covers_every_interpolator():
    # integration/rust/ninterp/src/cuda.rs (source only): every sys:: call names a declared function and passes as
    # many arguments as the header takes; every sys:: constant exists; every interpolator of the reference gets
    # to_cuda; the ExtrapolateError text is the reference's
    header = (ROOT / "include" / "ninterp_cuda.h").read_text()
    wrapper = (ROOT / "integration" / "rust" / "ninterp" / "src" / "cuda.rs").read_text()
    declared = _declared()
    calls = 0
    for m in re.finditer(r"sys::(nint_\w+)\s*\(", wrapper):
        name = m.group(1)
        assert name in declared, f"{name} is not in include/ninterp_cuda.h"
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(wrapper[i], 0)
            i += 1
        n_rust = len(_split_args(wrapper[m.end():i - 1]))
        c_args = re.search(r"\b%s\s*\(([^)]*)\)" % name, header, flags=re.S).group(1)
        n_c = 0 if c_args.strip() in ("", "void") else c_args.count(",") + 1
        assert n_c == n_rust, f"{name}: {n_c} arguments in C, {n_rust} in the Rust call"
        calls += 1
    assert calls >= 9
    for const in set(re.findall(r"sys::(NINT_[A-Z0-9_]+)", wrapper)):
        assert re.search(r"\b%s\b" % const, header), f"{const} is not in the header"
    for interp, strat in (("Interp1D", "Strategy1D"), ("Interp2D", "Strategy2D"), ("Interp3D", "Strategy3D"), ("InterpND", "StrategyND")):
        assert re.search(r"to_cuda_impl!\(%s, %s, sys::NINT_KIND_(FIXED|ND)\)" % (interp, strat), wrapper), interp
    assert "to_cuda_impl!(InterpND, StrategyND, sys::NINT_KIND_ND)" in wrapper
    assert "impl<D> InterpolatorEnum<D>" in wrapper and "InterpolatorEnum::Interp0D(i) => CudaInterpolator::Constant" in wrapper
    for strat in ("Linear", "Nearest", "LeftNearest", "RightNearest", "Strategy1DEnum"):
        assert f"impl CudaStrategy for strategy::{'enums::' if strat.endswith('Enum') else ''}{strat}" in wrapper
    assert "two_variant_enum!(Strategy2DEnum, Strategy3DEnum, StrategyNDEnum)" in wrapper
    # three/mod.rs:226-229, one/mod.rs:198-203, n/mod.rs:326-329
    assert 'format!("\\n    point[{dim}] = {:?} is out of bounds for grid[{dim}] = {:?}", point[dim], g)' in wrapper


def test_no_cpu_fallback_without_gpu():
    import torch

    import ninterp_b200 as nb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nb.EngineError):
        nb.Interp1D([0., 1., 2.], [0., 1., 2.])


def _create_code(kind, axes, values, strategy, extrap, dtype=np.float64):
    """Return code of nint_create straight through the C ABI (handle destroyed if one was made)."""
    from ninterp_b200 import _lib

    lib = _lib.lib()
    axes = [np.ascontiguousarray(a, dtype=dtype) for a in axes]
    values = np.ascontiguousarray(values, dtype=dtype)
    ng = len(axes)
    axis_len = (C.c_size_t * max(ng, 1))(*[len(a) for a in axes])
    ptrs = (C.c_void_p * max(ng, 1))(*[a.ctypes.data if len(a) else None for a in axes])
    shape = (C.c_size_t * max(values.ndim, 1))(*values.shape)
    fill = np.zeros(1, dtype=dtype)
    h = C.c_void_p()
    rc = lib.nint_create(0, kind, 1 if dtype == np.float64 else 0, ng, axis_len, ptrs, values.ndim, shape,
                         C.c_void_p(values.ctypes.data) if values.size else None, None, strategy, extrap,
                         C.c_void_p(fill.ctypes.data), C.byref(h))
    if rc == 0:
        lib.nint_destroy(h)
    return rc, lib.nint_last_error_dim()


@pytest.mark.parametrize("v", K.VALIDATE_KATS, ids=[v["name"] for v in K.VALIDATE_KATS])
def test_validate_kats_through_c_abi(v):
    rc, _ = _create_code(v["kind"], v["axes"], v["values"], v["strategy"], v["extrap"])
    if v["code"] == 0:
        assert rc in (0, -9)  # passes validation; -9 = no CUDA device in this container
    else:
        assert rc == v["code"]


def test_validation_matches_oracle_on_random_bad_inputs():
    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(300):
        nd = int(rng.integers(1, 4))
        kind = int(rng.integers(0, 2))
        lens = [int(rng.integers(0, 4)) for _ in range(nd)]
        axes = []
        for L in lens:
            g = np.sort(rng.integers(0, 5, L).astype(np.float64))
            if L and rng.random() < 0.2:
                g = g[::-1].copy()
            if L and rng.random() < 0.1:
                g[int(rng.integers(0, L))] = np.nan
            axes.append(g)
        shape = [L if rng.random() < 0.8 else L + 1 for L in lens]
        values = np.zeros(shape)
        strategy = int(rng.integers(0, 2))
        extrap = int(rng.integers(0, 5))
        want = O.Oracle(axes, values, strategy, extrap, kind=kind).validate()
        rc, dim = _create_code(kind, axes, values, strategy, extrap)
        if want[0] == 0:
            assert rc in (0, -9)
        else:
            assert (rc, dim) == want, (kind, axes, shape, strategy, extrap)
            checked += 1
    assert checked > 50


def test_argument_errors():
    # Left/RightNearest exist for Interp1D only (strategy/enums/one.rs:8-13 vs three.rs:8-11)
    assert _create_code(0, [[0., 1.], [0., 1.]], np.zeros((2, 2)), 2, 4)[0] == -11
    assert _create_code(1, [[0., 1.]], np.zeros(2), 3, 4)[0] == -11
    # a hard-coded interpolator needs as many axes as value dimensions
    assert _create_code(0, [[0., 1.]], np.zeros((2, 2)), 0, 4)[0] == -10


def test_host_result_buffer_checks():
    # caller-supplied result buffers reach the C ABI as raw pointers: the mirror rejects a wrong dtype, a short,
    # a read-only or a non-contiguous buffer before the call (no GPU needed for the check itself)
    import ninterp_b200 as nb
    from ninterp_b200.interpolator import _check_host_out

    _check_host_out(np.empty(10), np.float64, 10)
    _check_host_out(np.empty(12, dtype=np.float32), np.float32, 10)
    for bad in (np.empty(10, dtype=np.float32), np.empty(9), np.empty(20)[::2], [0.0] * 10):
        with pytest.raises(nb.EngineError):
            _check_host_out(bad, np.float64, 10)
    ro = np.empty(10)
    ro.flags.writeable = False
    with pytest.raises(nb.EngineError):
        _check_host_out(ro, np.float64, 10)
