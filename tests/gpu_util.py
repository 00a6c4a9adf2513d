"""Helpers for the -m gpu parity tests: build the same interpolator on the CUDA path and in the oracle."""
import numpy as np

import oracle_binding as O

STRATS = None


def _nb():
    import ninterp_b200 as nb

    return nb


def make_pair(kind, axes, values, strategy, extrap, fill=0.0, dtype=np.float64):
    """(gpu interpolator, oracle) for identical data.  Codes are those of include/ninterp_cuda.h."""
    nb = _nb()
    strat = {0: nb.strategy.Linear, 1: nb.strategy.Nearest, 2: nb.strategy.LeftNearest, 3: nb.strategy.RightNearest}[strategy]
    ext = {0: nb.Extrapolate.Enable, 1: nb.Extrapolate.Fill(fill), 2: nb.Extrapolate.Clamp, 3: nb.Extrapolate.Wrap,
           4: nb.Extrapolate.Error}[extrap]
    axes = [np.asarray(a, dtype=dtype) for a in axes]
    values = np.asarray(values, dtype=dtype)
    if kind == O.KIND_ND:
        g = nb.InterpND(axes, values, strat, ext, dtype=dtype)
    else:
        g = [nb.Interp1D, nb.Interp2D, nb.Interp3D][len(axes) - 1](*axes, values, strat, ext, dtype=dtype)
    o = O.Oracle(axes, values, strategy, extrap, fill=fill, kind=kind, dtype=dtype)
    return g, o


def gpu_eval(g, coords, aos=False):
    """coords: list of numpy columns.  Returns (values, status, info) as numpy / dict via the device batch call."""
    import torch

    dev = torch.device("cuda", g.device)
    cols = [torch.from_numpy(np.ascontiguousarray(c, dtype=g.dtype)).to(dev) for c in coords]
    n = len(coords[0])
    status = torch.full((n,), 77, dtype=torch.uint8, device=dev)
    info = torch.zeros(4, dtype=torch.int64, device=dev)
    if aos:
        pts = torch.stack(cols, dim=1).contiguous()
        out = g.interpolate_batch(pts, status=status, info=info)
    else:
        out = g.interpolate_batch(cols, status=status, info=info)
    torch.cuda.synchronize(dev)
    i = info.cpu().numpy().view(np.uint64)
    return out.cpu().numpy(), status.cpu().numpy(), dict(n_error=int(i[0]), n_panic=int(i[1]), first_error=int(i[2]),
                                                        first_panic=int(i[3]))


def assert_same(got, st_got, want, st_want, ctx=""):
    """Bit-exact values (NaN payloads aside) and identical statuses."""
    assert np.array_equal(st_got, st_want), f"{ctx}: status mismatch at {np.flatnonzero(st_got != st_want)[:10]}"
    a = np.asarray(got)
    b = np.asarray(want)
    same = (a.view(np.uint64 if a.dtype == np.float64 else np.uint32) ==
            b.view(np.uint64 if b.dtype == np.float64 else np.uint32)) | (np.isnan(a) & np.isnan(b))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, f"{ctx}: {bad.size} value mismatches, first at {bad[:5]}: got {a[bad[:5]]} want {b[bad[:5]]}"


def expected_info(status):
    err = np.flatnonzero(status == 1)
    pan = np.flatnonzero(status == 0xFF)
    u = 2**64 - 1
    return dict(n_error=err.size, n_panic=pan.size, first_error=int(err[0]) if err.size else u,
                first_panic=int(pan[0]) if pan.size else u)
