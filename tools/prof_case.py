#!/usr/bin/env python
"""Small driver for ncu captures: builds one BASELINE workload at reduced size and launches it a few times.

usage: python tools/prof_case.py --case random|sorted|nonuniform|nd5|2d|1d --points 1e8 --launches 3
Prints the CUDA-event time per launch (never quote a number printed under ncu).
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import ninterp_b200 as nb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="random")
    ap.add_argument("--points", type=float, default=1e8)
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--status", action="store_true")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--extrap", default="", help="nd5: wrap (default) or fill")
    ap.add_argument("--dtype", default="f64", help="random/sorted/nonuniform: f64 or f32")
    ap.add_argument("--strategy", default="linear", help="random/sorted/nonuniform/1d: linear or nearest")
    a = ap.parse_args()
    n = int(a.points)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    if a.case in ("random", "sorted", "nonuniform", "tilesorted", "morton"):
        axes, values = bench.grid_and_values(uniform=a.case != "nonuniform", GRID=a.grid)
        strat = nb.strategy.Nearest if a.strategy == "nearest" else nb.strategy.Linear
        npdt = np.float32 if a.dtype == "f32" else np.float64
        g = nb.Interp3D(*[ax.astype(npdt) for ax in axes], values.astype(npdt), strat, nb.Extrapolate.Error, dtype=npdt)
        cols = bench.gen_random(torch, n, dev, seed=1)
        if a.case == "sorted":
            bench.fill_cell_sorted(torch, cols, dev, seed=2, GRID=a.grid)
        elif a.case == "tilesorted":
            bench.fill_tile_sorted(torch, cols, dev, seed=2, GRID=a.grid)
        elif a.case == "morton":
            bench.fill_morton(torch, cols, dev, seed=2, GRID=a.grid)
        tdt = torch.float32 if a.dtype == "f32" else torch.float64
        if a.dtype == "f32":
            cols = [c.to(torch.float32).clamp_(0.0, 1.0) for c in cols]
    elif a.case == "nd5":
        axes = [np.cumsum(rng.uniform(0.5, 1.5, 32)).astype(np.float32) for _ in range(5)]
        values = rng.uniform(0, 1, (32,) * 5).astype(np.float32)
        g = nb.InterpND(axes, values, nb.strategy.Linear, nb.Extrapolate.Fill(-1.0) if a.extrap == "fill" else nb.Extrapolate.Wrap, dtype=np.float32)
        cols = [(torch.rand(n, dtype=torch.float32, device=dev) * 1.5 - 0.25) * float(ax[-1] - ax[0]) + float(ax[0]) for ax in axes]
        tdt = torch.float32
    elif a.case == "2d":
        G2 = a.grid if a.grid != 256 else 2048
        axes = [np.arange(float(G2))] * 2
        values = rng.uniform(0, 1, (G2, G2))
        g = nb.Interp2D(*axes, values, nb.strategy.Nearest if a.strategy == "nearest" else nb.strategy.Linear, nb.Extrapolate.Clamp)
        cols = [torch.rand(n, dtype=torch.float64, device=dev) * (G2 - 1) * 1.1 - 0.05 * (G2 - 1) for _ in range(2)]
        tdt = torch.float64
    else:
        gx = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, 999))]) if a.case == "1d" else np.arange(1000.0)
        g = nb.Interp1D(gx, rng.uniform(0, 1, 1000), nb.strategy.Nearest if a.strategy == "nearest" else nb.strategy.Linear, nb.Extrapolate.Error)
        cols = [torch.rand(n, dtype=torch.float64, device=dev) * gx[-1]]
        tdt = torch.float64
    out = torch.empty(n, dtype=tdt, device=dev)
    status = torch.empty(n, dtype=torch.uint8, device=dev) if a.status else None
    torch.cuda.synchronize()
    for k in range(a.launches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.interpolate_batch(cols, out=out, status=status)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{a.case} launch {k}: {ms:.3f} ms  {n / ms / 1e6:.2f} Gpt/s", flush=True)


if __name__ == "__main__":
    main()
