#!/usr/bin/env python
"""bench.py — Interp3D Linear f64, 256^3 uniform grid, 1e9 query points per GPU (BASELINE.json configs[2]).

One step = one pass of the hot path (`Interpolator::interpolate` for every point of the batch) over the
whole synthetic batch.  `value` is device-timed with the queries already resident in HBM; `e2e` is the
same metric through the C ABI's host-buffer call (pinned host points in, host values out, copies inside
the timed region).  `--impl reference` times the CPU restatement of the reference (the oracle; the
Rust crate cannot be built in this image) on all host threads.

Launch: `python bench.py --gpus 1`, or under torchrun with one rank per GPU (weak scaling: every rank
evaluates its own `--points` queries against a replicated grid; there is no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gpoints/s Interp3D Linear f64"
UNIT = "Gpoints/s"
GRID = 256
BYTES_PER_POINT = 32  # 3 f64 coordinates in + 1 f64 value out; table and status bytes excluded


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=float, default=1e9, help="query points per GPU per step")
    ap.add_argument("--e2e-points", type=float, default=None, help="points per end-to-end step (default: --points, bounded by host RAM)")
    ap.add_argument("--no-variants", action="store_true", help="skip the cell-sorted / non-uniform variants")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def grid_and_values(seed=1234567890, uniform=True, GRID=GRID):
    rng = np.random.default_rng(seed)  # benches/benchmark.rs:13 seed
    if uniform:
        axes = [np.arange(GRID, dtype=np.float64) * (1.0 / (GRID - 1)) for _ in range(3)]
    else:
        axes = []
        for _ in range(3):
            g = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, GRID - 1))])
            axes.append(g / g[-1])
    values = rng.uniform(0.0, 1.0, (GRID, GRID, GRID))
    return axes, values


def config_dict(points, n_gpus, workload="interp3d_linear_f64_256cubed_uniform_grid_1e9_uniform_random_queries"):
    return {
        "workload": workload,
        "interpolator": "Interp3D", "strategy": "Linear", "extrapolate": "Error",
        "grid": [GRID, GRID, GRID], "grid_spacing": "uniform",
        "points_per_gpu": int(points), "query_order": "uniform random (unsorted)",
        "query_layout": "struct of arrays, resident in HBM",
        "l2_policy": "inputs (32 GB/step) exceed L2; no flush needed",
        "sharding": f"queries sharded over {n_gpus} GPU(s), grid replicated, no collective",
    }


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping: look the device up by UUID
            import torch

            self.h = None
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """The reference's CPU implementation of the path (oracle port), all host threads, bounded sample/step."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as O

    axes, values = grid_and_values()
    threads = host_threads(O)
    o = O.Oracle(axes, values, O.LINEAR, O.ERROR)
    rng = np.random.default_rng(99)
    # calibrate, then size the per-step sample so the whole run stays within ~2 minutes
    cal = [rng.random(2_000_000) for _ in range(3)]
    t0 = time.perf_counter()
    o.eval(cal, nthreads=threads)
    rate = 2_000_000 / (time.perf_counter() - t0)
    budget = 90.0 / max(args.steps + args.warmup, 1)
    n = int(min(args.points, max(1_000_000, rate * min(budget, 5.0))))
    cols = [rng.random(n) for _ in range(3)]
    for _ in range(args.warmup):
        o.eval(cols, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.eval(cols, nthreads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt / 1e9
    sample = f"{n} uniform random points per step of the 1e9-point workload, {threads} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.points, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of ninterp 0.8.1 (oracle/), not the Rust crate: no cargo/rustc in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpu": _cpu_model(),
    }
    print(json.dumps(line), flush=True)


def host_threads(O):
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the CPU arm wants)."""
    try:
        return max(O.max_threads(), len(os.sched_getaffinity(0)))
    except Exception:
        return max(O.max_threads(), os.cpu_count() or 1)


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------- b200 arm
def gen_random(torch, n, dev, seed, chunk=1 << 27):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    for c in cols:
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            c[s:e] = torch.rand(e - s, dtype=torch.float64, device=dev, generator=gen)
    return cols


def fill_cell_sorted(torch, cols, dev, seed, chunk=1 << 26, GRID=GRID):
    """Overwrites cols with queries visited in grid-cell order (what a pre-binned caller provides)."""
    n = cols[0].numel()
    ncell = (GRID - 1) ** 3
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    h = 1.0 / (GRID - 1)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = torch.arange(s, e, dtype=torch.int64, device=dev)
        cell = (i.to(torch.float64) * (ncell / n)).to(torch.int64).clamp_(max=ncell - 1)
        cz = cell % (GRID - 1)
        cy = (cell // (GRID - 1)) % (GRID - 1)
        cx = cell // ((GRID - 1) * (GRID - 1))
        for col, c in zip(cols, (cx, cy, cz)):
            u = torch.rand(e - s, dtype=torch.float64, device=dev, generator=gen)
            col[s:e] = ((c.to(torch.float64) + u) * h).clamp_(0.0, 1.0)
        del i, cell, cx, cy, cz


def time_steps(torch, dist, interp, cols, out, steps, warmup, world, sampler=None):
    """W warm-up + K timed steps; returns (max-over-ranks seconds for K steps, per-step kernel ms list, clocks).
    time_steps.launches = kernels this rank launched inside the timed region (a batch call is one launch, or an
    order probe plus its passes on tables well beyond L2)."""
    from ninterp_b200 import _lib

    for _ in range(warmup):
        interp.interpolate_batch(cols, out=out)
    torch.cuda.synchronize()
    launches0 = _lib.lib().nint_launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        interp.interpolate_batch(cols, out=out)
        ev[k + 1].record()
    torch.cuda.synchronize()
    time_steps.launches = _lib.lib().nint_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    total_ms = ev[0].elapsed_time(ev[steps])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    return total_ms / 1e3, per_step, clocks


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and so its pinned buffers, by first touch) to the CPUs NVML reports as local
    to its GPU.  Host-side plumbing for the end-to-end leg; skipped silently when the topology is not visible."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)
    except Exception:
        pass
    return None


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import ninterp_b200 as nb
    from ninterp_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: ninterp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(args.points)
    lib = _lib.lib()

    axes, values = grid_and_values()
    interp = nb.Interp3D(*axes, values, nb.strategy.Linear, nb.Extrapolate.Error, device=local_rank)
    cols = gen_random(torch, n, dev, seed=1234567890 + rank)
    out = torch.empty(n, dtype=torch.float64, device=dev)

    sampler = ClockSampler(local_rank)
    secs, per_step, clocks = time_steps(torch, dist, interp, cols, out, args.steps, args.warmup, world, sampler)
    launches_rank = int(time_steps.launches)
    launches = torch.tensor([launches_rank], dtype=torch.int64, device=dev)
    order = int(lib.nint_last_query_order(interp._h))  # 0 no probe, 1 coherent, 2 incoherent (slab passes)
    if world > 1:
        dist.all_reduce(launches)
    value = world * n * args.steps / secs / 1e9
    kernel_ms = float(np.mean(per_step))

    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = n * BYTES_PER_POINT / (kernel_ms * 1e-3) / 1e9
    # One batch call is the unit of accounting: a single interp_kernel launch, or (incoherent queries on a table
    # well beyond L2) the order probe plus the slab passes, whose summed time and summed DRAM traffic are reported.
    per_call = launches_rank // max(args.steps, 1)
    slab = order == 2
    traffic = None
    try:
        key = "interp3d_linear_f64_uniform_random_slab_bytes_per_call" if slab else "interp3d_linear_f64_uniform_random_bytes_per_launch"
        traffic = json.load(open(ROOT / "profiles" / "traffic.json")).get(key)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "kernel": (f"{per_call - 2} x slab_kernel<double,3,fixed,plain,uniform> + order_probe_kernel per batch call"
                           if slab else "interp_kernel<double,3,fixed,Linear,smem,plain,uniform>"),
                "kernel_ms": kernel_ms, "launches_per_call": per_call,
                "algorithmic_bytes_per_point": BYTES_PER_POINT, "points_per_call": n}

    # parity sample of the main workload, kept for the cpu_baseline leg
    sample_n = 0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = min(n, 64_000_000)
        sample_cols = [c[:sample_n].cpu().numpy() for c in cols]
        sample_out = out[:sample_n].cpu().numpy()

    variants = {}
    if not args.no_variants:
        # (ii) queries pre-sorted into grid cells
        fill_cell_sorted(torch, cols, dev, seed=7 + rank)
        s2, ps2, ck2 = time_steps(torch, dist, interp, cols, out, args.steps, args.warmup, world, ClockSampler(local_rank))
        ms2 = float(np.mean(ps2))
        variants["uniform_grid_cell_sorted_queries"] = {
            "value": world * n * args.steps / s2 / 1e9, "unit": UNIT, "kernel_ms": ms2, "kernel_ms_min": float(np.min(ps2)),
            "roofline_frac": n * BYTES_PER_POINT / (ms2 * 1e-3) / 1e9 / peak, "clocks": ck2}
        # (iii) non-uniform axes, unsorted random queries
        axes_nu, _ = grid_and_values(uniform=False)
        interp_nu = nb.Interp3D(*axes_nu, values, nb.strategy.Linear, nb.Extrapolate.Error, device=local_rank)
        del cols
        cols = gen_random(torch, n, dev, seed=1234567890 + rank)
        s3, ps3, _ = time_steps(torch, dist, interp_nu, cols, out, args.steps, args.warmup, world)
        ms3 = float(np.mean(ps3))
        variants["nonuniform_grid_random_queries"] = {
            "value": world * n * args.steps / s3 / 1e9, "unit": UNIT, "kernel_ms": ms3,
            "roofline_frac": n * BYTES_PER_POINT / (ms3 * 1e-3) / 1e9 / peak}
        interp_nu.close()

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, interp, cols, args, world, dev)

    del cols, out
    torch.cuda.empty_cache()

    cpu_baseline = None
    if sample_n:
        cpu_baseline = run_cpu_baseline(axes, values, sample_cols, sample_out, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(n, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches.item()), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "variants": variants, "gpu": torch.cuda.get_device_name(local_rank), "library": lib.nint_version().decode(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(torch, dist, interp, cols, args, world, dev):
    import psutil

    n_dev = cols[0].numel()
    want = int(args.e2e_points) if args.e2e_points else n_dev
    avail = psutil.virtual_memory().available
    cap = int(avail * 0.35 / world / BYTES_PER_POINT)  # pinned in + out per rank, well inside free RAM
    n = max(1_000_000, min(want, n_dev, cap))
    pts = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    res = torch.empty(n, dtype=torch.float64, pin_memory=True)
    chunk = 1 << 25
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        pts[s:e].copy_(torch.stack([c[s:e] for c in cols], dim=1))
    torch.cuda.synchronize()
    pts_np, res_np = pts.numpy(), res.numpy()
    steps = max(3, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        interp.interpolate_batch_host(pts_np, out=res_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        interp.interpolate_batch_host(pts_np, out=res_np)  # returns when res_np is complete on the host
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    checksum = float(res_np[:: max(1, n // 1000)].sum())
    out = {"value": world * n * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": n * 24 * world,
           "host_cpus_rank0": len(os.sched_getaffinity(0)),
           "d2h_bytes_per_step": n * 8 * world, "points_per_step_per_gpu": n, "steps": steps, "ms_per_step": dt / steps * 1e3,
           "api": "nint_interpolate_batch_host (AoS points in pinned host memory -> host values)", "checksum": checksum}
    del pts, res
    return out


def run_cpu_baseline(axes, values, sample_cols, sample_out, target_seconds):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as O

    threads = host_threads(O)
    o = O.Oracle(axes, values, O.LINEAR, O.ERROR)
    cal_n = min(2_000_000, len(sample_cols[0]))
    t0 = time.perf_counter()
    o.eval([c[:cal_n] for c in sample_cols], nthreads=threads)
    rate = cal_n / (time.perf_counter() - t0)
    n = int(min(len(sample_cols[0]), max(cal_n, rate * target_seconds)))
    cols = [c[:n] for c in sample_cols]
    t0 = time.perf_counter()
    want, st = o.eval(cols, nthreads=threads)
    dt = time.perf_counter() - t0
    got = sample_out[:n]
    mism = int(np.count_nonzero(want.view(np.uint64) != got.view(np.uint64)))
    # single-thread figure on a smaller slice
    n1 = max(100_000, int(n / threads / 4))
    t0 = time.perf_counter()
    o.eval([c[:n1] for c in cols], nthreads=1)
    dt1 = time.perf_counter() - t0
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} points of the step's batch, {dt:.1f} s on {threads} OpenMP threads",
            "single_thread_value": n1 / dt1 / 1e9, "host_cpu": _cpu_model(),
            "gpu_vs_oracle_mismatches_on_sample": mism, "status_nonzero": int(np.count_nonzero(st)),
            "note": "C++ restatement of ninterp 0.8.1 (oracle/), not the Rust crate: no cargo/rustc in this image"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
