#!/usr/bin/env python
"""bench.py — Interp3D Linear f64, 256^3 uniform grid, 1e9 uniformly random query points sharded over the GPUs
(BASELINE.json configs[2]; SURVEY.md section 8d cfg 3).

One step = one pass of the hot path (`Interpolator::interpolate` for every point of the batch) over the whole
synthetic batch.  `value` is device-timed with the queries already resident in HBM; `e2e` is the same metric
through the C ABI's host-buffer call (pinned host points in, host values out, copies inside the timed region),
with a plain-copy ceiling measured beside it.  `--impl reference` times the CPU restatement of the reference
(the oracle; the Rust crate cannot be built in this image) on all host threads.

Launch: `python bench.py --gpus 1`, or under torchrun with one rank per GPU.  Strong scaling: the 1e9 points are
split into contiguous ranges of n/G points (the loop being replaced is benches/benchmark.rs:149-151), the grid is
replicated and there is no data-path collective.  `variants` carries the other BASELINE configurations and query
orders, each with its own roofline fraction, clocks and parity sample.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
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
    ap.add_argument("--points", type=float, default=1e9, help="query points per step, all GPUs together")
    ap.add_argument("--e2e-points", type=float, default=None, help="points per end-to-end step per GPU (default: the rank's share, bounded by host RAM)")
    ap.add_argument("--no-variants", action="store_true", help="skip the other query orders and BASELINE configurations")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--parity-seconds", type=float, default=40.0, help="CPU time budget of the whole-batch parity report")
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
        "points_total": int(points), "points_per_gpu": int(points) // max(n_gpus, 1), "query_order": "uniform random (unsorted)",
        "query_layout": "struct of arrays, resident in HBM",
        "l2_policy": "inputs (32 GB/step over all GPUs) exceed L2; no flush needed",
        "sharding": f"queries split into {n_gpus} contiguous range(s) of n/G points, grid replicated, no collective",
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
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
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


def _git_head():
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:
        head = ""
    if head:
        return head
    stamp = ROOT / "ninterp_b200" / "_build" / "commit"  # written by ninterp_b200.build where .git exists (not on the GPU box)
    return stamp.read_text().strip() if stamp.exists() else None


def kernel_source_sha():
    """Hash of the CUDA sources: profiles/traffic.json records the one its ncu capture was taken with."""
    h = hashlib.sha256()
    for f in sorted((ROOT / "ninterp_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------- query generators (device)
def gen_random(torch, n, dev, seed, chunk=1 << 27, ncols=3, dtype=None, lo=0.0, hi=1.0):
    dtype = dtype or torch.float64
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cols = [torch.empty(n, dtype=dtype, device=dev) for _ in range(ncols)]
    for c in cols:
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            c[s:e] = torch.rand(e - s, dtype=dtype, device=dev, generator=gen)
            if lo != 0.0 or hi != 1.0:
                c[s:e].mul_(hi - lo).add_(lo)
    return cols


def _fill_from_cells(torch, cols, dev, seed, cell_of, chunk=1 << 26, GRID=GRID):
    """cols[d][i] = (cell_d(i) + U[0,1)) * h, where cell_of(i) -> (cx, cy, cz) int64 tensors for an index range."""
    n = cols[0].numel()
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    h = 1.0 / (GRID - 1)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = torch.arange(s, e, dtype=torch.int64, device=dev)
        cells = cell_of(i, n)
        for col, c in zip(cols, cells):
            u = torch.rand(e - s, dtype=torch.float64, device=dev, generator=gen)
            col[s:e] = ((c.clamp_(max=GRID - 2).to(torch.float64) + u) * h).clamp_(0.0, 1.0)
        del i, cells


def fill_cell_sorted(torch, cols, dev, seed, GRID=GRID):
    """Queries visited in grid-cell order, C order of the cells (what a caller that sorts by cell provides)."""
    ncell = (GRID - 1) ** 3

    def cell_of(i, n):
        cell = (i.to(torch.float64) * (ncell / n)).to(torch.int64).clamp_(max=ncell - 1)
        return cell // ((GRID - 1) * (GRID - 1)), (cell // (GRID - 1)) % (GRID - 1), cell % (GRID - 1)

    _fill_from_cells(torch, cols, dev, seed, cell_of, GRID=GRID)


def fill_tile_sorted(torch, cols, dev, seed, GRID=GRID, T=8):
    """Queries grouped by tiles of T^3 cells (tiles in C order), uniformly random inside a tile: what a cheap
    binning pass over coarse tiles emits."""
    nt = (GRID - 1 + T - 1) // T
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + 1)

    def cell_of(i, n):
        tile = (i.to(torch.float64) * (nt ** 3 / n)).to(torch.int64).clamp_(max=nt ** 3 - 1)
        out = []
        for t in (tile // (nt * nt), (tile // nt) % nt, tile % nt):
            out.append(t * T + torch.randint(0, T, t.shape, dtype=torch.int64, device=dev, generator=gen))
        return tuple(out)

    _fill_from_cells(torch, cols, dev, seed, cell_of, GRID=GRID)


def fill_morton(torch, cols, dev, seed, GRID=GRID):
    """Queries visited along the Morton (Z-order) curve of the cells."""
    bits = int(np.ceil(np.log2(GRID)))

    def cell_of(i, n):
        m = (i.to(torch.float64) * ((1 << (3 * bits)) / n)).to(torch.int64).clamp_(max=(1 << (3 * bits)) - 1)
        cx, cy, cz = torch.zeros_like(m), torch.zeros_like(m), torch.zeros_like(m)
        for b in range(bits):
            cz |= ((m >> (3 * b)) & 1) << b
            cy |= ((m >> (3 * b + 1)) & 1) << b
            cx |= ((m >> (3 * b + 2)) & 1) << b
        return cx, cy, cz

    _fill_from_cells(torch, cols, dev, seed, cell_of, GRID=GRID)


# ------------------------------------------------------------------------------------- timing
def time_steps(torch, dist, interp, cols, out, steps, warmup, world, sampler=None, status=None):
    """W warm-up + K timed steps; returns (max-over-ranks seconds for K steps, per-step kernel ms list, clocks).
    time_steps.launches = kernels this rank launched inside the timed region."""
    from ninterp_b200 import _lib

    for _ in range(warmup):
        interp.interpolate_batch(cols, out=out, status=status)
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
        interp.interpolate_batch(cols, out=out, status=status)
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


# ------------------------------------------------------------------------------------- parity
class Checker:
    """Compares device results with the CPU oracle (test infrastructure: the checker, never the thing measured)."""

    def __init__(self):
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_binding as O

        self.O = O
        self.threads = host_threads(O)

    def oracle(self, axes, values, strategy, extrap, fill=0.0, kind=0, dtype=np.float64):
        return self.O.Oracle(axes, values, strategy, extrap, fill=fill, kind=kind, dtype=dtype)

    @staticmethod
    def mismatches(got, want):
        u = np.uint64 if got.dtype == np.float64 else np.uint32
        same = (got.view(u) == want.view(u)) | (np.isnan(got) & np.isnan(want))
        return int(np.count_nonzero(~same))

    def check_slices(self, torch, o, cols, out, status, slices):
        """slices: list of (start, stop, step).  Returns a parity report over their union."""
        checked = mism = st_mism = 0
        oracle_s = 0.0
        t0 = time.perf_counter()
        for (a, b, s) in slices:
            sub = [c[a:b:s].contiguous().cpu().numpy() for c in cols]
            got = out[a:b:s].contiguous().cpu().numpy()
            t1 = time.perf_counter()
            want, st_want = o.eval(sub, nthreads=self.threads)
            oracle_s += time.perf_counter() - t1
            mism += self.mismatches(got, want)
            if status is not None:
                st_mism += int(np.count_nonzero(status[a:b:s].contiguous().cpu().numpy() != st_want))
            else:
                st_mism += int(np.count_nonzero(st_want))  # the timed workloads are all in range: every status is Ok
            checked += len(got)
        return {"points_checked": checked, "mismatches": mism, "status_mismatches": st_mism,
                "seconds": round(time.perf_counter() - t0, 2), "oracle_seconds": round(oracle_s, 3)}


def sample_slices(n, sample=64_000_000, tail=4_000_000):
    """A strided sample across the whole batch plus its last points."""
    step = max(1, n // sample)
    sl = [(0, n, step)] if step > 1 else [(0, n, 1)]
    if step > 1:
        sl.append((max(0, n - tail), n, 1))
    return sl


def whole_batch_report(torch, ck, o, cols, out, seconds):
    """Zero-mismatch report over the whole batch where the CPU keeps up (about 16 s per 1e9 points on 16 threads),
    else over evenly spread blocks (always the first and the last) that fit the time budget."""
    n = cols[0].numel()
    block = 1 << 25
    blocks = [(s, min(n, s + block), 1) for s in range(0, n, block)]
    rep = ck.check_slices(torch, o, cols, out, None, blocks[:1])
    per_block = max(rep["seconds"], 1e-3)
    k = int(min(len(blocks) - 1, max(1, seconds / per_block - 1)))
    rest = blocks[1:]
    pick = sorted({int(round(j * (len(rest) - 1) / max(k - 1, 1))) for j in range(k)}) if rest else []
    if rest and (len(rest) - 1) not in pick:
        pick.append(len(rest) - 1)
    rep2 = ck.check_slices(torch, o, cols, out, None, [rest[j] for j in pick]) if pick else {"points_checked": 0, "mismatches": 0, "status_mismatches": 0, "seconds": 0}
    total = {k_: rep[k_] + rep2.get(k_, 0) for k_ in ("points_checked", "mismatches", "status_mismatches", "oracle_seconds")}
    total["seconds"] = round(rep["seconds"] + rep2["seconds"], 2)
    total["coverage"] = ("whole batch" if total["points_checked"] == n else
                         f"{1 + len(pick)} of {len(blocks)} blocks of {block} points spread evenly over the batch, first and last included")
    total["oracle_threads"] = ck.threads
    return total


def bits_checksum(torch, out):
    """Order-independent checksum of the result bits (wrapping 64-bit sum)."""
    v = out.view(torch.int64) if out.dtype == torch.float64 else out.view(torch.int32).to(torch.int64)
    return int(v.sum().item()) & 0xFFFFFFFFFFFFFFFF


def routes_checksum(torch, interp, cols, out, status=None):
    """The same batch through every route the library has for it (slab lists, binned tiles, slab passes, single
    pass): the result bits must not depend on the route."""
    sums = {}
    saved = {k: os.environ.get(k) for k in ("NINT_ROUTE",)}
    try:
        for route in ("part", "tile", "slab", "single"):
            os.environ["NINT_ROUTE"] = route
            out.zero_()
            interp.interpolate_batch(cols, out=out, status=status)
            torch.cuda.synchronize()
            sums[route] = bits_checksum(torch, out)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return {"checksums": {k: f"{v:016x}" for k, v in sums.items()}, "identical": len(set(sums.values())) == 1}


# ------------------------------------------------------------------------------------- b200 arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import ninterp_b200 as nb
    from ninterp_b200 import _lib
    from ninterp_b200.sharding import shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: ninterp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_total = int(args.points)
    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo  # this rank's contiguous share (strong scaling)
    lib = _lib.lib()

    axes, values = grid_and_values()
    interp = nb.Interp3D(*axes, values, nb.strategy.Linear, nb.Extrapolate.Error, device=local_rank)
    cols = gen_random(torch, n, dev, seed=1234567890 + rank)
    out = torch.empty(n, dtype=torch.float64, device=dev)

    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    sampler = ClockSampler(local_rank)
    secs, per_step, clocks = time_steps(torch, dist, interp, cols, out, args.steps, args.warmup, world, sampler)
    launches_rank = int(time_steps.launches)
    launches = torch.tensor([launches_rank], dtype=torch.int64, device=dev)
    order = int(lib.nint_last_query_order(interp._h))  # 0 no probe, 1 coherent, 2 incoherent
    route = int(lib.nint_last_incoherent_route(interp._h)) if order == 2 else 0  # 1 slab passes, 2 binned tiles, 3 slab lists
    if world > 1:
        dist.all_reduce(launches)
    value = n_total * args.steps / secs / 1e9
    kernel_ms = float(np.mean(per_step))

    achieved = n * BYTES_PER_POINT / (kernel_ms * 1e-3) / 1e9
    # One batch call is the unit of accounting: a single interp_kernel launch, or (incoherent queries on a table
    # well beyond L2) the order probe plus bin + evaluate launches per segment (or the slab passes); their summed
    # time and summed DRAM traffic are what is reported.
    per_call = launches_rank // max(args.steps, 1)
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(ROOT / "profiles" / "traffic.json"))
        key = {3: "interp3d_linear_f64_uniform_random_lists_bytes_per_call", 2: "interp3d_linear_f64_uniform_random_tiles_bytes_per_call",
               1: "interp3d_linear_f64_uniform_random_slab_bytes_per_call"}.get(route, "interp3d_linear_f64_uniform_random_bytes_per_launch")
        traffic = tj.get(key)
        if traffic is not None and world > 1:
            traffic = traffic * n / 1e9  # the capture is of a 1e9-point call; a rank's share scales with its points
        traffic_note = {"source": "profiles/traffic.json (ncu capture, dram__bytes_read.sum + dram__bytes_write.sum)",
                        "capture_commit": tj.get("commit"), "capture_kernel_source_sha": tj.get("kernel_source_sha"),
                        "stale": tj.get("kernel_source_sha") != kernel_source_sha()}
    except Exception:
        pass
    kernel_name = {3: f"order_probe2_kernel + {(per_call - 2) // 3} x (part_scatter_kernel<double> + part_eval_kernel<double,...> + part_merge_kernel<double>) per batch call",
                   2: f"order_probe2_kernel + {(per_call - 2) // 2} x (tile_bin_kernel<double,...> + tile_eval_kernel<double>) per batch call",
                   1: f"order_probe2_kernel + {per_call - 2} x slab_kernel<double,3,...> per batch call"}.get(
        route, "interp_kernel<double,3,fixed,Linear,smem,plain,uniform>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "kernel": kernel_name, "kernel_ms": kernel_ms, "launches_per_call": per_call,
                "algorithmic_bytes_per_point": BYTES_PER_POINT, "points_per_call": n}

    ck = Checker() if (rank == 0 and not args.no_cpu_baseline) else None
    parity = None
    if ck:
        o3 = ck.oracle(axes, values, ck.O.LINEAR, ck.O.ERROR)
        parity = whole_batch_report(torch, ck, o3, cols, out, args.parity_seconds if world == 1 else 6.0)
        parity["routes"] = routes_checksum(torch, interp, cols, out)
    cpu_sample = None
    if ck and world == 1:
        m = min(n, 32_000_000)
        cpu_sample = ([c[:m].cpu().numpy() for c in cols], out[:m].cpu().numpy())

    def frac(points, ms, bpp=BYTES_PER_POINT):
        return points * bpp / (ms * 1e-3) / 1e9 / peak

    variants = {}
    if not args.no_variants:
        vsteps = max(5, min(args.steps, 10))

        def timed(interp_v, cols_v, out_v, status=None, steps=vsteps, bpp=BYTES_PER_POINT, oracle=None, sample=64_000_000):
            s, ps, ckk = time_steps(torch, dist, interp_v, cols_v, out_v, steps, args.warmup, world, ClockSampler(local_rank), status=status)
            ms = float(np.mean(ps))
            m = cols_v[0].numel()
            r = {"value": world * m * steps / s / 1e9, "unit": UNIT, "kernel_ms": ms, "kernel_ms_min": float(np.min(ps)),
                 "roofline_frac": frac(m, ms, bpp), "algorithmic_bytes_per_point": bpp, "points_per_gpu": m, "steps": steps,
                 "clocks": ckk, "query_order_verdict": int(lib.nint_last_query_order(interp_v._h)),
                 "incoherent_route": int(lib.nint_last_incoherent_route(interp_v._h))}
            if ck and oracle is not None:
                r["parity_report"] = ck.check_slices(torch, oracle, cols_v, out_v, status, sample_slices(m, sample, min(4_000_000, sample)))
                r["parity_report"]["checksum"] = f"{bits_checksum(torch, out_v):016x}"
            return r

        o3 = ck.oracle(axes, values, ck.O.LINEAR, ck.O.ERROR) if ck else None
        # cfg 3 +status: the same batch with the per-point status mask written (33 B/point)
        status = torch.empty(n, dtype=torch.uint8, device=dev)
        variants["random_queries_with_status_mask"] = timed(interp, cols, out, status=status, bpp=33, oracle=o3, sample=16_000_000)
        del status
        # cfg 3 forced onto the other routes, for the record: slab passes, binned tiles (TMA-staged, nint_tile.cuh), single pass
        for name, route_env in (("random_queries_slab_passes", "slab"), ("random_queries_binned_tiles", "tile"), ("random_queries_single_pass", "single")):
            os.environ["NINT_ROUTE"] = route_env
            variants[name] = timed(interp, cols, out, steps=5)
            os.environ.pop("NINT_ROUTE", None)
        # query orders a pre-binning caller provides
        for name, fill, seed in (("uniform_grid_cell_sorted_queries", fill_cell_sorted, 7), ("tile_sorted_8cubed_queries", fill_tile_sorted, 11),
                                 ("morton_queries", fill_morton, 13)):
            fill(torch, cols, dev, seed=seed + rank)
            variants[name] = timed(interp, cols, out, steps=args.steps if name.startswith("uniform_grid_cell") else vsteps,
                                   oracle=o3, sample=16_000_000)
        # non-uniform axes, unsorted random queries
        axes_nu, _ = grid_and_values(uniform=False)
        interp_nu = nb.Interp3D(*axes_nu, values, nb.strategy.Linear, nb.Extrapolate.Error, device=local_rank)
        del cols
        cols = gen_random(torch, n, dev, seed=1234567890 + rank)
        variants["nonuniform_grid_random_queries"] = timed(
            interp_nu, cols, out, oracle=ck.oracle(axes_nu, values, ck.O.LINEAR, ck.O.ERROR) if ck else None, sample=16_000_000)
        interp_nu.close()
        if world > 1:
            variants["note"] = "variants run on every rank's share; values are whole-job aggregates"
        variants.update(other_configs(torch, dist, nb, ck, timed, dev, local_rank, rank, world))

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, interp, cols, args, world, dev, rank, local_rank, axes, values)

    del cols, out
    torch.cuda.empty_cache()

    cpu_baseline = None
    if cpu_sample:
        cpu_baseline = run_cpu_baseline(axes, values, cpu_sample[0], cpu_sample[1], args.cpu_seconds)
        if parity and parity.get("oracle_seconds", 0) > cpu_baseline.get("seconds", 0):
            # the whole-batch parity report already ran the port over (most of) the workload: that is the larger sample
            cpu_baseline.update({
                "value": parity["points_checked"] / parity["oracle_seconds"] / 1e9,
                "sample": f"{parity['points_checked']} points of the step's batch ({parity['coverage']}), "
                          f"{parity['oracle_seconds']:.1f} s of oracle time on {parity['oracle_threads']} OpenMP threads",
                "gpu_vs_oracle_mismatches_on_sample": parity["mismatches"]})

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(n_total, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches.item()), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_report": parity, "variants": variants, "gpu": torch.cuda.get_device_name(local_rank),
            "library": lib.nint_version().decode(), "commit": _git_head(), "kernel_source_sha": kernel_source_sha(),
            "numa_cpus_rank0": len(numa_cpus) if numa_cpus else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(torch, dist, nb, ck, timed, dev, local_rank, rank, world):
    """BASELINE.json configs 1, 2 and 4 at their stated sizes (per GPU), device-resident inputs."""
    v = {}
    rng = np.random.default_rng(20261020)
    O = ck.O if ck else None
    # cfg 1: Interp1D Linear f64, 1,000-node non-uniform grid, Extrapolate::Error, 1e6 and 4e8 points (16 B/point)
    g = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, 999))])
    vals = rng.uniform(0, 1, 1000)
    it = nb.Interp1D(g, vals, nb.strategy.Linear, nb.Extrapolate.Error, device=local_rank)
    o = ck.oracle([g], vals, O.LINEAR, O.ERROR) if ck else None
    for npts, name in ((1_000_000, "cfg1_interp1d_linear_f64_1000_nodes_1e6_points"), (400_000_000, "cfg1_interp1d_linear_f64_1000_nodes_4e8_points")):
        cols = gen_random(torch, npts, dev, seed=21 + rank, ncols=1, hi=float(g[-1]))
        out = torch.empty(npts, dtype=torch.float64, device=dev)
        v[name] = timed(it, cols, out, bpp=16, oracle=o, sample=2_000_000)  # the reference scans all nodes per point: small sample
        del cols, out
    it.close()
    # cfg 2: Interp2D Linear + Nearest f64, 2048 x 2048, Clamp, 1e8 points, ~10 % out of range (24 B/point)
    ax2 = [np.arange(2048.0)] * 2
    v2 = rng.uniform(0, 1, (2048, 2048))
    cols = gen_random(torch, 100_000_000, dev, seed=22 + rank, ncols=2, lo=-0.05 * 2047, hi=1.05 * 2047)
    out = torch.empty(100_000_000, dtype=torch.float64, device=dev)
    for strat, code, name in ((nb.strategy.Linear, 0, "cfg2_interp2d_linear_f64_2048sq_clamp_1e8_points"),
                              (nb.strategy.Nearest, 1, "cfg2_interp2d_nearest_f64_2048sq_clamp_1e8_points")):
        it = nb.Interp2D(*ax2, v2, strat, nb.Extrapolate.Clamp, device=local_rank)
        v[name] = timed(it, cols, out, bpp=24, oracle=ck.oracle(ax2, v2, code, O.CLAMP) if ck else None, sample=16_000_000)
        it.close()
    del cols, out
    # cfg 4: InterpND Linear f32 N=5, 32^5 non-uniform grid, 2e8 points over [-0.25, 1.25) of the span, Wrap and Fill (24 B/point)
    ax5 = [np.cumsum(rng.uniform(0.5, 1.5, 32)).astype(np.float32) for _ in range(5)]
    v5 = rng.uniform(0, 1, (32,) * 5).astype(np.float32)
    cols = gen_random(torch, 200_000_000, dev, seed=24 + rank, ncols=5, dtype=torch.float32, lo=-0.25, hi=1.25)
    for c, a in zip(cols, ax5):
        c.mul_(float(a[-1] - a[0])).add_(float(a[0]))
    out = torch.empty(200_000_000, dtype=torch.float32, device=dev)
    for ext, code, fill, name in ((nb.Extrapolate.Wrap, 3, 0.0, "cfg4_interpnd_linear_f32_n5_32pow5_wrap_2e8_points"),
                                  (nb.Extrapolate.Fill(-1.0), 1, -1.0, "cfg4_interpnd_linear_f32_n5_32pow5_fill_2e8_points")):
        it = nb.InterpND(ax5, v5, nb.strategy.Linear, ext, dtype=np.float32, device=local_rank)
        v[name] = timed(it, cols, out, bpp=24,
                        oracle=ck.oracle(ax5, v5, O.LINEAR, code, fill=fill, kind=O.KIND_ND, dtype=np.float32) if ck else None, sample=2_000_000)
        it.close()
    del cols, out
    torch.cuda.empty_cache()
    return v


def copy_ceiling(torch, pts, res, n, steps):
    """The same bytes as one end-to-end step, as plain pinned copies (H2D of the points and D2H of the values on two
    streams at once, no kernels): the ceiling the host pipeline can reach on this link."""
    d_in = torch.empty((min(n, 1 << 26), 3), dtype=torch.float64, device="cuda")
    d_out = torch.empty(min(n, 1 << 26), dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    blk = d_in.shape[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for s in range(0, n, blk):
            e = min(n, s + blk)
            with torch.cuda.stream(s1):
                d_in[: e - s].copy_(pts[s:e], non_blocking=True)
            with torch.cuda.stream(s2):
                res[s:e].copy_(d_out[: e - s], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt / steps


def run_e2e(torch, dist, interp, cols, args, world, dev, rank, local_rank, axes, values):
    import psutil

    import ninterp_b200 as nb

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
    checksum = float(res_np[:: max(1, n // 1000)].sum())
    # the plain-copy ceiling, all ranks at once like the leg above
    if world > 1:
        dist.barrier()
    ceil_s = copy_ceiling(torch, pts, res, n, steps=max(2, steps // 2))
    if world > 1:
        t = torch.tensor([dt, ceil_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, ceil_s = float(t[0].item()), float(t[1].item())
    value = world * n * steps / dt / 1e9
    ceiling = world * n / ceil_s / 1e9
    out = {"value": value, "unit": UNIT, "h2d_bytes_per_step": n * 24 * world,
           "host_cpus_rank0": len(os.sched_getaffinity(0)),
           "d2h_bytes_per_step": n * 8 * world, "points_per_step_per_gpu": n, "steps": steps, "ms_per_step": dt / steps * 1e3,
           "api": "nint_interpolate_batch_host (AoS points in pinned host memory -> host values)", "checksum": checksum,
           "copy_ceiling": {"value": ceiling, "unit": UNIT, "h2d_gb_s": world * n * 24 / ceil_s / 1e9, "d2h_gb_s": world * n * 8 / ceil_s / 1e9,
                            "what": "the step's bytes as concurrent pinned cudaMemcpyAsync H2D + D2H on every rank, no kernels"},
           "frac_of_copy_ceiling": value / ceiling}
    # one process, all GPUs: nint_interpolate_batch_host_multi from rank 0 while the other ranks wait
    if world > 1:
        dist.barrier()
        multi = None
        if rank == 0:
            try:
                from ninterp_b200.multi import ReplicatedInterpolator

                rep = ReplicatedInterpolator(lambda d: nb.Interp3D(*axes, values, nb.strategy.Linear, nb.Extrapolate.Error, device=d), range(world))
                rep.interpolate_batch_host(pts_np, out=res_np)
                t0 = time.perf_counter()
                for _ in range(steps):
                    rep.interpolate_batch_host(pts_np, out=res_np)
                dtm = time.perf_counter() - t0
                multi = {"value": n * steps / dtm / 1e9, "unit": UNIT, "points_per_step": n, "gpus": world,
                         "api": "nint_interpolate_batch_host_multi (one process, one handle and one host thread per GPU)",
                         "checksum": float(res_np[:: max(1, n // 1000)].sum())}
                rep.close()
            except Exception as ex:  # the leg is informational
                multi = {"error": repr(ex)}
        dist.barrier()
        out["e2e_multi"] = multi
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
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
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
