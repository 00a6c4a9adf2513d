#!/usr/bin/env python
"""Turns the text files tools/gpu_profiles.sh brings back (gpurun_out/r2p_*) into profiles/r2_* and profiles/traffic.json.
usage: python tools/postprocess_profiles.py COMMIT   (the commit the capture was taken at)"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

NAMES = {"lists": "interp3d_linear_f64_256cubed_random_slab_lists", "slab": "interp3d_linear_f64_256cubed_random_slab_passes", "tiles": "interp3d_linear_f64_256cubed_random_binned_tiles_2.6e8_points",
         "sorted": "interp3d_linear_f64_256cubed_cell_sorted", "tilesorted": "interp3d_linear_f64_256cubed_tile_sorted_8cubed",
         "fused1d": "interp1d_linear_f64_1000_nodes_fused_4e8_points", "nd5fill": "interpnd5_linear_f32_32pow5_fill_inrange_compaction",
         "nd5wrap": "interpnd5_linear_f32_32pow5_wrap", "nd5wraplists": "interpnd5_linear_f32_32pow5_wrap_slab_lists", "near2d": "interp2d_nearest_f64_2048sq_clamp", "lin2d": "interp2d_linear_f64_2048sq_clamp"}


def kernels(path):
    ks, cur = [], None
    for ln in open(path):
        if ln.startswith("== kernel:"):
            cur = {"name": ln[11:].strip()}
            ks.append(cur)
        elif cur is not None and ln.startswith("dram__bytes_"):
            p = ln.split()
            cur[p[0]] = float(p[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(p[2], 1)
    return ks


def main():
    commit = sys.argv[1]
    # optional: --only key,key  (re)writes just those profile files (the others in gpurun_out/ are an earlier capture)
    only = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else None
    earlier = sys.argv[sys.argv.index("--earlier") + 1] if "--earlier" in sys.argv else None  # commit of that earlier capture
    out = ROOT / "gpurun_out"
    prof = ROOT / "profiles"
    for key, name in NAMES.items():
        summ = out / f"r2p_summary_{key}.txt"
        if not summ.exists() or (only is not None and key not in only):
            continue
        plain = (out / f"r2p_plain_{key}.log").read_text().strip().splitlines()
        text = [f"# ncu --set full --clock-control none --import-source on, B200, round 2, captured at commit {commit} (tools/gpu_profiles.sh)",
                f"# event-timed run of the same command, not under ncu (ONE cold launch incl. first-call setup; bench.py holds the warmed numbers): {plain[-1] if plain else "?"}", summ.read_text(), "",
                "## source page (tools/ncu_source.py): memory instructions and top stall sites", (out / f"r2p_source_{key}.txt").read_text()]
        (prof / f"r2_ncu_full_{name}.txt").write_text("\n".join(text))
    # launch list of the bench command
    src = out / "r2p_bench_launches.csv"
    if src.exists():
        (prof / "r2_ncu_launch_list_bench.csv").write_text(src.read_text())
        rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot = collections.OrderedDict()
        for r in rows[1:]:
            name = r[ki].split("(")[0].replace("void ", "")
            if "nint::" not in name:
                name = "(torch: synthetic data generation, checks)"
            v = float(r[vi]) / {"ns": 1e6, "us": 1e3, "ms": 1.0}.get(r[ui].strip(), 1e6)
            t = tot.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += v
        lines = [f"# launch list of `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` under ncu --metrics gpu__time_duration.sum, B200, commit {commit}",
                 "# per-launch times are cold-cache and serialised: compare SHARES", f"{'kernel':100s} {'launches':>8s} {'total ms':>12s}"]
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{k[:100]:100s} {n:8d} {ms:12.3f}")
        (prof / "r2_ncu_launch_list_bench_summary.txt").write_text("\n".join(lines) + "\n")
    # DRAM traffic per batch call
    tot = lambda k: k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]
    old = json.load(open(prof / "traffic.json"))
    have = all((out / f"r2p_summary_{k}.txt").exists() for k in ("slab", "tiles", "sorted", "lists"))
    if not have:
        print("not every route was captured: traffic.json left as it is")
        return
    slab, tiles, srt, lists = (kernels(out / f"r2p_summary_{k}.txt") for k in ("slab", "tiles", "sorted", "lists"))
    tj = {"note": "dram__bytes_read.sum + dram__bytes_write.sum from `ncu --set full` captures (profiles/r2_ncu_full_*.txt), per batch call of "
                  "1e9 points: the slab lists are scaled from a 2.6e8-point capture (scatter + evaluate + merge); the slab route is the order "
                  "probe + 3 slab_kernel passes; the binned-tiles figure is scaled from a 2.6e8-point capture (bin + evaluate); the single-pass figure on random queries is the round-1 capture (that kernel's random path is unchanged)",
          "commit": commit, "kernel_source_sha": bench.kernel_source_sha(), "algorithmic_bytes_per_launch": 32e9,
          **({"other_routes_captured_at": earlier + " (their kernels are unchanged since)"} if earlier else {}),
          "interp3d_linear_f64_uniform_random_lists_bytes_per_call": sum(tot(k) for k in lists) * 1e9 / 2.6e8,
          "interp3d_linear_f64_uniform_random_slab_bytes_per_call": sum(tot(k) for k in slab),
          "interp3d_linear_f64_uniform_random_slab_bytes_per_pass": tot(slab[0]),
          "interp3d_linear_f64_uniform_random_tiles_bytes_per_call": sum(tot(k) for k in tiles) * 1e9 / 2.6e8,
          "interp3d_linear_f64_uniform_random_bytes_per_launch": old["interp3d_linear_f64_uniform_random_bytes_per_launch"],
          "interp3d_linear_f64_cell_sorted_bytes_per_launch": tot(srt[0])}
    json.dump(tj, open(prof / "traffic.json", "w"), indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
