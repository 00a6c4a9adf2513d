#!/usr/bin/env python
"""Summarises `ncu --page source --csv` output: memory instructions with their L1/L2 work, and the top stall sites.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_source.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
g = lambda r, k: int(float(r[ix[k]] or 0)) if k in ix else 0
tot = sum(g(r, "Instructions Executed") for r in data)
samp = sum(g(r, "# Samples") for r in data)
print(f"warp instructions {tot}  samples {samp}")
print("memory instructions:")
for r in data:
    src = r[ix["Source"]].strip()
    ie = g(r, "Instructions Executed")
    if any(k in src for k in ("LDG", "STG", "ATOM", "LDS", "STS", "RED.", "LDL", "STL", "UTMA", "UBLK", "SYNCS")) and ie > tot / 2000:
        print(f"  {src[:64]:66s} inst={ie:>10} tag={g(r, 'L1 Tag Requests Global'):>10} shwf={g(r, 'L1 Wavefronts Shared'):>10} "
              f"ideal={g(r, 'L1 Wavefronts Shared Ideal'):>9} l2sec={g(r, 'L2 Theoretical Sectors Global'):>10} samp={g(r, '# Samples')}")
print("top stall sites:")
for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:16]:
    reasons = {k[6:]: g(r, k) for k in ix if k.startswith("stall_") and g(r, k) > 0}
    top = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {g(r, '# Samples'):>7} {r[ix['Source']].strip()[:70]:72s} {top}")
