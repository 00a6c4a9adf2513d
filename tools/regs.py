#!/usr/bin/env python
"""Prints registers / stack / spills per kernel instantiation from the nvcc -Xptxas -v logs."""
import re
import subprocess
import sys
from pathlib import Path

pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(Path(__file__).resolve().parent.parent.glob("ninterp_b200/_build/*.log")):
    txt = f.read_text()
    ents = re.findall(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", txt)
    for name, stack, spill, regs in ents:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(\w+_kernel<.*?>)\(", dem)
        label = m.group(1) if m else dem[:80]
        if pat in label:
            print(f"{label:70s} regs {regs:>3s} stack {stack:>5s} spill {spill}")
