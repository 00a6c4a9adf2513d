#!/bin/bash
# Sweeps the evict_last fraction of the value-table L2 policy on the headline workload (run on the GPU box).
for f in 0 0.25 0.5 0.625 0.75 0.875 1.0; do
  for c in random sorted; do
    echo "NINT_L2_FRAC=$f case=$c"
    NINT_L2_FRAC=$f python tools/prof_case.py --case $c --points 4e8 --launches 4 | tail -2
  done
done
