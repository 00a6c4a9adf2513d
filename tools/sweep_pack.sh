#!/bin/bash
# Plain vs cell-packed table on the headline workload, and plain-layout behaviour against table size.
for pk in 0 1; do
  for c in random sorted nonuniform 2d nd5 1d; do
    echo "NINT_PACK=$pk case=$c"
    NINT_PACK=$pk python tools/prof_case.py --case $c --points 4e8 --launches 4 | tail -1
  done
done
for g in 96 128 160 192 224 256 320; do
  for pk in 0 1; do
    echo "grid=$g NINT_PACK=$pk random"
    NINT_PACK=$pk python tools/prof_case.py --case random --grid $g --points 2e8 --launches 4 | tail -1
  done
done
