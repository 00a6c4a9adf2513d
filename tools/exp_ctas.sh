#!/bin/bash
# Tuning experiment: rebuild with extra nvcc defines and time the headline cases.
# usage: exp_ctas.sh "256:4 256:3:-DNINT_PREFETCH_L2 ..."   (block:min_ctas[:extra define])
for cfg in ${1:-"256:4 256:3"}; do
  IFS=: read b m x <<< "$cfg"
  export NINT_NVCC_EXTRA="-DNINT_BLOCK=$b -DNINT_MIN_CTAS=$m $x"
  python ninterp_b200/build.py > /dev/null 2>&1
  echo "BLOCK=$b MIN_CTAS=$m $x sorted 1e9: $(python tools/prof_case.py --case sorted --points 1e9 --launches 4 | tail -1)"
done
export NINT_NVCC_EXTRA=""
python ninterp_b200/build.py > /dev/null 2>&1
