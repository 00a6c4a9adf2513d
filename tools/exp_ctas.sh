#!/bin/bash
# Register-budget experiment: rebuild with different block sizes / minimum resident CTAs per SM and time
# the headline cases.  usage: exp_ctas.sh "256:4 256:3 128:7 ..."
for cfg in ${1:-"256:4 256:3"}; do
  b=${cfg%%:*}; m=${cfg##*:}
  export NINT_NVCC_EXTRA="-DNINT_BLOCK=$b -DNINT_MIN_CTAS=$m"
  python ninterp_b200/build.py > /dev/null 2>&1
  for c in sorted random; do
    echo "BLOCK=$b MIN_CTAS=$m case=$c: $(python tools/prof_case.py --case $c --points 4e8 --launches 4 | tail -1)"
  done
done
export NINT_NVCC_EXTRA=""
python ninterp_b200/build.py > /dev/null 2>&1
