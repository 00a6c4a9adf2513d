#!/bin/bash
# Register-budget experiment: rebuild with different minimum resident CTAs/SM and time the headline cases.
for m in 4 3 2; do
  NINT_NVCC_EXTRA="-DNINT_MIN_CTAS=$m" python ninterp_b200/build.py > /dev/null 2>&1
  for c in sorted random; do
    echo "MIN_CTAS=$m case=$c: $(NINT_NVCC_EXTRA="-DNINT_MIN_CTAS=$m" python tools/prof_case.py --case $c --points 4e8 --launches 4 | tail -1)"
  done
done
NINT_NVCC_EXTRA="" python ninterp_b200/build.py > /dev/null 2>&1
