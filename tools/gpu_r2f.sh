#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
{
for c in sorted tilesorted morton; do
  echo "== $c blocked"; timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
  echo "== $c grid-stride"; NINT_NO_BLOCKED=1 timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
done
echo "== f32 sorted blocked"; timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== f32 sorted grid-stride"; NINT_NO_BLOCKED=1 timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2f_times.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_case.py > gpurun_out/r2f_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2f_memcheck.log
tail -5 gpurun_out/r2f_pytest.log; grep -v "^+" gpurun_out/r2f_times.log | grep -v "launch 0"; tail -12 gpurun_out/r2f_memcheck.log
