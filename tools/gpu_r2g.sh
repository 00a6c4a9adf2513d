#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "analytic or binned" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
{
for c in sorted tilesorted morton; do
  echo "== $c analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
done
echo "== f32 sorted analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== 2d analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case 2d --points 4e8 --launches 4
echo "== 2d"; timeout 300 python tools/prof_case.py --case 2d --points 4e8 --launches 4
echo "== random single pass analytic"; NINT_ROUTE=single NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== random 128 analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case random --grid 128 --points 1e9 --launches 3
echo "== random 128"; timeout 300 python tools/prof_case.py --case random --grid 128 --points 1e9 --launches 3
} > gpurun_out/r2g_times.log 2>&1
tail -3 gpurun_out/r2g_pytest.log; grep -v "^+" gpurun_out/r2g_times.log | grep -v "launch 0"
