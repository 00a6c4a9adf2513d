#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
{
for c in sorted morton tilesorted; do
  echo "== $c pipe"; timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
  echo "== $c classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
done
echo "== f32 sorted pipe"; timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== f32 sorted classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== nearest sorted pipe"; timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 4
echo "== nearest sorted classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 4
echo "== random single pipe"; NINT_ROUTE=single NINT_PIPE=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== random 128 pipe"; NINT_PIPE=1 timeout 300 python tools/prof_case.py --case random --grid 128 --points 1e9 --launches 3
echo "== random 128 classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case random --grid 128 --points 1e9 --launches 3
echo "== 2d linear pipe"; NINT_PIPE=1 timeout 300 python tools/prof_case.py --case 2d --grid 1024 --points 4e8 --launches 4
echo "== 2d linear classic"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case 2d --grid 1024 --points 4e8 --launches 4
echo "== slab analytic 4 CTAs"; NINT_LIB_VARIANT=ana4 NINT_SLAB_ANALYTIC=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
} > gpurun_out/r2q_times.log 2>&1
tail -4 gpurun_out/r2q_pytest.log; grep -v "^+" gpurun_out/r2q_times.log | grep -v "launch 0"
