#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "pipelined or config3 or slab_passes or binned" > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
{
for c in sorted morton tilesorted; do
  echo "== $c wide"; timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
  echo "== $c narrow"; NINT_NO_WIDE=1 timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
  echo "== $c wide classic kernel"; NINT_PIPE=0 timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
done
echo "== f32 sorted wide"; timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== f32 sorted narrow"; NINT_NO_WIDE=1 timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4
echo "== nonuniform sorted wide"; timeout 300 python tools/prof_case.py --case sorted --points 1e9 --launches 3
} > gpurun_out/r2x_times.log 2>&1
tail -3 gpurun_out/r2x_pytest.log; grep -v "^+" gpurun_out/r2x_times.log | grep -v "launch 0" | grep -v "launch 1"
