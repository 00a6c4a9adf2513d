#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "pipelined or config3 or slab_passes or binned" > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
{
for c in sorted morton tilesorted; do for pf in 0 1; do
  echo "== $c prefetch=$pf"; NINT_PIPE_PREFETCH=$pf timeout 300 python tools/prof_case.py --case $c --points 1e9 --launches 4
done; done
for pf in 0 1; do echo "== f32 sorted prefetch=$pf"; NINT_PIPE_PREFETCH=$pf timeout 300 python tools/prof_case.py --case sorted --dtype f32 --points 1e9 --launches 4; done
} > gpurun_out/r2r_times.log 2>&1
tail -3 gpurun_out/r2r_pytest.log; grep -v "^+" gpurun_out/r2r_times.log | grep -v "launch 0" | grep -v "launch 1"
