#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "binned" > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
{
for sc in 22 23; do for fl in 0 4 6 8; do
  echo "== tile sc=$sc flags=$fl"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=$sc NINT_TILE_FLAGS=$fl timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
done; done
echo "== tile sc=21 flags=0"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=21 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== tile sc=22 page=8"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=22 NINT_TILE_PAGE_SHIFT=8 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
} > gpurun_out/r2d_times.log 2>&1
tail -3 gpurun_out/r2d_pytest.log; grep -v "^+" gpurun_out/r2d_times.log | grep -v "launch 0"
