#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "binned" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
{
for sc in 22 23; do for fl in 0 16; do
  echo "== tile sc=$sc flags=$fl"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=$sc NINT_TILE_FLAGS=$fl timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
done; done
echo "== tile sc=22 no analytic"; NINT_NO_ANALYTIC=1 NINT_ROUTE=tile NINT_TILE_SC_SHIFT=22 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== bin only"; NINT_TILE_BIN_ONLY=1 NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== bin only no analytic"; NINT_NO_ANALYTIC=1 NINT_TILE_BIN_ONLY=1 NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== f32 sc=23"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=23 timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 3
echo "== f32 sc=24"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=24 timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 3
} > gpurun_out/r2e_times.log 2>&1
for rows in 1 4; do timeout 200 tools/probes/tma_gather4_probe $rows > gpurun_out/r2e_gather4_box$rows.log 2>&1; echo "probe rows=$rows rc=$?"; done
tail -3 gpurun_out/r2e_pytest.log; grep -v "^+" gpurun_out/r2e_times.log | grep -v "launch 0"; cat gpurun_out/r2e_gather4_box*.log
