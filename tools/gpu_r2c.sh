#!/bin/bash
# round 2, call C: L2 policies of the binned route, ncu full on both kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "binned" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
{
for fl in 0 1 2 4 6; do
  echo "== tile flags=$fl"; NINT_TILE_FLAGS=$fl timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
echo "== bin only flags=0"; NINT_TILE_BIN_ONLY=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
echo "== bin only flags=1"; NINT_TILE_BIN_ONLY=1 NINT_TILE_FLAGS=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
for sc in 21 22; do
echo "== tile sc_shift=$sc"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=$sc timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
echo "== f32 tile"; timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2c_times.log 2>&1
timeout 300 python tools/prof_case.py --case random --points 2.6e8 --launches 2 > gpurun_out/r2c_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_ -c 2 -o gpurun_out/r2c_prof_tiles python tools/prof_case.py --case random --points 2.6e8 --launches 2 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; grep -v "^+" gpurun_out/r2c_times.log
