#!/bin/bash
# round 2, call B: paged lists + pipelined evaluate: parity, timings, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "binned or concurrent" > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
{
for ps in 6 8 13; do
  echo "== tile page_shift=$ps"; NINT_TILE_PAGE_SHIFT=$ps timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
  echo "== tile page_shift=$ps bin only"; NINT_TILE_BIN_ONLY=1 NINT_TILE_PAGE_SHIFT=$ps timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
echo "== tile sc_shift=22"; NINT_TILE_SC_SHIFT=22 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
echo "== tile sc_shift=24"; NINT_TILE_SC_SHIFT=24 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
echo "== tile nonuniform"; timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 4
echo "== f32 tile"; timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2b_times.log 2>&1
timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 2 > gpurun_out/r2b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv python tools/prof_case.py --case random --points 1e9 --launches 2 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_pytest.log; grep -v "^+" gpurun_out/r2b_times.log
