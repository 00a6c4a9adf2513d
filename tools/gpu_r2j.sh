#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
{
for sh in "4,4,5" "3,4,5" "3,3,5" "3,3,4" "2,3,4"; do
echo "== bin only sh=$sh"; NINT_TILE_SH=$sh NINT_TILE_BIN_ONLY=1 NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
done
echo "== tile sh=3,4,5"; NINT_TILE_SH=3,4,5 NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== tile default"; NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== tile default nonuniform"; NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 3
echo "== slab nonuniform"; timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 3
} > gpurun_out/r2j_times.log 2>&1
NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 2.6e8 --launches 2 > gpurun_out/r2j_plain.log 2>&1 &&
NINT_ROUTE=tile timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_eval -c 1 -o gpurun_out/r2j_prof_eval python tools/prof_case.py --case random --points 2.6e8 --launches 2 > gpurun_out/r2j_ncu.log 2>&1
tail -4 gpurun_out/r2j_pytest.log; grep -v "^+" gpurun_out/r2j_times.log | grep -v "launch 0"
