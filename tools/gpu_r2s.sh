#!/bin/bash
set -x
mkdir -p gpurun_out
{
for v in "" bin5 bin6; do
echo "== bin only variant=$v"; NINT_LIB_VARIANT=$v NINT_TILE_BIN_ONLY=1 NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== tile variant=$v"; NINT_LIB_VARIANT=$v NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
done
} > gpurun_out/r2s_times.log 2>&1
grep -v "^+" gpurun_out/r2s_times.log | grep -v "launch 0"
