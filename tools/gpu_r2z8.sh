#!/bin/bash
# slab lists, scatter kernel: the next chunk's first column requested before the third column's round trip
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
echo "== parity ex"; NINT_LIB_VARIANT=ex timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_list_route" 2>&1 | tail -n 3
echo "== default"; timeout 300 $P --case random --points 1e9 --launches 5
echo "== early x"; NINT_LIB_VARIANT=ex timeout 300 $P --case random --points 1e9 --launches 5
echo "== f32 default"; timeout 300 $P --case random --dtype f32 --points 1e9 --launches 4
echo "== f32 early x"; NINT_LIB_VARIANT=ex timeout 300 $P --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2z8_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z8_launches_ex.csv env NINT_LIB_VARIANT=ex $P --case random --points 2.6e8 --launches 2 > gpurun_out/r2z8_ncu_ex.log 2>&1
grep -v "^+" gpurun_out/r2z8_times.log | grep -v "launch 0"; grep "part_" gpurun_out/r2z8_launches_ex.csv | awk -F'","' '{print $5, $NF}' | tail -n 3
