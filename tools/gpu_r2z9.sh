#!/bin/bash
# slab lists: scatter as one CTA of 1024 threads per SM with the next chunk's columns a whole turn ahead (variant s1k)
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
echo "== parity s1k"; NINT_LIB_VARIANT=s1k timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_list_route" 2>&1 | tail -n 3
echo "== default"; timeout 300 $P --case random --points 1e9 --launches 5
echo "== s1k"; NINT_LIB_VARIANT=s1k timeout 300 $P --case random --points 1e9 --launches 5
echo "== f32 default"; timeout 300 $P --case random --dtype f32 --points 1e9 --launches 4
echo "== f32 s1k"; NINT_LIB_VARIANT=s1k timeout 300 $P --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2z9_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z9_launches_s1k.csv env NINT_LIB_VARIANT=s1k $P --case random --points 2.6e8 --launches 2 > gpurun_out/r2z9_ncu.log 2>&1
grep -v "^+" gpurun_out/r2z9_times.log | grep -v "launch 0"; grep "part_" gpurun_out/r2z9_launches_s1k.csv | awk -F'","' '{print $5, $NF}' | tail -n 3
