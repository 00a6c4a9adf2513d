#!/bin/bash
# static blend for four and five axes: full GPU suite, cfg 4 on every route
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2z5_pytest.log 2>&1
{
echo "== nd5 wrap, lists (default)"; timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap, lists, slab 96 MB"; NINT_PART_SLAB_MB=96 timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap, single pass"; NINT_PART_ND=0 timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 fill (in-range compaction)"; timeout 300 $P --case nd5 --extrap fill --points 2e8 --launches 4
echo "== nd5 wrap, lists, 3 CTAs"; NINT_LIB_VARIANT=c3 timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap, lists, 3 CTAs, slab 96 MB"; NINT_PART_SLAB_MB=96 NINT_LIB_VARIANT=c3 timeout 300 $P --case nd5 --points 2e8 --launches 4
} > gpurun_out/r2z5_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z5_launches_nd5.csv $P --case nd5 --points 2e8 --launches 2 > gpurun_out/r2z5_ncu_nd5.log 2>&1
tail -5 gpurun_out/r2z5_pytest.log; grep -v "^+" gpurun_out/r2z5_times.log | grep -v "launch 0"; grep "part_" gpurun_out/r2z5_launches_nd5.csv | awk -F'","' '{print $5, $NF}' | tail -3
