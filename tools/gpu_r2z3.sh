#!/bin/bash
# slab lists for InterpND with four and five axes: parity (regular and bounds-checked build), cfg 4 under Wrap against the single pass
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_list_route" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_checked_build.py -x -q 2>&1 | tail -15
} > gpurun_out/r2z3_pytest.log 2>&1
{
echo "== headline"; timeout 300 $P --case random --points 1e9 --launches 4
echo "== nd5 wrap, lists (default)"; timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap, single pass"; NINT_PART_ND=0 timeout 300 $P --case nd5 --points 2e8 --launches 4
for mb in 20 30 56; do
  echo "== nd5 wrap, lists, slab $mb MB"; NINT_PART_SLAB_MB=$mb timeout 300 $P --case nd5 --points 2e8 --launches 4
done
} > gpurun_out/r2z3_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z3_launches_nd5.csv $P --case nd5 --points 2e8 --launches 2 > gpurun_out/r2z3_ncu_nd5.log 2>&1
grep -v "^+" gpurun_out/r2z3_pytest.log; grep -v "^+" gpurun_out/r2z3_times.log | grep -v "launch 0"; grep "part_" gpurun_out/r2z3_launches_nd5.csv | awk -F'","' '{print $5, $NF}' | tail -3
