#!/bin/bash
# slab lists: evaluate kernel with several groups per trip (e*), scatter/merge chunk shapes (s*), slab size; per-kernel times
set -x
mkdir -p gpurun_out
{
for v in "" e222 s1 s4; do
  echo "== parity variant '$v'"; NINT_LIB_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_list_route" 2>&1 | tail -3
done
} > gpurun_out/r2z_pytest.log 2>&1
{
for v in "" e222 e242 e332 e442 s1 s4 s5; do
  echo "== variant '$v'"; NINT_LIB_VARIANT=$v timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
for mb in 28 52; do
  echo "== slab $mb MB"; NINT_PART_SLAB_MB=$mb timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
for v in "" e222 e442; do
  echo "== nonuniform axes, lists forced, variant '$v'"; NINT_ROUTE=part NINT_LIB_VARIANT=$v timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 3
done
echo "== f32 default"; timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
echo "== f32 e222"; NINT_LIB_VARIANT=e222 timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2z_times.log 2>&1
for v in "" e222 s1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_$v.csv env NINT_LIB_VARIANT=$v python tools/prof_case.py --case random --points 2.6e8 --launches 2 > gpurun_out/r2z_ncu_$v.log 2>&1
done
cat gpurun_out/r2z_pytest.log | grep -v "^+"; grep -v "^+" gpurun_out/r2z_times.log | grep -v "launch 0"
