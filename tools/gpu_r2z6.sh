#!/bin/bash
# slab lists, 5-D f32 under Wrap at 3 CTAs: slab size, 4 CTAs; parity of the N-D route on the new build
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checked_build.py -x -q -k "slab_list_route or bounds_checked" 2>&1 | tail -5
} > gpurun_out/r2z6_pytest.log 2>&1
{
for mb in 96 128 160 224; do
  echo "== nd5 wrap, lists, slab $mb MB"; NINT_PART_SLAB_MB=$mb timeout 300 $P --case nd5 --points 2e8 --launches 4
done
echo "== nd5 wrap, lists, 4 CTAs, slab 96"; NINT_LIB_VARIANT=c4 timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== headline"; timeout 300 $P --case random --points 1e9 --launches 4
} > gpurun_out/r2z6_times.log 2>&1
grep -v "^+" gpurun_out/r2z6_pytest.log; grep -v "^+" gpurun_out/r2z6_times.log | grep -v "launch 0"
