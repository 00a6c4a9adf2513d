#!/bin/bash
# slab lists: slab prefetch at the start of a list, segment size, deeper record prefetch at 2 CTAs; ncu capture of the three kernels
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
echo "== parity pf auto"; NINT_PART_PREFETCH_KB=-1 NINT_LIB_VARIANT=pf timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_list_route" 2>&1 | tail -3
for kb in 0 -1 16 64 256; do
  echo "== pf lib, prefetch $kb KB"; NINT_PART_PREFETCH_KB=$kb NINT_LIB_VARIANT=pf timeout 300 $P --case random --points 1e9 --launches 4
done
for sh in 29 30; do
  echo "== pf lib, prefetch auto, segment 2^$sh"; NINT_PART_SEG_SHIFT=$sh NINT_PART_PREFETCH_KB=-1 NINT_LIB_VARIANT=pf timeout 300 $P --case random --points 1e9 --launches 4
  echo "== pf lib, no prefetch, segment 2^$sh"; NINT_PART_SEG_SHIFT=$sh NINT_LIB_VARIANT=pf timeout 300 $P --case random --points 1e9 --launches 4
done
echo "== a42"; NINT_LIB_VARIANT=a42 timeout 300 $P --case random --points 1e9 --launches 4
echo "== a42 prefetch auto"; NINT_PART_PREFETCH_KB=-1 NINT_LIB_VARIANT=a42 timeout 300 $P --case random --points 1e9 --launches 4
echo "== f32 pf auto"; NINT_PART_PREFETCH_KB=-1 NINT_LIB_VARIANT=pf timeout 300 $P --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2z2_times.log 2>&1
for kb in 0 -1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z2_launches_pf$kb.csv env NINT_PART_PREFETCH_KB=$kb NINT_LIB_VARIANT=pf $P --case random --points 2.6e8 --launches 2 > gpurun_out/r2z2_ncu_$kb.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_ -c 3 -o /tmp/r2z2_lists env NINT_LIB_VARIANT=pf $P --case random --points 2.6e8 --launches 1 > gpurun_out/r2z2_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r2z2_lists.ncu-rep "pf lib, no prefetch, 2.6e8 points" > gpurun_out/r2z2_summary_lists.txt 2>&1
ncu -i /tmp/r2z2_lists.ncu-rep --page source --csv > /tmp/r2z2_src.csv 2>/dev/null && python tools/ncu_source.py /tmp/r2z2_src.csv | awk '!seen[$0]++' > gpurun_out/r2z2_source_lists.txt 2>&1
ncu -i /tmp/r2z2_lists.ncu-rep --page details > gpurun_out/r2z2_details_lists.txt 2>&1
grep -v "^+" gpurun_out/r2z2_times.log | grep -v "launch 0"
