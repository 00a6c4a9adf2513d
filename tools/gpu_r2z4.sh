#!/bin/bash
# slab lists, 5-D f32 under Wrap: slab size; ncu capture of the evaluate kernel
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
for mb in 56 72 96 128; do
  echo "== nd5 wrap, lists, slab $mb MB"; NINT_PART_SLAB_MB=$mb timeout 300 $P --case nd5 --points 2e8 --launches 4
done
} > gpurun_out/r2z4_times.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_eval -c 1 -o /tmp/r2z4_nd5 env NINT_PART_SLAB_MB=56 $P --case nd5 --points 2e8 --launches 1 > gpurun_out/r2z4_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r2z4_nd5.ncu-rep "nd5 wrap lists, slab 56 MB" > gpurun_out/r2z4_summary_nd5.txt 2>&1
ncu -i /tmp/r2z4_nd5.ncu-rep --page source --csv > /tmp/r2z4_src.csv 2>/dev/null && python tools/ncu_source.py /tmp/r2z4_src.csv | awk '!seen[$0]++' > gpurun_out/r2z4_source_nd5.txt 2>&1
grep -v "^+" gpurun_out/r2z4_times.log | grep -v "launch 0"
