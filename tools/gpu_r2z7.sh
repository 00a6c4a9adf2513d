#!/bin/bash
# 5-D f32 single pass and in-range compaction at 3 CTAs (NINT_ND_CTAS=3) after the blend fix; larger slabs for the wide lists
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
{
echo "== nd5 fill, compaction, 2 CTAs"; timeout 300 $P --case nd5 --extrap fill --points 2e8 --launches 4
echo "== nd5 fill, compaction, 3 CTAs"; NINT_LIB_VARIANT=nd3 timeout 300 $P --case nd5 --extrap fill --points 2e8 --launches 4
echo "== nd5 wrap, single pass, 2 CTAs"; NINT_PART_ND=0 timeout 300 $P --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap, single pass, 3 CTAs"; NINT_PART_ND=0 NINT_LIB_VARIANT=nd3 timeout 300 $P --case nd5 --points 2e8 --launches 4
for mb in 320 640; do
  echo "== nd5 wrap, lists, slab $mb MB"; NINT_PART_SLAB_MB=$mb timeout 300 $P --case nd5 --points 2e8 --launches 4
done
} > gpurun_out/r2z7_times.log 2>&1
grep -v "^+" gpurun_out/r2z7_times.log | grep -v "launch 0"
