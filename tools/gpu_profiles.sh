#!/bin/bash
# round 2: ncu captures of every route at the final code (one gpurun call; summaries are made on the CPU box)
set -x
mkdir -p gpurun_out
P="python tools/prof_case.py"
cap() {  # name, kernel regex, count, command...
  local name=$1 regex=$2 count=$3; shift 3
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $name "; then return; fi
  "$@" > gpurun_out/r2p_plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -c $count -o /tmp/r2p_$name "$@" > gpurun_out/r2p_ncu_$name.log 2>&1
  echo "$name rc=$?"
  # the reports are tens of MB each (gpurun brings back at most 64 MiB): summarise here, keep the text
  python tools/ncu_summary.py /tmp/r2p_$name.ncu-rep "command: $*" > gpurun_out/r2p_summary_$name.txt 2>&1
  ncu -i /tmp/r2p_$name.ncu-rep --page source --csv > /tmp/r2p_src_$name.csv 2>/dev/null && python tools/ncu_source.py /tmp/r2p_src_$name.csv | awk '!seen[$0]++' > gpurun_out/r2p_source_$name.txt 2>&1
  rm -f /tmp/r2p_$name.ncu-rep /tmp/r2p_src_$name.csv
}
cap lists "part_" 3 $P --case random --points 2.6e8 --launches 1
cap slab "slab_kernel" 3 env NINT_ROUTE=slab $P --case random --points 1e9 --launches 1
cap sorted "interp_kernel" 1 $P --case sorted --points 1e9 --launches 1
cap tilesorted "interp_kernel" 1 $P --case tilesorted --points 1e9 --launches 1
cap tiles "tile_" 2 env NINT_ROUTE=tile $P --case random --points 2.6e8 --launches 1
cap fused1d "interp1d_fused" 1 $P --case 1d --points 4e8 --launches 1
cap nd5fill "slab_kernel" 1 $P --case nd5 --extrap fill --points 2e8 --launches 1
cap nd5wrap "interp_kernel" 1 env NINT_PART_ND=0 $P --case nd5 --points 2e8 --launches 1
cap nd5wraplists "part_" 3 $P --case nd5 --points 2e8 --launches 1
cap near2d "interp_kernel" 1 $P --case 2d --strategy nearest --points 1e8 --launches 1
cap lin2d "interp_kernel" 1 $P --case 2d --points 1e8 --launches 1
[ -n "$ONLY" ] || python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_bench_plain.log 2>&1 &&
[ -n "$ONLY" ] ||
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_bench_ncu.log 2>&1
echo "bench launch list rc=$?"
du -sh gpurun_out
