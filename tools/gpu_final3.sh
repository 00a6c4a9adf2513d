#!/bin/bash
# final validation + re-capture of the slab lists (3-D and 5-D) + the launch list of the bench command
bash tools/gpu_final.sh
ONLY="lists nd5wraplists" bash tools/gpu_profiles.sh
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_bench_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_bench_ncu.log 2>&1
echo "bench launch list rc=$?"
