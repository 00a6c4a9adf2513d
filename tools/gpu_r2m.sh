#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
echo "bench n1 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
echo "bench ref rc=$?"
tail -c 600 gpurun_out/r2m_bench_n1.err
