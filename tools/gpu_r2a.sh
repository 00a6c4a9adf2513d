#!/bin/bash
# round 2, call A: parity of the binned route + first timings
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "binned or slab or concurrent or plain_table" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -k "config3" >> gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
{
for r in tile slab; do
  echo "== NINT_ROUTE=$r random 1e9"; NINT_ROUTE=$r timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 5
done
for sc in 21 22 24; do
  echo "== tile sc_shift=$sc"; NINT_TILE_SC_SHIFT=$sc timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
done
echo "== tile nonuniform"; timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 4
echo "== sorted"; timeout 300 python tools/prof_case.py --case sorted --points 1e9 --launches 4
echo "== f32 tile"; timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
echo "== f32 slab"; NINT_ROUTE=slab timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
} > gpurun_out/r2a_times.log 2>&1
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_times.log | grep -v "^+"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2a_bench.err
