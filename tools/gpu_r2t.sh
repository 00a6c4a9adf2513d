#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "slab or config3 or concurrent or compaction" > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
{
echo "== slab random"; timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 5
echo "== slab nonuniform"; timeout 300 python tools/prof_case.py --case nonuniform --points 1e9 --launches 4
echo "== slab f32"; timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 4
echo "== slab 320"; timeout 300 python tools/prof_case.py --case random --grid 320 --points 1e9 --launches 3
} > gpurun_out/r2t_times.log 2>&1
tail -3 gpurun_out/r2t_pytest.log; grep -v "^+" gpurun_out/r2t_times.log | grep -v "launch 0"
