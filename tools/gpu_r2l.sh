#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
{
echo "== 2d nearest clamp"; timeout 300 python tools/prof_case.py --case 2d --strategy nearest --points 4e8 --launches 4
echo "== 2d linear clamp"; timeout 300 python tools/prof_case.py --case 2d --points 4e8 --launches 4
} > gpurun_out/r2l_times.log 2>&1
tail -4 gpurun_out/r2l_pytest.log; grep -v "^+" gpurun_out/r2l_times.log | grep -v "launch 0"
