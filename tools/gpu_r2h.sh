#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
{
echo "== 1d fused"; timeout 300 python tools/prof_case.py --case 1d --points 4e8 --launches 4
echo "== 1d general"; NINT_FUSED_1D=0 timeout 300 python tools/prof_case.py --case 1d --points 4e8 --launches 4
echo "== 1d uniform fused"; timeout 300 python tools/prof_case.py --case 1d_uniform --points 4e8 --launches 4
echo "== 1d uniform general"; NINT_FUSED_1D=0 timeout 300 python tools/prof_case.py --case 1d_uniform --points 4e8 --launches 4
echo "== slab analytic"; NINT_SLAB_ANALYTIC=1 timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
echo "== slab"; timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 4
} > gpurun_out/r2h_times.log 2>&1
tail -4 gpurun_out/r2h_pytest.log; grep -v "^+" gpurun_out/r2h_times.log | grep -v "launch 0"
