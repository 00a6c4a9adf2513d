#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
{
echo "== 2d linear clamp (analytic default)"; timeout 300 python tools/prof_case.py --case 2d --points 4e8 --launches 4
echo "== 2d linear clamp (records)"; NINT_ANALYTIC=0 timeout 300 python tools/prof_case.py --case 2d --points 4e8 --launches 4
echo "== 2d nearest (analytic default)"; timeout 300 python tools/prof_case.py --case 2d --strategy nearest --points 4e8 --launches 4
echo "== 2d nearest (records)"; NINT_ANALYTIC=0 timeout 300 python tools/prof_case.py --case 2d --strategy nearest --points 4e8 --launches 4
echo "== 3d nearest random analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case random --strategy nearest --points 1e9 --launches 3
echo "== 3d nearest random"; timeout 300 python tools/prof_case.py --case random --strategy nearest --points 1e9 --launches 3
echo "== 3d nearest sorted analytic"; NINT_ANALYTIC=1 timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 3
echo "== 3d nearest sorted"; timeout 300 python tools/prof_case.py --case sorted --strategy nearest --points 1e9 --launches 3
echo "== tile (no L2 promotion)"; NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
} > gpurun_out/r2k_times.log 2>&1
tail -4 gpurun_out/r2k_pytest.log; grep -v "^+" gpurun_out/r2k_times.log | grep -v "launch 0"
