#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
{
echo "== nd5 wrap (binned tiles)"; timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap single"; NINT_ROUTE=single timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap tiles sc=24"; NINT_TILE_SC_SHIFT=24 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap tiles sc=22"; NINT_ROUTE=tile NINT_TILE_SC_SHIFT=22 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap bin only"; NINT_TILE_BIN_ONLY=1 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== 3d f64 tile"; NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --points 1e9 --launches 3
echo "== 3d f32 tile"; NINT_ROUTE=tile timeout 300 python tools/prof_case.py --case random --dtype f32 --points 1e9 --launches 3
} > gpurun_out/r2u_times.log 2>&1
tail -4 gpurun_out/r2u_pytest.log; grep -v "^+" gpurun_out/r2u_times.log | grep -v "launch 0"
