#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
{
echo "== nd5 wrap single (fast wrap)"; NINT_ROUTE=single timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
for v in nd3 nd4; do echo "== nd5 wrap single variant=$v"; NINT_LIB_VARIANT=$v NINT_ROUTE=single timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3; done
echo "== nd5 wrap tiles"; timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap bin only"; NINT_TILE_BIN_ONLY=1 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap tiles binnd3"; NINT_LIB_VARIANT=binnd3 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 wrap bin only binnd3"; NINT_LIB_VARIANT=binnd3 NINT_TILE_BIN_ONLY=1 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 3
echo "== nd5 fill"; timeout 300 python tools/prof_case.py --case nd5 --extrap fill --points 2e8 --launches 3
for v in nd3 nd4; do echo "== nd5 fill variant=$v"; NINT_LIB_VARIANT=$v timeout 300 python tools/prof_case.py --case nd5 --extrap fill --points 2e8 --launches 3; done
} > gpurun_out/r2v_times.log 2>&1
tail -4 gpurun_out/r2v_pytest.log; grep -v "^+" gpurun_out/r2v_times.log | grep -v "launch 0"
