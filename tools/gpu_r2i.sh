#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
{
echo "== 1d fused ieee div"; timeout 300 python tools/prof_case.py --case 1d --points 4e8 --launches 4
echo "== 1d fused rcp"; NINT_FUSED_RCP=1 timeout 300 python tools/prof_case.py --case 1d --points 4e8 --launches 4
echo "== 1d uniform fused ieee"; timeout 300 python tools/prof_case.py --case 1d_uniform --points 4e8 --launches 4
echo "== nd5 fill compact"; timeout 300 python tools/prof_case.py --case nd5 --extrap fill --points 2e8 --launches 4
echo "== nd5 fill old"; NINT_COMPACT=0 timeout 300 python tools/prof_case.py --case nd5 --extrap fill --points 2e8 --launches 4
echo "== nd5 wrap"; timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 4
echo "== nd5 wrap plain table"; NINT_PACK=0 timeout 300 python tools/prof_case.py --case nd5 --points 2e8 --launches 4
} > gpurun_out/r2i_times.log 2>&1
tail -4 gpurun_out/r2i_pytest.log; grep -v "^+" gpurun_out/r2i_times.log | grep -v "launch 0"
