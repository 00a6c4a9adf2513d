#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
NINT_LIB_VARIANT=checked timeout 1800 python -m pytest tests -x -q -m gpu -k "not division_fast_path" > gpurun_out/r2w_pytest_checked.log 2>&1
echo "pytest (bounds-checked build) rc=$?" >> gpurun_out/r2w_pytest_checked.log
tail -3 gpurun_out/r2w_pytest.log; tail -3 gpurun_out/r2w_pytest_checked.log
