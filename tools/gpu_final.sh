#!/bin/bash
# final validation of a round: the full GPU suite, smoke(), the bench as the driver runs it, the reference arm
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 1200 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
tail -4 gpurun_out/final_pytest.log; tail -2 gpurun_out/final_smoke.log
