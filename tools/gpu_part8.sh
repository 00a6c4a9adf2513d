#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x -m gpu -k "slab_list_route or config3 or concurrent" > gpurun_out/part_test.log 2>&1; echo "test rc=$?"
tail -3 gpurun_out/part_test.log
P="python tools/prof_case.py"
timeout 300 $P --case random --points 1e9 --launches 4 2>&1 | tail -2
timeout 300 $P --case sorted --points 1e9 --launches 4 2>&1 | tail -2
timeout 300 $P --case morton --points 1e9 --launches 4 2>&1 | tail -2
timeout 300 $P --case random --points 1e9 --launches 3 --dtype f32 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:part_ -c 3 --csv --log-file gpurun_out/part_launches.csv $P --case random --points 1e9 --launches 1 > gpurun_out/part_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/part_launches.csv | cut -d, -f5,15- | head -5
