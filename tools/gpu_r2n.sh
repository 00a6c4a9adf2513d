#!/bin/bash
# N = 2 (and N = 8) runs of the bench as the driver launches them
set -x
mkdir -p gpurun_out
N=$1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
echo "bench n$N rc=$?"
tail -c 800 gpurun_out/r2n_bench_n$N.err
