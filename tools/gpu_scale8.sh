#!/bin/bash
# strong scaling on 8 GPUs as the driver launches it (1e9 points split n/G)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 --no-variants > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
echo "rc=$?"; tail -c 1500 gpurun_out/scale_n8.json
