#!/bin/bash
# round-2 GPU run 13 (8 GPUs): the driver's launch line at N = 8 with all secondary rows
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_run13_bench_8gpu.json 2> gpurun_out/r2_run13_bench_8gpu.err
tail -3 gpurun_out/r2_run13_bench_8gpu.err | cut -c1-300
