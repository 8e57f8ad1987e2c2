#!/bin/bash
# round-2 GPU run 7 (2 GPUs): the driver's multi-GPU launch line, with the secondary block (strong-scaling rows at N = 2)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_run7_bench_2gpu.json 2> gpurun_out/r2_run7_bench_2gpu.err
tail -5 gpurun_out/r2_run7_bench_2gpu.err
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 300 -k "sharding" 2>&1 | tail -3
