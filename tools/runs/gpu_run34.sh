#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_run34_bench_8gpu.json 2> gpurun_out/r2_run34_bench_8gpu.err; echo "8-GPU bench rc=$?"
grep "^{" gpurun_out/r2_run34_bench_8gpu.json | cut -c1-300
