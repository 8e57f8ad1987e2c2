#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 --secondary cfg4,cfg5 --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run44_bench_4gpu.json 2> gpurun_out/r2_run44_bench_4gpu.err; echo "4-GPU bench rc=$?"
grep "^{" gpurun_out/r2_run44_bench_4gpu.json | cut -c1-260
