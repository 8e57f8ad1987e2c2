#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_run28_bench_2gpu.json 2> gpurun_out/r2_run28_bench_2gpu.err; echo "2-GPU bench rc=$?"
grep "^{" gpurun_out/r2_run28_bench_2gpu.json | cut -c1-300
