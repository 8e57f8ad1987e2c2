#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --secondary train --no-cpu-baseline > gpurun_out/r2_run27_bench.json 2> gpurun_out/r2_run27_bench.err; echo "bench rc=$?"
grep -v "Warning\|warn" gpurun_out/r2_run27_bench.err | tail -40 | cut -c1-250
./tools/mma_cost > gpurun_out/r2_run27_mma_cost.txt 2>&1; echo "mma_cost rc=$?"
cat gpurun_out/r2_run27_mma_cost.txt
