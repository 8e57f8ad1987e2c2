#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_ops.py ln 2>&1 | tail -8 | cut -c1-160
timeout 900 python bench.py --workload cfg2a --steps 20 --warmup 3 --no-secondary --no-torch-gpu --no-cpu-baseline --kernel-table gpurun_out/r2_run40_kernel_table_cfg2a.txt > gpurun_out/r2_run40_bench_cfg2a.json 2> gpurun_out/r2_run40_bench_cfg2a.err; echo "bench cfg2a rc=$?"
grep "device-resident" gpurun_out/r2_run40_bench_cfg2a.err | cut -c1-100
head -10 gpurun_out/r2_run40_kernel_table_cfg2a.txt | cut -c1-150
