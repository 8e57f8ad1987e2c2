#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" > gpurun_out/r2_run23_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -15 gpurun_out/r2_run23_attn_tests.log | cut -c1-200
echo "--- two query tiles per CTA"
timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run23_attn_pair.log
echo "--- one query tile per CTA"
MI_ATTN_PAIR=0 timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run23_attn_single.log
timeout 900 python bench.py --workload cfg2a --steps 20 --warmup 3 --no-secondary --no-torch-gpu --no-cpu-baseline --kernel-table gpurun_out/r2_run23_kernel_table_cfg2a.txt > gpurun_out/r2_run23_bench_cfg2a.json 2> gpurun_out/r2_run23_bench_cfg2a.err; echo "bench cfg2a rc=$?"
tail -3 gpurun_out/r2_run23_bench_cfg2a.err | cut -c1-300
head -14 gpurun_out/r2_run23_kernel_table_cfg2a.txt | cut -c1-150
