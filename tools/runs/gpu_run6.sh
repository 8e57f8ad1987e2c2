#!/bin/bash
# round-2 GPU run 6: gn_apply channel-slab version (op tests, micro-bench, step), gap statistics, steady-state ncu launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "groupnorm or gn_" 2>&1 | tail -4 > gpurun_out/r2_run6_gn_tests.log
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "cfg3_full_size_vs_oracle or tiny_unet or tensor_core_configs or cascade" 2>&1 | grep -E "rel-L2|passed|failed|FAILED" > gpurun_out/r2_run6_unet_tests.log
timeout 300 python tools/bench_ops.py gn > gpurun_out/r2_run6_bench_ops_gn.log 2>&1
timeout 300 python bench.py --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run6_kernel_table.txt > gpurun_out/r2_run6_bench.json 2> gpurun_out/r2_run6_bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run6_launches.csv python bench.py --steps 2 --warmup 3 --profiler-range --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run6_ncu_launches.log 2>&1
ls -la gpurun_out | tail -8
