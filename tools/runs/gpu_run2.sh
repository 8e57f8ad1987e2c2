#!/bin/bash
# round-2 GPU run 2: fused GroupNorm-prologue conv (op test, full-network parity, bench A/B), failing parity tests with output
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "fused_groupnorm" -rA 2>&1 | tail -40 > gpurun_out/r2_run2_fused_op.log
MI_FUSE_GN_CONV=1 timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "cfg3_full_size or cfg3_structure or tensor_core_configs or cfg5" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_run2_fused_net.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "more_seeds or tensor_core_configs" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_run2_unfused_net.log
timeout 300 python bench.py --fuse on --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run2_kernel_table_fused.txt > gpurun_out/r2_run2_bench_fused.json 2> gpurun_out/r2_run2_bench_fused.err
timeout 300 python bench.py --fuse off --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run2_bench_unfused.json 2> gpurun_out/r2_run2_bench_unfused.err
timeout 900 python bench.py > gpurun_out/r2_run2_bench_default.json 2> gpurun_out/r2_run2_bench_default.err
ls -la gpurun_out | tail -12
