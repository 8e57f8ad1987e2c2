#!/bin/bash
# round-2 GPU run 9: pair fused kernel with software-pipelined prologue (setmaxnreg) -- op tests, parity, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "fused_groupnorm" 2>&1 | tail -6 > gpurun_out/r2_run9_fused_op_pair.log
MI_FUSE_GN_CONV=pair timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "cfg3_full_size_vs_oracle or cfg3_structure or cascade" 2>&1 | grep -E "rel-L2|passed|failed|FAILED|Error" > gpurun_out/r2_run9_parity_pair.log
MI_FUSE_OVER_FOLD=0 timeout 300 python bench.py --fuse pair --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run9_kernel_table_pair_foldfirst.txt > gpurun_out/r2_run9_bench_pair_foldfirst.json 2> gpurun_out/r2_run9_bench_pair_foldfirst.err
timeout 300 python bench.py --fuse pair --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run9_kernel_table_pair.txt > gpurun_out/r2_run9_bench_pair.json 2> gpurun_out/r2_run9_bench_pair.err
timeout 300 python bench.py --fuse off --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run9_bench_off.json 2> gpurun_out/r2_run9_bench_off.err
ls -la gpurun_out | tail -8
