#!/bin/bash
# round-2 GPU run 3: training path (native backward kernels), fused conv (C_out = 128 rule) A/B, parity numbers per seed
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training.py tests/test_gpu_ops.py -m gpu -q --timeout 600 -s -k "training or backward or gradcheck or p_losses or fused_groupnorm" > gpurun_out/r2_run3_training_full.log 2>&1
grep -E "rel-L2|passed|failed|FAILED|Error|error|unet [01]:|tensor-core Block" gpurun_out/r2_run3_training_full.log | tail -40 > gpurun_out/r2_run3_training.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "more_seeds or cfg5 or tensor_core_configs" > gpurun_out/r2_run3_parity_full.log 2>&1
grep -E "rel-L2|passed|failed|FAILED" gpurun_out/r2_run3_parity_full.log | tail -30 > gpurun_out/r2_run3_parity.log
MI_FUSE_GN_CONV=1 timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "cfg3_full_size_vs_oracle or cfg3_structure" 2>&1 | grep -E "rel-L2|passed|failed|FAILED" > gpurun_out/r2_run3_parity_fused.log
timeout 300 python bench.py --fuse on --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run3_kernel_table_fused128.txt > gpurun_out/r2_run3_bench_fused128.json 2> gpurun_out/r2_run3_bench_fused128.err
timeout 300 python bench.py --fuse off --no-secondary --no-cpu-baseline > gpurun_out/r2_run3_bench_unfused.json 2> gpurun_out/r2_run3_bench_unfused.err
ls -la gpurun_out | tail -12
