#!/bin/bash
# round-2 GPU run 8: CTA-pair fused GroupNorm conv kernel -- op tests (pair and single-CTA forms), network parity, bench A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "fused_groupnorm" -rA 2>&1 | tail -25 > gpurun_out/r2_run8_fused_op_pair.log
MI_GN_NO_PAIR=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "fused_groupnorm" 2>&1 | tail -4 > gpurun_out/r2_run8_fused_op_single.log
MI_FUSE_GN_CONV=pair timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 600 -s -k "cfg3_full_size_vs_oracle or cfg3_structure or tensor_core_configs or cascade" 2>&1 | grep -E "rel-L2|passed|failed|FAILED|Error" > gpurun_out/r2_run8_parity_pair.log
timeout 300 python bench.py --fuse pair --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run8_kernel_table_pair.txt > gpurun_out/r2_run8_bench_pair.json 2> gpurun_out/r2_run8_bench_pair.err
MI_FUSE_OVER_FOLD=0 timeout 300 python bench.py --fuse pair --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run8_kernel_table_pair_foldfirst.txt > gpurun_out/r2_run8_bench_pair_foldfirst.json 2> gpurun_out/r2_run8_bench_pair_foldfirst.err
timeout 300 python bench.py --fuse on --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run8_bench_on.json 2> gpurun_out/r2_run8_bench_on.err
timeout 300 python bench.py --fuse off --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run8_bench_off.json 2> gpurun_out/r2_run8_bench_off.err
ls -la gpurun_out | tail -10
