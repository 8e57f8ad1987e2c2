#!/bin/bash
# round-2 GPU run 4: full -m gpu suite (folded res_conv default on), bench A/B: fold on/off, fused-GN v2 (C_out = 128)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_run4_pytest_full.log 2>&1
grep -E "rel-L2|passed|failed|FAILED|ERROR|unet [01]:|tensor-core Block" gpurun_out/r2_run4_pytest_full.log | tail -60 > gpurun_out/r2_run4_pytest.log
timeout 300 python bench.py --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run4_kernel_table_fold.txt > gpurun_out/r2_run4_bench_fold.json 2> gpurun_out/r2_run4_bench_fold.err
MI_FOLD_RES_CONV=0 timeout 300 python bench.py --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run4_bench_nofold.json 2> gpurun_out/r2_run4_bench_nofold.err
timeout 300 python bench.py --fuse on --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run4_kernel_table_fused128v2.txt > gpurun_out/r2_run4_bench_fused128v2.json 2> gpurun_out/r2_run4_bench_fused128v2.err
MI_SUBPIX_PAIR=1 timeout 300 python bench.py --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run4_bench_subpix_pair.json 2> gpurun_out/r2_run4_bench_subpix_pair.err
ls -la gpurun_out | tail -8
