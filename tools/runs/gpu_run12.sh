#!/bin/bash
# round-2 GPU run 12: full -m gpu suite on the final code + the default bench (all secondary rows incl. the training step)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_run12_pytest_full.log 2>&1
grep -E "rel-L2|passed|failed|FAILED|ERROR|unet [01]:|tensor-core Block|grad rel-L2" gpurun_out/r2_run12_pytest_full.log | tail -60 > gpurun_out/r2_run12_pytest.log
timeout 900 python bench.py --kernel-table gpurun_out/r2_run12_kernel_table.txt > gpurun_out/r2_run12_bench_default.json 2> gpurun_out/r2_run12_bench_default.err
python __graft_entry__.py smoke > gpurun_out/r2_run12_smoke.log 2>&1
ls -la gpurun_out | tail -6
