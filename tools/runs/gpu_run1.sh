#!/bin/bash
# round-2 GPU run 1: full -m gpu suite, default bench (+secondary, kernel table), PDL experiment, ncu L2/smem look at halo_t
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA 2>&1 | tail -150 > gpurun_out/r2_run1_pytest.log
timeout 900 python bench.py --kernel-table gpurun_out/r2_run1_kernel_table.txt > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err
timeout 300 python bench.py --pdl 1 --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run1_kernel_table_pdl.txt > gpurun_out/r2_run1_bench_pdl.json 2> gpurun_out/r2_run1_bench_pdl.err
timeout 600 ncu --set full --clock-control none -k regex:halo_t -s 30 -c 6 -o gpurun_out/r2_run1_halo_t --force-overwrite python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run1_ncu.log 2>&1
ls -la gpurun_out | tail -20
