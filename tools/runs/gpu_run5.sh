#!/bin/bash
# round-2 GPU run 5: PDL (late trigger) A/B, full default bench, ncu launch list + DRAM capture of the dominant kernel family
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -m gpu -q --timeout 300 2>&1 | tail -5 > gpurun_out/r2_run5_training.log
timeout 300 python bench.py --pdl 1 --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run5_kernel_table_pdl.txt > gpurun_out/r2_run5_bench_pdl.json 2> gpurun_out/r2_run5_bench_pdl.err
timeout 900 python bench.py --kernel-table gpurun_out/r2_run5_kernel_table.txt > gpurun_out/r2_run5_bench_default.json 2> gpurun_out/r2_run5_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_run5_bench_reference.json 2> gpurun_out/r2_run5_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_run5_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run5_ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:halo_t -c 3000 --csv --log-file gpurun_out/r2_run5_halo_t_dram.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run5_ncu_dram.log 2>&1
ls -la gpurun_out | tail -10
