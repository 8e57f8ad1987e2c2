#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --kernel-table gpurun_out/r2_run48_kernel_table.txt > gpurun_out/r2_run48_bench.json 2> gpurun_out/r2_run48_bench.err; echo "bench rc=$?"
grep "secondary" gpurun_out/r2_run48_bench.err | cut -c1-160 | tail -8
