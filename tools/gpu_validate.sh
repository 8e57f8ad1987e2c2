#!/bin/bash
# Full single-GPU validation: GPU parity suite, smoke, default bench (+ kernel table), reference arm.  Usage: gpurun -- 'bash tools/gpu_validate.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/validate_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/validate_gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/validate_smoke.log | cut -c1-200
timeout 1500 python bench.py --kernel-table gpurun_out/validate_kernel_table.txt > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
grep "secondary" gpurun_out/validate_bench.err | cut -c1-200 | tail -8
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/validate_bench_reference.json 2> gpurun_out/validate_bench_reference.err; echo "reference arm rc=$?"
