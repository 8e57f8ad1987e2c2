#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" > gpurun_out/r2_run47_memcheck_attn.log 2>&1; echo "memcheck attention rc=$?"
tail -4 gpurun_out/r2_run47_memcheck_attn.log | cut -c1-200
