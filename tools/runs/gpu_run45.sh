#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" > gpurun_out/r2_run45_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -12 gpurun_out/r2_run45_attn_tests.log | cut -c1-200
echo "--- persistent CTAs"
timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | cut -c1-90
echo "--- one work item per CTA"
MI_ATTN_PERSISTENT=0 timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | cut -c1-90
