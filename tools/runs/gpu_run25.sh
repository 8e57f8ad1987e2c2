#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" > gpurun_out/r2_run25_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -3 gpurun_out/r2_run25_attn_tests.log | cut -c1-200
for P in 0 1; do
  echo "--- two MMA issuers, MI_ATTN_POLY=$P"
  MI_ATTN_POLY=$P timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | cut -c1-90 | tee gpurun_out/r2_run25_attn_poly$P.log
done
