#!/bin/bash
for O in 0 1; do
  echo "--- MI_ATTN_ORDER=$O"
  MI_ATTN_ORDER=$O timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | cut -c1-90
done
MI_ATTN_ORDER=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" 2>&1 | tail -2
