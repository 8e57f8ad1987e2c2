#!/bin/bash
mkdir -p gpurun_out
for P in 0 1 2 3; do
  echo "--- MI_ATTN_POLY=$P"
  MI_ATTN_POLY=$P timeout 300 python tools/bench_ops.py attn 2>&1 | grep "n=4096 m=4096\|n=1024 m=1024\|n=4096 m=258" | cut -c1-90
done
MI_ATTN_POLY=3 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" 2>&1 | tail -2
AB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc2_kernel -s 2 -c 1 -o gpurun_out/r2_run24_attn2 -f python tools/attn_one.py > gpurun_out/r2_run24_ncu.log 2>&1; echo "ncu rc=$?"
AB=64 AN=4096 AM=258 timeout 600 ncu --set full --clock-control none -k regex:attn_tc2_kernel -s 2 -c 1 -o gpurun_out/r2_run24_attn2_cross -f python tools/attn_one.py > gpurun_out/r2_run24_ncu2.log 2>&1; echo "ncu2 rc=$?"
