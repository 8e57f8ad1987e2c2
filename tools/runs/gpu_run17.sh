#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/r2_run17_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -4 gpurun_out/r2_run17_attn_tests.log
timeout 300 python tools/bench_ops.py attn > gpurun_out/r2_run17_attn.log 2>&1; echo "attn bench rc=$?"
cat gpurun_out/r2_run17_attn.log | tail -8
echo "--- BK=128 forced"
MI_ATTN_BK=128 timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run17_attn_bk128.log
echo "--- BK=256 forced"
MI_ATTN_BK=256 timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run17_attn_bk256.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -o gpurun_out/r2_run17_attn -f python tools/attn_one.py > gpurun_out/r2_run17_ncu.log 2>&1; echo "ncu rc=$?"
