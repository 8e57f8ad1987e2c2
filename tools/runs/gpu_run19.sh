#!/bin/bash
mkdir -p gpurun_out
echo "--- attention, MUFU only"
timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run19_attn_mufu.log
echo "--- attention, every 4th exp on the FMA pipe"
MI_ATTN_POLY=1 timeout 300 python tools/bench_ops.py attn 2>&1 | tail -6 | tee gpurun_out/r2_run19_attn_poly.log
MI_ATTN_POLY=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/r2_run19_attn_tests_poly.log 2>&1; echo "attn tests (poly) rc=$?"
tail -2 gpurun_out/r2_run19_attn_tests_poly.log
TB=8 timeout 600 python tools/train_graph.py > gpurun_out/r2_run19_train_graph_b8.log 2>&1; echo "train graph b8 rc=$?"
head -40 gpurun_out/r2_run19_train_graph_b8.log | cut -c1-250
TB=32 timeout 300 python tools/profile_train.py > gpurun_out/r2_run19_train_profile_b32.txt 2>&1; echo "train profile b32 rc=$?"
head -30 gpurun_out/r2_run19_train_profile_b32.txt | cut -c1-160
