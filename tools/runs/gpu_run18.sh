#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu -s > gpurun_out/r2_run18_train_tests.log 2>&1; echo "train tests rc=$?"
grep "wgrad tc\|downsample\|linear \|tensor-core\|grad rel-L2\|passed\|failed\|Error\|assert" gpurun_out/r2_run18_train_tests.log | tail -22 | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/r2_run18_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -2 gpurun_out/r2_run18_attn_tests.log
timeout 300 python tools/bench_wgrad.py > gpurun_out/r2_run18_wgrad.log 2>&1; echo "wgrad bench rc=$?"
cat gpurun_out/r2_run18_wgrad.log
TB=8 timeout 600 python tools/train_graph.py > gpurun_out/r2_run18_train_graph_b8.log 2>&1; echo "train graph b8 rc=$?"
cat gpurun_out/r2_run18_train_graph_b8.log | cut -c1-300
TB=32 timeout 600 python tools/train_graph.py > gpurun_out/r2_run18_train_graph_b32.log 2>&1; echo "train graph b32 rc=$?"
cat gpurun_out/r2_run18_train_graph_b32.log | cut -c1-300
TB=8 timeout 300 python tools/profile_train.py > gpurun_out/r2_run18_train_profile.txt 2>&1; echo "train profile rc=$?"
head -16 gpurun_out/r2_run18_train_profile.txt | cut -c1-160
