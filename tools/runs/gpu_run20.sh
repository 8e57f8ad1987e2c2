#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu > gpurun_out/r2_run20_train_tests.log 2>&1; echo "train tests rc=$?"
tail -3 gpurun_out/r2_run20_train_tests.log | cut -c1-250
TB=8 timeout 600 python tools/train_graph.py > gpurun_out/r2_run20_train_graph_b8.log 2>&1; echo "train graph b8 rc=$?"
grep -v Warning gpurun_out/r2_run20_train_graph_b8.log | head -30 | cut -c1-250
TB=32 timeout 600 python tools/train_graph.py > gpurun_out/r2_run20_train_graph_b32.log 2>&1; echo "train graph b32 rc=$?"
grep "^b=\|loss\|failed" gpurun_out/r2_run20_train_graph_b32.log | cut -c1-250
TB=32 timeout 300 python tools/profile_train.py > gpurun_out/r2_run20_train_profile_b32.txt 2>&1; echo "train profile b32 rc=$?"
head -14 gpurun_out/r2_run20_train_profile_b32.txt | cut -c1-160
