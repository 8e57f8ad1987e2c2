#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu > gpurun_out/r2_run30_train_tests.log 2>&1; echo "train tests rc=$?"
tail -2 gpurun_out/r2_run30_train_tests.log | cut -c1-250
TB=32 timeout 300 python tools/profile_train.py > gpurun_out/r2_run30_train_profile_b32.txt 2>&1; echo "train profile b32 rc=$?"
head -26 gpurun_out/r2_run30_train_profile_b32.txt | cut -c1-150
TB=32 timeout 600 python tools/train_graph.py > gpurun_out/r2_run30_train_graph_b32.log 2>&1; echo "train graph b32 rc=$?"
grep "^b=" gpurun_out/r2_run30_train_graph_b32.log | head -3 | cut -c1-250
