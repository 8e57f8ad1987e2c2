#!/bin/bash
mkdir -p gpurun_out
TB=32 timeout 600 python tools/train_graph.py > gpurun_out/r2_run31_train_graph_b32.log 2>&1; echo "train graph b32 rc=$?"
grep "^b=\|loss after" gpurun_out/r2_run31_train_graph_b32.log | cut -c1-250
TB=8 timeout 600 python tools/train_graph.py > gpurun_out/r2_run31_train_graph_b8.log 2>&1; echo "train graph b8 rc=$?"
grep "^b=\|loss after" gpurun_out/r2_run31_train_graph_b8.log | cut -c1-250
timeout 300 python -m pytest tests/test_gpu_unet.py -x -q -m gpu -k "tiny or smoke or cfg1" 2>&1 | tail -2
