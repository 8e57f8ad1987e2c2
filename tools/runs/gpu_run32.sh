#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu > gpurun_out/r2_run32_train_tests.log 2>&1; echo "train tests rc=$?"
tail -2 gpurun_out/r2_run32_train_tests.log | cut -c1-200
TB=8 timeout 300 python tools/profile_train_host.py > gpurun_out/r2_run32_host_profile.txt 2>&1; echo "host profile rc=$?"
