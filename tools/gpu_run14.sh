#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu -s > gpurun_out/r2_run14_train_tests.log 2>&1; echo "train tests rc=$?"
tail -5 gpurun_out/r2_run14_train_tests.log
timeout 300 python tools/bench_wgrad.py > gpurun_out/r2_run14_wgrad.log 2>&1; echo "wgrad bench rc=$?"
cat gpurun_out/r2_run14_wgrad.log
timeout 600 python bench.py --steps 10 --warmup 3 --secondary train --no-torch-gpu > gpurun_out/r2_run14_bench_train.json 2> gpurun_out/r2_run14_bench_train.err; echo "bench rc=$?"
grep "training_step" gpurun_out/r2_run14_bench_train.err | tail -2
