#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/r2_run16_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -4 gpurun_out/r2_run16_attn_tests.log
timeout 300 python tools/bench_ops.py attn > gpurun_out/r2_run16_attn.log 2>&1; echo "attn bench rc=$?"
cat gpurun_out/r2_run16_attn.log | tail -8
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu -s > gpurun_out/r2_run16_train_tests.log 2>&1; echo "train tests rc=$?"
grep "wgrad tc\|downsample\|linear \|tensor-core\|grad rel-L2\|passed\|failed\|Error\|assert" gpurun_out/r2_run16_train_tests.log | tail -22 | cut -c1-250
TB=8 timeout 300 python tools/profile_train.py > gpurun_out/r2_run16_train_profile.txt 2>&1; echo "train profile rc=$?"
head -24 gpurun_out/r2_run16_train_profile.txt | cut -c1-160
timeout 900 python bench.py --steps 10 --warmup 3 --secondary cfg2a,train > gpurun_out/r2_run16_bench.json 2> gpurun_out/r2_run16_bench.err; echo "bench rc=$?"
grep "secondary" gpurun_out/r2_run16_bench.err | cut -c1-900
