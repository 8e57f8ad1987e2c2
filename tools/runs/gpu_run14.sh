#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu -s > gpurun_out/r2_run14_train_tests.log 2>&1; echo "train tests rc=$?"
grep "wgrad tc\|tensor-core\|grad rel-L2\|passed\|failed\|Error" gpurun_out/r2_run14_train_tests.log | tail -14
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/r2_run14_attn_tests.log 2>&1; echo "attn tests rc=$?"
tail -4 gpurun_out/r2_run14_attn_tests.log
timeout 300 python tools/bench_wgrad.py > gpurun_out/r2_run14_wgrad.log 2>&1; echo "wgrad bench rc=$?"
cat gpurun_out/r2_run14_wgrad.log
timeout 300 python tools/bench_ops.py attn > gpurun_out/r2_run14_attn_single.log 2>&1; echo "attn bench rc=$?"
MI_ATTN_TWO_SWEEP=1 timeout 300 python tools/bench_ops.py attn > gpurun_out/r2_run14_attn_two.log 2>&1
echo "--- single sweep"; cat gpurun_out/r2_run14_attn_single.log | tail -12
echo "--- two sweeps"; cat gpurun_out/r2_run14_attn_two.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --secondary cfg2a,train --no-torch-gpu > gpurun_out/r2_run14_bench.json 2> gpurun_out/r2_run14_bench.err; echo "bench rc=$?"
grep "secondary" gpurun_out/r2_run14_bench.err | cut -c1-400
