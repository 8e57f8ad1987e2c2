#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention_single_sweep or (test_attention and 1024)" > gpurun_out/r2_run35_memcheck_attn.log 2>&1; echo "memcheck attention rc=$?"
tail -5 gpurun_out/r2_run35_memcheck_attn.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_training.py -x -q -m gpu -k "conv_wgrad_tensor_core or downsample or linear_gradients or backward_ops" > gpurun_out/r2_run35_memcheck_train.log 2>&1; echo "memcheck training rc=$?"
tail -5 gpurun_out/r2_run35_memcheck_train.log | cut -c1-200
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/r2_run35_memcheck_attn.log gpurun_out/r2_run35_memcheck_train.log
