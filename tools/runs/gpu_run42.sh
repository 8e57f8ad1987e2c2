#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_run42_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r2_run42_gpu_tests.log | cut -c1-250
