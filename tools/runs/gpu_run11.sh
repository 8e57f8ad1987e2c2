#!/bin/bash
# round-2 GPU run 11: what bounds the pair fused kernel?  (profiling switches: results of these runs are WRONG, only kernel times are read)
set -x
mkdir -p gpurun_out
for dbg in 0 1 2 3 7; do
MI_GN_DBG=$dbg MI_FUSE_OVER_FOLD=1 timeout 300 python bench.py --fuse pair --steps 20 --no-secondary --no-cpu-baseline --no-torch-gpu --kernel-table gpurun_out/r2_run11_kt_dbg$dbg.txt > gpurun_out/r2_run11_bench_dbg$dbg.json 2> gpurun_out/r2_run11_bench_dbg$dbg.err
head -3 gpurun_out/r2_run11_kt_dbg$dbg.txt
done
