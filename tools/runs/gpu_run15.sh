#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -o gpurun_out/r2_run15_attn -f python tools/attn_one.py > gpurun_out/r2_run15_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_run15_ncu.log
TB=8 timeout 300 python tools/profile_train.py > gpurun_out/r2_run15_train_profile.txt 2>&1; echo "train profile rc=$?"
head -45 gpurun_out/r2_run15_train_profile.txt | cut -c1-200
