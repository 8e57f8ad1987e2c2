#!/bin/bash
mkdir -p gpurun_out
for B in 4 8; do
 for P in 0 1; do
  timeout 300 python bench.py --batch $B --pdl $P --steps 50 --warmup 5 --no-secondary --no-torch-gpu --no-cpu-baseline > gpurun_out/r2_run36_b${B}_pdl${P}.json 2> gpurun_out/r2_run36_b${B}_pdl${P}.err
  python - <<PY
import json
for line in open("gpurun_out/r2_run36_b${B}_pdl${P}.json"):
    if line.startswith("{"):
        j = json.loads(line); print("batch $B pdl $P:", round(j["value"], 2), "steps/s", round(j["ms_per_step"], 3), "ms/step, e2e", round(j["e2e"]["value"], 2))
PY
 done
done
timeout 300 python bench.py --batch 4 --steps 20 --warmup 5 --no-secondary --no-torch-gpu --no-cpu-baseline --kernel-table gpurun_out/r2_run36_kernel_table_b4.txt > /dev/null 2>&1
head -12 gpurun_out/r2_run36_kernel_table_b4.txt | cut -c1-140; grep "idle gaps" gpurun_out/r2_run36_kernel_table_b4.txt | cut -c1-200
