#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q -m gpu > gpurun_out/r2_run41_unet_tests.log 2>&1; echo "unet tests rc=$?"
tail -4 gpurun_out/r2_run41_unet_tests.log | cut -c1-250
timeout 600 python bench.py --steps 50 --warmup 5 --secondary cfg2a --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run41_bench.json 2> gpurun_out/r2_run41_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/r2_run41_bench.json"):
    if line.startswith("{"):
        j = json.loads(line); print("headline", round(j["value"], 2), round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"], 2), "launches/step", j.get("launches_per_step"), "cfg2a", round(j["secondary"]["cfg2a"]["steps_per_s"], 2))
PY
