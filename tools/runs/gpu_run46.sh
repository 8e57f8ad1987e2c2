#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_run46_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/r2_run46_gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-200
timeout 600 python bench.py --steps 50 --warmup 5 --secondary cfg2a,cfg4 --no-cpu-baseline --no-torch-gpu > gpurun_out/r2_run46_bench.json 2> gpurun_out/r2_run46_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/r2_run46_bench.json"):
    if line.startswith("{"):
        j = json.loads(line); s = j["secondary"]
        print("headline", round(j["value"], 2), round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"], 2), "cfg2a", round(s["cfg2a"]["steps_per_s"], 2), round(s["cfg2a"]["whole_step_frac"], 3),
              "cfg4", {k: (round(v["steps_per_s"], 2), round(v["whole_step_frac"], 3)) for k, v in s["cfg4"]["stages"].items()})
PY
