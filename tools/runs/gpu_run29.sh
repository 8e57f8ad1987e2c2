#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training.py -x -q -m gpu > gpurun_out/r2_run29_train_tests.log 2>&1; echo "train tests rc=$?"
tail -3 gpurun_out/r2_run29_train_tests.log | cut -c1-250
TB=32 timeout 600 python tools/train_graph.py > gpurun_out/r2_run29_train_graph_b32.log 2>&1; echo "train graph b32 rc=$?"
grep "^b=\|loss\|failed" gpurun_out/r2_run29_train_graph_b32.log | cut -c1-250
TB=8 timeout 600 python tools/train_graph.py > gpurun_out/r2_run29_train_graph_b8.log 2>&1; echo "train graph b8 rc=$?"
grep "^b=\|loss\|failed" gpurun_out/r2_run29_train_graph_b8.log | cut -c1-250
timeout 900 python bench.py --steps 5 --warmup 3 --secondary train --no-cpu-baseline > gpurun_out/r2_run29_bench.json 2> gpurun_out/r2_run29_bench.err; echo "bench rc=$?"
grep "secondary training" gpurun_out/r2_run29_bench.err | cut -c1-200
python - <<'PY'
import json
for line in open("gpurun_out/r2_run29_bench.json"):
    if line.startswith("{"):
        t = json.loads(line)["secondary"]["training_step"]
        print({k: v for k, v in t.items() if k != "workload"})
PY
