#!/usr/bin/env bash
# round 2, GPU call 47: last check of the final tree: suite, smoke, default bench (no CPU legs)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c47_tests.log 2>&1
echo "suite rc=$?"; tail -1 gpurun_out/r2c47_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c47_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c47_smoke.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c47_bench.json 2> gpurun_out/r2c47_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c47_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["featurenet"], d["gpu_launches_per_step"], d["roofline"]["frac"])
PY
