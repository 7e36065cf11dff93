#!/usr/bin/env bash
# round 2, GPU call 48: bench.py with the spread of `value` (two more timed regions)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c48_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d["value_repeats"], round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["steps"], d["warmup"])
PY
