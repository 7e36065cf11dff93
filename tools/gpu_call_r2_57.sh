#!/usr/bin/env bash
# round 2, GPU call 57: final check of the tree (GPU suite, smoke, the driver's two bench commands with default flags)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c57_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2c57_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c57_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c57_smoke.log
timeout 600 python bench.py > gpurun_out/r2c57_bench.json 2> gpurun_out/r2c57_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c57_bench_ref.json 2> gpurun_out/r2c57_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c57_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d.get("value_repeats"), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["gpu_launches_per_step"],
      round(d["roofline"]["frac"], 4), d["roofline"]["traffic"], d["cpu_baseline"]["value"], d.get("gpu_stock_ref", {}).get("value"), d["clocks"], d["stage_ms"])
r = json.loads(open("gpurun_out/r2c57_bench_ref.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("impl", "value", "unit") if k in r})
PY
