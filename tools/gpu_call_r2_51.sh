#!/usr/bin/env bash
# round 2, GPU call 51: final check of the tree (GPU suite incl. the new switch test, smoke, the driver's two bench commands)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c51_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2c51_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c51_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c51_smoke.log
timeout 600 python bench.py > gpurun_out/r2c51_bench.json 2> gpurun_out/r2c51_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c51_bench_ref.json 2> gpurun_out/r2c51_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c51_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d.get("value_repeats"), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["gpu_launches_per_step"],
      round(d["roofline"]["frac"], 4), d["cpu_baseline"], d.get("gpu_stock_ref"), d["clocks"])
r = json.loads(open("gpurun_out/r2c51_bench_ref.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("impl", "value", "unit", "cpu_baseline", "e2e") if k in r})
PY
