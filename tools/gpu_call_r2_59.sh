#!/usr/bin/env bash
# round 2, GPU call 59: final check of the last tree (GPU suite, smoke, default-flag bench, reference arm, config 5 stress shape)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c59_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2c59_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c59_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c59_smoke.log
timeout 600 python bench.py > gpurun_out/r2c59_bench.json 2> gpurun_out/r2c59_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c59_bench_ref.json 2> gpurun_out/r2c59_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c59_bench_cfg5.json 2> gpurun_out/r2c59_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c59_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d.get("value_repeats"), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["gpu_launches_per_step"],
      round(d["roofline"]["frac"], 4), d["cpu_baseline"]["value"], d.get("gpu_stock_ref", {}).get("value"), d["clocks"], d["stage_ms"])
r = json.loads(open("gpurun_out/r2c59_bench_ref.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("impl", "value", "unit") if k in r})
c = json.loads(open("gpurun_out/r2c59_bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5", round(c["value"], 1), round(c["e2e"]["value"], 1), round(c["single_stream"]["value"], 1), c["memory"]["max_allocated_MB"], c["stage_ms"])
PY
