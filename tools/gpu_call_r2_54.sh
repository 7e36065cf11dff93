#!/usr/bin/env bash
# round 2, GPU call 54: padded level 3 + one prologue launch as the default: GPU suite, smoke, bench A/B against IMVS_TUNE_WC_PAD3=0
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c54_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2c54_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c54_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c54_smoke.log
for i in 1 2 3; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c54_bench_pad3_$i.json 2> gpurun_out/r2c54_bench.err
  IMVS_TUNE_WC_PAD3=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c54_bench_off_$i.json 2>> gpurun_out/r2c54_bench.err
done
python - <<'PY'
import json
for tag in ("pad3", "off"):
    for i in (1, 2, 3):
        d = json.loads(open(f"gpurun_out/r2c54_bench_{tag}_{i}.json").read().strip().splitlines()[-1])
        st = d["stage_ms"]
        print(tag, i, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["gpu_launches_per_step"],
              "frac", round(d["roofline"]["frac"], 4), "iter_ms", round(d["roofline"]["avg_launch_ms"] * 1e3, 2), "init", st["warpcorr_init"], "compose", st["compose"])
PY
