#!/usr/bin/env bash
# round 2, GPU call 23: suite (f-4 device path, CorrNet tcgen05 option, uint8 normalisation fix), bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c23_tests.log 2>&1
echo "suite rc=$?"; grep -E "prefetch|corrnet tcgen05|pipeline with CorrNet" gpurun_out/r2c23_tests.log; tail -6 gpurun_out/r2c23_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err
tail -1 gpurun_out/r2c23_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c23_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"], d.get("e2e_uint8_images", {}).get("value"))
PY
