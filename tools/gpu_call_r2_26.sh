#!/usr/bin/env bash
# round 2, GPU call 26: SM share of the persistent launches (1 = whole GPU, 2 = half, ...) with 4 reference views in flight
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 1 2 3 4; do
  IMVS_TUNE_TC5P_SHARE=$v timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c26_bench_share$v.json 2> gpurun_out/r2c26_bench_share$v.err
done
IMVS_TUNE_TC5P_SHARE=2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-u8 --in-flight 6 > gpurun_out/r2c26_bench_share2_if6.json 2> gpurun_out/r2c26_bench_share2_if6.err
IMVS_TUNE_TC5P_SHARE=2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-u8 --in-flight 3 > gpurun_out/r2c26_bench_share2_if3.json 2> gpurun_out/r2c26_bench_share2_if3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c26_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["featurenet"], d["stage_ms"]["gru"])
    except Exception as e:
        print(f, "unreadable", e)
PY
