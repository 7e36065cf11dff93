#!/usr/bin/env bash
# round 2, GPU call 44: FeatureNet conv1 with four rows per thread: suite, launch list, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c44_tests.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2c44_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c44_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c44_ncu1.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c44_bench.json 2> gpurun_out/r2c44_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c44_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["featurenet"], d.get("e2e_uint8_images", {}).get("value"))
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c44_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:3]: print(r[ki][:80], r[vi])
PY
