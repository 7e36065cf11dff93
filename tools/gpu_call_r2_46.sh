#!/usr/bin/env bash
# round 2, GPU call 46: lateral epilogue reads fp32 chunk-planar coarse maps (no hi+lo reconstruction): suite, launch list, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c46_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c46_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c46_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c46_ncu1.log 2>&1
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c46_bench.json 2> gpurun_out/r2c46_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c46_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"])
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c46_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:21]:
    if 'Lateral' in r[ki] or '48, 16' in r[ki] or '48, 32' in r[ki]: print(r[ki][:100], r[vi])
PY
