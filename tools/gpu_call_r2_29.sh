#!/usr/bin/env bash
# round 2, GPU call 29: tile iterators without divisions in the tc5p roles: suite, bench default vs fused CorrNet, launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c29_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c29_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c29_bench_base.json 2> gpurun_out/r2c29_bench_base.err
IMVS_TUNE_CORR_FUSED=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c29_bench_fused.json 2> gpurun_out/r2c29_bench_fused.err
IMVS_TUNE_CORR_FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c29_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c29_ncu1.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c29_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"])
    except Exception as e:
        print(f, "unreadable", e)
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c29_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:21]+rows[28:34]+rows[36:40]:
    print(r[ki][:100], r[vi])
PY
