#!/usr/bin/env bash
# round 2, GPU call 42: tile-resident CorrNet with vertical strips (conflict-free shared-memory reads)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "corrnet_on_tma" > gpurun_out/r2c42_t1.log 2>&1
echo "t1 rc=$?"; tail -1 gpurun_out/r2c42_t1.log
IMVS_TUNE_CORR_TILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c42_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c42_ncu1.log 2>&1
IMVS_TUNE_CORR_TILE=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c42_bench.json 2> gpurun_out/r2c42_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c42_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["corrnet"], d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c42_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    if 'corrnet_tile' in r[ki]: print(r[ki][:60], r[vi])
PY
