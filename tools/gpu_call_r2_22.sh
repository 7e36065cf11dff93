#!/usr/bin/env bash
# round 2, GPU call 22: CorrNet on the TMA + tcgen05 kernel (multi-set weights, stride-2, transposed layers)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "corrnet or evaluation_i" > gpurun_out/r2c22_t1.log 2>&1
echo "t1 rc=$?"; grep -E "passed|failed|Error|error|assert" gpurun_out/r2c22_t1.log | tail -8
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c22_tests.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/r2c22_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c22_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c22_ncu1.log 2>&1
tail -1 gpurun_out/r2c22_ncu1.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench.err
tail -2 gpurun_out/r2c22_bench.err
IMVS_TUNE_TC5P_CORR=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c22_bench_corroff.json 2> gpurun_out/r2c22_bench_corroff.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c22_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"])
    except Exception as e:
        print(f, "unreadable", e)
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c22_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[21:60]:
    print(r[ki][:100], r[vi])
PY
