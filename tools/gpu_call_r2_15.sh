#!/usr/bin/env bash
# round 2, GPU call 15: tc5p v2 (uniform issue, 2 MMAs per tap/k-step): ubench, operator tests, suite, bench, launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 ./tools/ubench/umma_chain > gpurun_out/r2c15_ubench.txt 2>&1; head -14 gpurun_out/r2c15_ubench.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "conv3x3_tma" > gpurun_out/r2c15_op.log 2>&1
echo "op rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/r2c15_op.log | tail -8
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c15_tests.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/r2c15_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c15_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c15_ncu1.log 2>&1
tail -1 gpurun_out/r2c15_ncu1.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err
for v in "MB=1" "ST=2"; do
  env IMVS_TUNE_TC5P_$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c15_bench_$v.json 2> gpurun_out/r2c15_bench_$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c15_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["featurenet"])
    except Exception as e:
        print(f, "unreadable", e)
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c15_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:20]:
    print(r[ki][:90], r[vi])
PY
