#!/usr/bin/env bash
# round 2, GPU call 17: interleaved weight order (one bulk copy per tap): suite, launch list, ncu --set full of tc5p kernels, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c17_tests.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/r2c17_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c17_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c17_ncu1.log 2>&1
tail -1 gpurun_out/r2c17_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc5p_conv_kernel -c 17 -o gpurun_out/r2c17_tc5p \
    python tools/profile_forward.py 1 > gpurun_out/r2c17_ncu2.log 2>&1
tail -1 gpurun_out/r2c17_ncu2.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --in-flight 2 > gpurun_out/r2c17_bench_if2.json 2> gpurun_out/r2c17_bench_if2.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --in-flight 8 > gpurun_out/r2c17_bench_if8.json 2> gpurun_out/r2c17_bench_if8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c17_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"])
    except Exception as e:
        print(f, "unreadable", e)
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c17_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:19]+rows[46:50]:
    print(r[ki][:100], r[vi])
PY
