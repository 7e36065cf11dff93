#!/usr/bin/env bash
# round 2, GPU call 13: tc5p operator tests, launch list, ncu --set full of the tc5p kernels, tunable sweep
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "conv3x3_tma or cfg5 or config5" > gpurun_out/r2c13_op.log 2>&1
echo "op rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/r2c13_op.log | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c13_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c13_ncu1.log 2>&1
tail -1 gpurun_out/r2c13_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc5p_conv_kernel -c 12 -o gpurun_out/r2c13_tc5p \
    python tools/profile_forward.py 1 > gpurun_out/r2c13_ncu2.log 2>&1
tail -1 gpurun_out/r2c13_ncu2.log
for v in "MB=1" "MB=2" "ST=2" "ST=3"; do
  env IMVS_TUNE_TC5P_$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c13_bench_$v.json 2> gpurun_out/r2c13_bench_$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c13_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["featurenet"])
    except Exception as e:
        print(f, "unreadable", e)
PY
