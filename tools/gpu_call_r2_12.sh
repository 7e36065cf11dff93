#!/usr/bin/env bash
# round 2, GPU call 12: the persistent TMA + tcgen05 convolution (tc5pconv.cuh): operator test, suite, bench A/B, launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "conv3x3_tma" > gpurun_out/r2c12_op.log 2>&1
echo "op rc=$?"; grep -E "tc5p|passed|failed|Error|error" gpurun_out/r2c12_op.log | tail -30
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c12_tests.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/r2c12_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
tail -2 gpurun_out/r2c12_bench.err
IMVS_TUNE_TC5P=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench_off.json 2> gpurun_out/r2c12_bench_off.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c12_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c12_ncu.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c12_bench.json", "gpurun_out/r2c12_bench_off.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["single_stream"]["value"], d["stage_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
