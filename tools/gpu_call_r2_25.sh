#!/usr/bin/env bash
# round 2, GPU call 25 (final tree, 1 GPU): suite, smoke, full bench line with the reference arms, launch list, config 5 line,
# ncu --set full of the roofline kernel and of the largest tcgen05 kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c25_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c25_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c25_smoke.log
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2c25_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c25_bench_ref.json 2> gpurun_out/r2c25_bench_ref.err
echo "ref rc=$?"
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c25_bench_cfg5.json 2> gpurun_out/r2c25_bench_cfg5.err
echo "cfg5 rc=$?"; tail -1 gpurun_out/r2c25_bench_cfg5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c25_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c25_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:warpcorr_iter_kernel -c 1 -o gpurun_out/r2c25_warpcorr \
    python tools/profile_forward.py 1 > gpurun_out/r2c25_ncu2.log 2>&1
tail -1 gpurun_out/r2c25_ncu2.log
python - <<'PY'
import json
for f in ("gpurun_out/r2c25_bench.json", "gpurun_out/r2c25_bench_ref.json", "gpurun_out/r2c25_bench_cfg5.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), d.get("e2e", {}).get("value"), d.get("single_stream", {}).get("value"), d.get("stage_ms"), d.get("cpu_baseline", {}).get("value"), d.get("gpu_stock_ref", {}).get("value"), d.get("memory"))
    except Exception as e:
        print(f, "unreadable", e)
PY
