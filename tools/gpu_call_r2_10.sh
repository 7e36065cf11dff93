#!/usr/bin/env bash
# round 2, GPU call 10: tc5h v2 (staging / epilogue prefetch / 5 CTAs per SM), GRU on tc5h A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c10_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c10_tests.log
IMVS_TUNE_TC5H_GRU=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "cfg2 or fixture or gru or full_size" -s > gpurun_out/r2c10_tests_gru.log 2>&1
echo "gru-tc5h rc=$?"; grep -E "cfg2|passed|failed|config2" gpurun_out/r2c10_tests_gru.log | tail -4
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --breakdown > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err
tail -1 gpurun_out/r2c10_bench.err
IMVS_TUNE_TC5H_GRU=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --breakdown > gpurun_out/r2c10_bench_gru.json 2> gpurun_out/r2c10_bench_gru.err
tail -1 gpurun_out/r2c10_bench_gru.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c10_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c10_ncu1.log 2>&1
