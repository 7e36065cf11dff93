#!/usr/bin/env bash
# round 2, GPU call 5: fused tcgen05 head: parity tests, bench with/without, launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "fused_tcgen05_head" > gpurun_out/r2c5_head_test.log 2>&1
echo "head test rc=$?"; tail -15 gpurun_out/r2c5_head_test.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c5_tests.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/r2c5_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --breakdown > gpurun_out/r2c5_bench_fused.json 2> gpurun_out/r2c5_bench_fused.err
IMVS_TUNE_HEADFUSED=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --breakdown > gpurun_out/r2c5_bench_unfused.json 2> gpurun_out/r2c5_bench_unfused.err
tail -2 gpurun_out/r2c5_bench_fused.err gpurun_out/r2c5_bench_unfused.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 100 --csv --log-file gpurun_out/r2c5_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c5_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 2 -c 1 -o gpurun_out/r2c5_headfused \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c5_ncu2.log 2>&1
tail -3 gpurun_out/r2c5_ncu2.log
