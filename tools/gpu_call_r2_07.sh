#!/usr/bin/env bash
# round 2, GPU call 7: suite, default bench (all legs), config-4 stress bench, fused head ncu
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/r2c7_tests.log 2>&1
echo "suite rc=$?"; grep -E "head iter|passed|failed|cfg5|Error" gpurun_out/r2c7_tests.log | tail -8
timeout 600 python bench.py --steps 40 --warmup 5 --breakdown > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
tail -2 gpurun_out/r2c7_bench.err
timeout 900 python bench.py --config 4 --steps 20 --warmup 3 --breakdown > gpurun_out/r2c7_bench_cfg4.json 2> gpurun_out/r2c7_bench_cfg4.err
tail -2 gpurun_out/r2c7_bench_cfg4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 2 -c 1 -o gpurun_out/r2c7_headfused \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-u8 > gpurun_out/r2c7_ncu2.log 2>&1
