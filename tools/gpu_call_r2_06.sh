#!/usr/bin/env bash
# round 2, GPU call 6: fused head v2 (2 CTAs/SM): test, bench, ncu
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "fused_tcgen05_head or cfg2 or fixture" > gpurun_out/r2c6_head_test.log 2>&1
echo "head test rc=$?"; grep -E "head iter|passed|failed" gpurun_out/r2c6_head_test.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --breakdown > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
cat gpurun_out/r2c6_bench.err | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_fused -s 2 -c 1 -o gpurun_out/r2c6_headfused \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c6_ncu2.log 2>&1
