#!/usr/bin/env bash
# round 2, GPU call 8: tcgen05 fp16-split convolutions in FeatureNet: parity + bench A/B + launch list
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "featurenet or cfg2 or fixture" > gpurun_out/r2c8_t1.log 2>&1
echo "t1 rc=$?"; grep -E "passed|failed|Error|error|cfg2" gpurun_out/r2c8_t1.log | tail -6
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c8_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c8_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --breakdown > gpurun_out/r2c8_bench_tc5h.json 2> gpurun_out/r2c8_bench_tc5h.err
tail -1 gpurun_out/r2c8_bench_tc5h.err
IMVS_TUNE_TC5H=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 --breakdown > gpurun_out/r2c8_bench_mma.json 2> gpurun_out/r2c8_bench_mma.err
tail -1 gpurun_out/r2c8_bench_mma.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c8_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c8_ncu1.log 2>&1
tail -2 gpurun_out/r2c8_ncu1.log
