#!/usr/bin/env bash
# round 2, GPU call 3: band-based init kernel: timing + parity
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c3_ps.jsonl
: > $O
python tools/bench_planesweep.py --reps 20 --tag band_init --check gpurun_out/r2c2_ref.pt >> $O 2>> gpurun_out/r2c3_ps.err
python tools/bench_planesweep.py --reps 10 --config 5 --tag band_init_cfg5 >> $O 2>> gpurun_out/r2c3_ps.err
python tools/bench_planesweep.py --reps 10 --noise 0.05 --tag band_init_noise >> $O 2>> gpurun_out/r2c3_ps.err
python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -q -x -s > gpurun_out/r2c3_tests.log 2>&1
tail -3 gpurun_out/r2c3_tests.log
ncu --set full --clock-control none --import-source on -k regex:warpcorr_init_kernel -c 1 -o gpurun_out/r2c3_init \
    python tools/bench_planesweep.py --reps 1 --tag ncu > gpurun_out/r2c3_ncu.log 2>&1
