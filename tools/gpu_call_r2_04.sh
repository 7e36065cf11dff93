#!/usr/bin/env bash
# round 2, GPU call 4: hybrid band/direct init kernel: density threshold sweep at configs 2 and 5 + parity
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c4_ps.jsonl
: > $O
for dn in 0 2 3 4 6; do
  env IMVS_INIT_DENSE=$dn python tools/bench_planesweep.py --reps 15 --tag "dense$dn" >> $O 2>> gpurun_out/r2c4_ps.err
  env IMVS_INIT_DENSE=$dn python tools/bench_planesweep.py --reps 8 --config 5 --tag "cfg5_dense$dn" >> $O 2>> gpurun_out/r2c4_ps.err
done
python -m pytest tests/test_gpu_parity.py -q -x -k "init or fixture or cfg" > gpurun_out/r2c4_tests.log 2>&1
tail -3 gpurun_out/r2c4_tests.log
ncu --set full --clock-control none --import-source on -k regex:warpcorr_init_kernel -c 1 -o gpurun_out/r2c4_init \
    python tools/bench_planesweep.py --reps 1 --tag ncu > gpurun_out/r2c4_ncu.log 2>&1
