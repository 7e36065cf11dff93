#!/usr/bin/env bash
# round 2, GPU call 55: ncu of the padded-level-3 plane-sweep kernels (full set, one launch each) + the launch list of the final forward
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c55_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c55_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:warpcorr_iter_kernel -c 1 -o gpurun_out/r2c55_warpcorr_iter \
    python tools/profile_forward.py 1 > gpurun_out/r2c55_ncu2.log 2>&1
tail -1 gpurun_out/r2c55_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:warpcorr_init_kernel -c 1 -o gpurun_out/r2c55_warpcorr_init \
    python tools/profile_forward.py 1 > gpurun_out/r2c55_ncu3.log 2>&1
tail -1 gpurun_out/r2c55_ncu3.log
ls -la gpurun_out/r2c55*
