#!/usr/bin/env bash
# round 2, GPU call 43: ncu --set full of the tile-resident CorrNet kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
IMVS_TUNE_CORR_TILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:corrnet_tile -s 1 -c 1 -o gpurun_out/r2c43_ctile \
    python tools/profile_forward.py 1 > gpurun_out/r2c43_ncu.log 2>&1
tail -1 gpurun_out/r2c43_ncu.log
