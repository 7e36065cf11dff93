#!/usr/bin/env bash
# round 2, GPU call 61: ncu launch list of one eager forward of the LAST tree (after the tile-shape changes of calls 56 / 58)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c61_launches.csv \
    python tools/profile_forward.py 1 > gpurun_out/r2c61_ncu1.log 2>&1
tail -1 gpurun_out/r2c61_ncu1.log; wc -l gpurun_out/r2c61_launches.csv
