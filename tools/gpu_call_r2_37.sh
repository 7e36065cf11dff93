#!/usr/bin/env bash
# round 2, GPU call 37: ncu --set full of the two lateral (1x1 + bilinear) launches
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on --profile-from-start off -k regex:tc5p_conv_kernel -s 15 -c 4 -o gpurun_out/r2c37_lateral \
    python tools/profile_forward.py 1 > gpurun_out/r2c37_ncu.log 2>&1
tail -1 gpurun_out/r2c37_ncu.log
