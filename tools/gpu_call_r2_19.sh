#!/usr/bin/env bash
# round 2, GPU call 19: ncu --set full with dense stall sampling of the tc5p kernels (first 3 + output1 + GRU)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on --profile-from-start off -k regex:tc5p_conv_kernel -c 20 -o gpurun_out/r2c19_tc5p \
    python tools/profile_forward.py 1 > gpurun_out/r2c19_ncu2.log 2>&1
tail -1 gpurun_out/r2c19_ncu2.log
