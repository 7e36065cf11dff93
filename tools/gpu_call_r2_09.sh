#!/usr/bin/env bash
# round 2, GPU call 9: ncu --set full of the tc5h kernels (first 16->16 layer and output1 48->16)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc5h_conv_kernel -c 14 -o gpurun_out/r2c9_tc5h \
    python tools/profile_forward.py 1 > gpurun_out/r2c9_ncu.log 2>&1
tail -2 gpurun_out/r2c9_ncu.log
