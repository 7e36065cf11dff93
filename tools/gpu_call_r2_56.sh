#!/usr/bin/env bash
# round 2, GPU call 56: tile shapes of the PixelViewWeight convolution (5 120 CTAs of 8 rows by default)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 800 python tools/tune_sweep.py "K8=1" "PVW_TILE=1" "PVW_TILE=2" "PVW_TILE=3" "K8=1" "PVW_TILE=1" "PVW_TILE=3" 2>&1 | tee gpurun_out/r2c56_sweep.txt
