#!/usr/bin/env bash
# round 2, GPU call 62: A/B/A/B of 16-row tiles for CorrNet's two full-resolution layers on the last tree (bit-identical results, call 51 / 59 tests)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python tools/tune_sweep.py "K8=1" "CORR_TILE05=1" "K8=1" "CORR_TILE05=1" 2>&1 | tee gpurun_out/r2c62_sweep.txt
