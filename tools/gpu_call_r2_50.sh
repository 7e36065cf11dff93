#!/usr/bin/env bash
# round 2, GPU call 50: A/B/A/B of the call-49 switches that looked positive (alternating runs on one box)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
C="CORR_TILE05=1,CORR_TILEMID=1,WCI_VPER=2"
timeout 800 python tools/tune_sweep.py "K8=1" "$C" "K8=1" "$C" "K8=1" "$C" "CORR_TILE05=1,WCI_VPER=2" "CORR_TILE05=1,WCI_VPER=2" 2>&1 | tee gpurun_out/r2c50_sweep.txt
