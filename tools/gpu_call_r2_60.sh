#!/usr/bin/env bash
# round 2, GPU call 60: the depth head's 3x3 on 4-row tiles x 16-channel output blocks (IMVS_TUNE_HEAD=3: 640-1 280 CTAs instead of 320-640)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 500 python tools/tune_sweep.py "K8=1" "HEAD=3" "K8=1" "HEAD=3" 2>&1 | tee gpurun_out/r2c60_sweep.txt
IMVS_TUNE_HEAD=3 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pipeline or cfg2 or heads or fused_tcgen05_head or window" > gpurun_out/r2c60_tests.log 2>&1
echo "HEAD=3 parity rc=$?"; tail -1 gpurun_out/r2c60_tests.log
