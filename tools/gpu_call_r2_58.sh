#!/usr/bin/env bash
# round 2, GPU call 58: op-level test of the padded entry points; tile shapes of hidden_init's and the upsampling net's 3x3 convolutions
# (40 / 160 CTAs with 864 dependent MMAs per warp by default)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "padded_level3 or launch_granularity" > gpurun_out/r2c58_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r2c58_tests.log
timeout 800 python tools/tune_sweep.py "K8=1" "HINIT_TILE=1" "HINIT_TILE=2" "UPS_TILE=1" "UPS_TILE=2" "K8=1" "HINIT_TILE=2,UPS_TILE=2" "HINIT_TILE=1,UPS_TILE=1" 2>&1 | tee gpurun_out/r2c58_sweep.txt
for t in "HINIT_TILE=2 UPS_TILE=2" "HINIT_TILE=1 UPS_TILE=1"; do
  env $(for kv in $t; do echo IMVS_TUNE_$kv; done) timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pipeline or cfg2 or hinit or upsample" > gpurun_out/r2c58_tests_tiles.log 2>&1
  echo "[$t] parity rc=$?"; tail -1 gpurun_out/r2c58_tests_tiles.log
done
