#!/usr/bin/env bash
# round 2, GPU call 33: the fallback switches still give parity (featurenet vs oracle, cfg2 vs the reference fixture, GRU KAT)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in TC5P=0 TC5P_S2=0 TC5P_LAT=0 TC5P_GRU=0 HEADFUSED=0 TC5P_SHARE=2 TC5P_MB=1 TC5P_ST=2 CORR_FUSED=1 TC5P_CORR=1; do
  env IMVS_TUNE_$v timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "featurenet_vs_oracle or cfg2_matches or corrnet_pvw_gru or fused_tcgen05_head or fp16_range_guard" > gpurun_out/r2c33_$v.log 2>&1
  echo "$v rc=$? $(tail -1 gpurun_out/r2c33_$v.log)"
done
timeout 300 python -m pytest tests/test_fusion.py -q -m gpu 2>&1 | tail -1
