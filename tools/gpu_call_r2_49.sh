#!/usr/bin/env bash
# round 2, GPU call 49: tile-shape switches of CorrNet's mma.sync layers, view split of the init plane sweep,
# 26 / 28 warps in the iteration kernel (104 items per SM = 4 full rounds)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c49_ps.jsonl; : > $O
timeout 120 python tools/bench_planesweep.py --tag base --save /tmp/ps_ref.pt >> $O 2>gpurun_out/r2c49_ps.err
for w in 26 28; do IMVS_TUNE_WC_WARPS=$w timeout 120 python tools/bench_planesweep.py --tag warps$w --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c49_ps.err; done
for v in 1 2; do IMVS_TUNE_WCI_VPER=$v timeout 120 python tools/bench_planesweep.py --tag vper$v --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c49_ps.err; done
python - <<'PY'
import json
for l in open("gpurun_out/r2c49_ps.jsonl"):
    d = json.loads(l)
    print(d["tag"], "iter", round(d["iter_warm"]["median_us"], 1), round(d["iter_cold"]["median_us"], 1), "init", round(d["init_warm"]["median_us"], 1),
          round(d["init_cold"]["median_us"], 1), d.get("max_abs_diff_vs_ref"))
PY
timeout 600 python tools/tune_sweep.py "K8=1" "CORR_TILE05=1" "CORR_TILE05=2" "CORR_TILE05=3" "CORR_TILEMID=1" "WC_WARPS=26" "WCI_VPER=1" 2>&1 | tee gpurun_out/r2c49_sweep.txt
