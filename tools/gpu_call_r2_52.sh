#!/usr/bin/env bash
# round 2, GPU call 52: phase-A prefetch of the records' lines (IMVS_TUNE_WC_PF / WCI_PF: 1 = L1, 2 = L2)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c52_ps.jsonl; : > $O
timeout 120 python tools/bench_planesweep.py --tag base --save /tmp/ps_ref.pt >> $O 2>gpurun_out/r2c52_ps.err
for m in 1 2; do IMVS_TUNE_WC_PF=$m IMVS_TUNE_WCI_PF=$m timeout 120 python tools/bench_planesweep.py --tag pf$m --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c52_ps.err; done
timeout 120 python tools/bench_planesweep.py --tag base_again --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c52_ps.err
IMVS_TUNE_WC_PF=1 IMVS_TUNE_WCI_PF=1 timeout 120 python tools/bench_planesweep.py --tag pf1_again --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c52_ps.err
IMVS_TUNE_WC_PF=1 IMVS_TUNE_WCI_PF=1 timeout 200 python tools/bench_planesweep.py --config 5 --tag pf1_cfg5 >> $O 2>>gpurun_out/r2c52_ps.err
timeout 200 python tools/bench_planesweep.py --config 5 --tag base_cfg5 >> $O 2>>gpurun_out/r2c52_ps.err
python - <<'PY'
import json
for l in open("gpurun_out/r2c52_ps.jsonl"):
    d = json.loads(l)
    print(d["tag"], "iter", round(d["iter_warm"]["median_us"], 1), round(d["iter_cold"]["median_us"], 1), "init", round(d["init_warm"]["median_us"], 1),
          round(d["init_cold"]["median_us"], 1), d.get("max_abs_diff_vs_ref"))
PY
timeout 600 python tools/tune_sweep.py "K8=1" "WC_PF=1,WCI_PF=1" "K8=1" "WC_PF=1,WCI_PF=1" 2>&1 | tee gpurun_out/r2c52_sweep.txt
