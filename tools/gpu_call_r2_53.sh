#!/usr/bin/env bash
# round 2, GPU call 53: level 3 of the plane sweep from a pyramid padded to 256 bytes per texel (IMVS_TUNE_WC_PAD3=1)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c53_ps.jsonl; : > $O
timeout 120 python tools/bench_planesweep.py --tag base --save /tmp/ps_ref.pt >> $O 2>gpurun_out/r2c53_ps.err
timeout 120 python tools/bench_planesweep.py --tag pad3 --pad3 --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c53_ps.err
timeout 120 python tools/bench_planesweep.py --tag base_again --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c53_ps.err
timeout 120 python tools/bench_planesweep.py --tag pad3_again --pad3 --check /tmp/ps_ref.pt >> $O 2>>gpurun_out/r2c53_ps.err
timeout 200 python tools/bench_planesweep.py --config 5 --tag base_cfg5 >> $O 2>>gpurun_out/r2c53_ps.err
timeout 200 python tools/bench_planesweep.py --config 5 --tag pad3_cfg5 --pad3 >> $O 2>>gpurun_out/r2c53_ps.err
python - <<'PY'
import json
for l in open("gpurun_out/r2c53_ps.jsonl"):
    d = json.loads(l)
    print(d["tag"], "iter", round(d["iter_warm"]["median_us"], 1), round(d["iter_cold"]["median_us"], 1), "init", round(d["init_warm"]["median_us"], 1),
          round(d["init_cold"]["median_us"], 1), "pad", d.get("pad_warm", {}).get("median_us"), d.get("max_abs_diff_vs_ref"), d.get("ref_abs_max"))
PY
tail -3 gpurun_out/r2c53_ps.err
timeout 600 python tools/tune_sweep.py "K8=1" "WC_PAD3=1" "K8=1" "WC_PAD3=1" 2>&1 | tee gpurun_out/r2c53_sweep.txt
IMVS_TUNE_WC_PAD3=1 timeout 600 python -m pytest tests -q -m gpu -k "pipeline or cfg2 or cfg5 or full_size or size_independent or streaming or uint8" > gpurun_out/r2c53_tests_pad3.log 2>&1
echo "pad3 tests rc=$?"; tail -3 gpurun_out/r2c53_tests_pad3.log
