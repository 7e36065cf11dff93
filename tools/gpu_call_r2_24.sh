#!/usr/bin/env bash
# round 2, GPU call 24: occupancy variants of warpcorr_iter_kernel (24 warps x 1 block, 32 x 1, 16 x 2 per SM)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/r2c24_ps.jsonl
for v in 24 32 16; do
  IMVS_TUNE_WC_WARPS=$v timeout 200 python tools/bench_planesweep.py --config 2 --reps 30 --tag warps$v --save gpurun_out/r2c24_ref_$v.pt >> gpurun_out/r2c24_ps.jsonl 2>> gpurun_out/r2c24_ps.err
  IMVS_TUNE_WC_WARPS=$v timeout 200 python tools/bench_planesweep.py --config 5 --reps 10 --tag cfg5_warps$v >> gpurun_out/r2c24_ps.jsonl 2>> gpurun_out/r2c24_ps.err
done
IMVS_TUNE_WC_WARPS=32 timeout 200 python tools/bench_planesweep.py --config 2 --reps 5 --tag check32 --check gpurun_out/r2c24_ref_24.pt >> gpurun_out/r2c24_ps.jsonl 2>> gpurun_out/r2c24_ps.err
IMVS_TUNE_WC_WARPS=16 timeout 200 python tools/bench_planesweep.py --config 2 --reps 5 --tag check16 --check gpurun_out/r2c24_ref_24.pt >> gpurun_out/r2c24_ps.jsonl 2>> gpurun_out/r2c24_ps.err
cat gpurun_out/r2c24_ps.jsonl | cut -c1-400
for v in 32 16; do
  IMVS_TUNE_WC_WARPS=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-u8 > gpurun_out/r2c24_bench_w$v.json 2> gpurun_out/r2c24_bench_w$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c24_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["single_stream"]["value"], 1), d["stage_ms"]["warpcorr_iter"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "unreadable", e)
PY
rm -f gpurun_out/r2c24_ref_*.pt
