#!/usr/bin/env bash
# round 2, GPU call 2: plane-sweep iteration kernel diagnostics (warm-L1 second pass per level) and region prefetch variants
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c2_ps.jsonl
: > $O
run() { env "$@" python tools/bench_planesweep.py --reps 20 --tag "$*" --check gpurun_out/r2c2_ref.pt >> $O 2>> gpurun_out/r2c2_ps.err; }
python tools/bench_planesweep.py --reps 20 --tag base --save gpurun_out/r2c2_ref.pt >> $O 2>> gpurun_out/r2c2_ps.err
for L in 1 2 4; do
  run IMVS_WC_LEVELS=$L
  run IMVS_WC_LEVELS=$L IMVS_WC_REPEAT=2
  run IMVS_WC_LEVELS=$L IMVS_WC_PF=1 IMVS_WC_PF_LEVELS=$L
  run IMVS_WC_LEVELS=$L IMVS_WC_PF=2 IMVS_WC_PF_LEVELS=$L
done
run IMVS_WC_PF=1
run IMVS_WC_PF=2
run IMVS_WC_PF=1 IMVS_WC_PF_LEVELS=6
run IMVS_WC_PF=2 IMVS_WC_PF_LEVELS=6
run IMVS_WC_PF=2 IMVS_WC_PF_LEVELS=4
run IMVS_WC_REPEAT=2
python tools/bench_planesweep.py --reps 10 --noise 0.05 --tag noise_base >> $O 2>> gpurun_out/r2c2_ps.err
env IMVS_WC_PF=2 python tools/bench_planesweep.py --reps 10 --noise 0.05 --tag noise_pf2 >> $O 2>> gpurun_out/r2c2_ps.err
python -m pytest tests/test_gpu_parity.py -q -x -k "evaluation or cfg5 or all_predictions" -s > gpurun_out/r2c2_tests.log 2>&1
env IMVS_WC_PF=2 python -m pytest tests/test_gpu_parity.py -q -x -k "evaluation or cfg2 or fixture" > gpurun_out/r2c2_tests_pf2.log 2>&1
tail -3 gpurun_out/r2c2_tests.log gpurun_out/r2c2_tests_pf2.log
