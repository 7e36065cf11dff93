#!/usr/bin/env bash
# First GPU call of round 2: everything that was written after round 1's GPU budget was spent, in the order that stops
# at the cheapest failure.  Run from the repository root through gpurun, e.g.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_gpu_call.sh'
# Outputs land in gpurun_out/r2_*; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

# 1. the on-device twins of the training tests (kernels so far executed through tests/cusim only)
python -m pytest tests/test_gpu_training.py -x -q -s > gpurun_out/r2_training_tests.log 2>&1
echo "training tests rc=$?" | tee -a gpurun_out/r2_training_tests.log

# 2. the whole GPU suite (regression check of the inference path: mmaconv.cuh gained #ifdef CUSIM branches)
python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log

# 3. headline bench (unchanged path) and the training line (BASELINE configs[3]) at N = 1
python bench.py --steps 40 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err

# 4. launch list of one training step (which kernels dominate: cuDNN convolutions vs the plane-sweep backward)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_train_launches.csv \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/r2_train_ncu.log 2>&1

# 5. one full capture of the plane-sweep backward kernels
ncu --set full --clock-control none --import-source on -k regex:warpcorr_.*bwd -c 4 -o gpurun_out/r2_warpcorr_bwd \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/r2_bwd_ncu.log 2>&1
ls -la gpurun_out | tail -12
