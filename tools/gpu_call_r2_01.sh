#!/usr/bin/env bash
# round 2, GPU call 1: whole GPU suite (with printed parity numbers), headline bench with the real reference arms,
# reference arm, training bench at N=1 + launch list + one capture of the plane-sweep backward kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests -q -m gpu -s > gpurun_out/r2c1_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2c1_gpu_tests.log
python bench.py --steps 40 --warmup 5 --breakdown > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c1_bench_ref.json 2> gpurun_out/r2c1_bench_ref.err
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/r2c1_bench_train.json 2> gpurun_out/r2c1_bench_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c1_train_launches.csv \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/r2c1_train_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warpcorr_.*bwd -c 4 -o gpurun_out/r2c1_warpcorr_bwd \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/r2c1_bwd_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
ls -la gpurun_out | tail -12
