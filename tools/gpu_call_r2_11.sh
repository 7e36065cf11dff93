#!/usr/bin/env bash
# round 2, GPU call 11 (2 GPUs): suite on GPU 0, inference + training bench at N = 1 and N = 2
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2c11_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c11_tests.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c11_bench_n1.json 2> gpurun_out/r2c11_bench_n1.err
tail -1 gpurun_out/r2c11_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2c11_bench_n2.json 2> gpurun_out/r2c11_bench_n2.err
tail -2 gpurun_out/r2c11_bench_n2.err
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/r2c11_train_n1.json 2> gpurun_out/r2c11_train_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --mode train --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c11_train_n2.json 2> gpurun_out/r2c11_train_n2.err
tail -2 gpurun_out/r2c11_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --mode train --train-bf16-wire --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c11_train_n2_bf16.json 2> gpurun_out/r2c11_train_n2_bf16.err
tail -2 gpurun_out/r2c11_train_n2_bf16.err
