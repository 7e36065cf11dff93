#!/usr/bin/env bash
# round 2, GPU call 27 (2 GPUs, final tree): inference bench at N = 2, training bench at N = 2, reference arm under torchrun
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2c27_bench_n2.json 2> gpurun_out/r2c27_bench_n2.err
tail -1 gpurun_out/r2c27_bench_n2.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --mode train --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c27_train_n2.json 2> gpurun_out/r2c27_train_n2.err
tail -1 gpurun_out/r2c27_train_n2.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2c27_ref_n2.json 2> gpurun_out/r2c27_ref_n2.err
tail -1 gpurun_out/r2c27_ref_n2.json | cut -c1-300
