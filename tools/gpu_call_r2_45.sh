#!/usr/bin/env bash
# round 2, GPU call 45 (8 GPUs): training bench (config 4) at N = 8 and N = 4
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --mode train --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2c45_train_n8.json 2> gpurun_out/r2c45_train_n8.err
tail -1 gpurun_out/r2c45_train_n8.json | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --mode train --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2c45_train_n4.json 2> gpurun_out/r2c45_train_n4.err
tail -1 gpurun_out/r2c45_train_n4.json | cut -c1-200
