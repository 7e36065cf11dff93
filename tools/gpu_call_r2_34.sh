#!/usr/bin/env bash
# round 2, GPU call 34 (8 GPUs, final tree): inference bench at N = 8 (value, e2e with fp32 and with uint8 image upload)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2c34_bench_n8.json 2> gpurun_out/r2c34_bench_n8.err
tail -1 gpurun_out/r2c34_bench_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d.get('e2e_uint8_images',{}).get('value'), d['single_stream']['value'])"
nvidia-smi topo -m > gpurun_out/r2c34_topo.txt 2>&1; nproc >> gpurun_out/r2c34_topo.txt
