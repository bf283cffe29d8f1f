#!/bin/bash
# round 2, GPU call 11 (2 GPUs): bench at N = 2 with the progressive parity self-check; ncu of kg_train_kernel
set -u
mkdir -p gpurun_out
port=$((29500 + RANDOM % 400))
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c11_bench_n2.json 2> gpurun_out/c11_bench_n2.err
tail -c 300 gpurun_out/c11_bench_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:kg_train_kernel -s 1 -c 1 -f -o gpurun_out/r02_kg_train \
    python bench.py --workload rotate_fb15k237 --steps 1 --warmup 1 > gpurun_out/c11_ncu_kg.log 2>&1
tail -3 gpurun_out/c11_ncu_kg.log | cut -c1-200
