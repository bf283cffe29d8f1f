#!/bin/bash
# round 2, GPU call 9 (8 GPUs): bench at N = 8 with the parity self-check, RotatE at N = 4, the reference arm at N = 8 (stderr kept)
set -u
mkdir -p gpurun_out
port=$((29500 + RANDOM % 400))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 8 --steps 16 --warmup 8 > gpurun_out/c9_bench_n8.json 2> gpurun_out/c9_bench_n8.err
tail -c 300 gpurun_out/c9_bench_n8.err
port=$((29500 + RANDOM % 400))
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 4 --workload rotate_fb15k237 --steps 6 --warmup 2 > gpurun_out/c9_bench_rotate_n4.json 2> gpurun_out/c9_bench_rotate_n4.err
tail -c 300 gpurun_out/c9_bench_rotate_n4.err
port=$((29500 + RANDOM % 400))
GV_KERNEL_FLAGS=64 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 8 --steps 16 --warmup 8 --no-e2e --no-parity > gpurun_out/c9_bench_n8_persistent.json 2> gpurun_out/c9_bench_n8_persistent.err
timeout 240 python bench.py --impl reference --gpus 8 --steps 16 --warmup 8 > gpurun_out/c9_bench_reference_n8.json 2> gpurun_out/c9_bench_reference_n8.err
tail -c 600 gpurun_out/reference_stderr_n8.log gpurun_out/c9_bench_reference_n8.json
