#!/bin/bash
# round 2, GPU call 7 (2 GPUs): multi-rank parity tests, bench at N = 2 (with the parity self-check), sampler settings
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_w_multi.py -x -q -m gpu -k "2-" 2>&1 | tail -5 > gpurun_out/c7_multi_tests.txt
run() {
    local tag="$1"; shift
    local port=$((29500 + RANDOM % 400))
    env "$@" GV_LOG=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 $EXTRA > gpurun_out/c7_bench_$tag.json 2> gpurun_out/c7_bench_$tag.err
}
EXTRA="" run default
EXTRA="--no-e2e --no-parity" run direct_scatter GV_DIRECT_PEER_SCATTER=1
EXTRA="--no-e2e --no-parity" run persistent_r1 GV_KERNEL_FLAGS=64 GV_CHUNK_BATCHES=16
tail -c 400 gpurun_out/c7_bench_*.err
