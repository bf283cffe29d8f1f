#!/bin/bash
# 2 GPUs: the toy multi-rank parity worker in its three sampler delivery modes, full output kept
set -u
mkdir -p gpurun_out
run() {
    local tag="$1"; shift
    local port=$((29500 + RANDOM % 400))
    env "$@" GV_TEST_SOLVER=graph timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port $port tests/multi_rank_worker.py > gpurun_out/c8_worker_$tag.log 2>&1
    echo "$tag rc=$?" >> gpurun_out/c8_summary.txt
}
: > gpurun_out/c8_summary.txt
run staged GV_TEST_PARTITIONS=2
run direct GV_TEST_PARTITIONS=2 GV_DIRECT_PEER_SCATTER=1
run replicated GV_TEST_PARTITIONS=2 GV_REPLICATED_SAMPLING=1
run staged_p4 GV_TEST_PARTITIONS=4
run node2vec GV_TEST_PARTITIONS=2 GV_TEST_MODEL=node2vec
cat gpurun_out/c8_summary.txt
grep -h "Error\|error\|assert\|Mismatch\|rank . ok" gpurun_out/c8_worker_*.log | head -40
port=$((29500 + RANDOM % 400))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
tail -c 300 gpurun_out/c8_bench_n2.err
