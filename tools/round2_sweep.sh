#!/bin/bash
# Tunable sweep at N GPUs for the multi-GPU efficiency question left open in round 1 (inside one
# block of the 2-D partition the hub rows are P x hotter, and the samplers' kernels compete with a persistent train
# grid).  Run on an N-GPU box:   gpurun --gpus 2 --timeout 2400 -- 'bash tools/round2_sweep.sh 2'
# (about 1 minute per line; the lines are ordered by how much they are expected to tell, and every finished line is
# already in the output file if the call is cut short; N = 1 answers the sampler / train co-scheduling questions alone)
# Every line of gpurun_out/sweep_n<N>.jsonl is one bench.py result (value = edges/s over all ranks) tagged with the
# environment it ran under.
set -u
N=${1:-2}
OUT=gpurun_out/sweep_n${N}.jsonl
mkdir -p gpurun_out
: > $OUT
run() {
    local tag="$1"; shift
    local port=$((29500 + RANDOM % 400))
    local line
    line=$(env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus $N --steps 8 --warmup 3 --no-e2e 2>/dev/null | tail -1)
    echo "{\"tag\": \"$tag\", \"result\": ${line:-null}}" | tee -a $OUT
}
run "default" GV_LOG=0
run "direct_peer_scatter" GV_DIRECT_PEER_SCATTER=1
run "fill_per_walk (round-1 fill kernels)" GV_FILL_PER_WALK=1
run "rng_per_chunk (round-1 generator calls)" GV_RNG_PER_CHUNK=1 GV_RNG_SEQUENTIAL=1
run "dynamic_chunks" GV_KERNEL_FLAGS=8
run "dynamic_chunks,sampler_max_ctas=64" GV_KERNEL_FLAGS=8 GV_SAMPLER_MAX_CTAS=64
run "dynamic_chunks,reserve_sms=8" GV_KERNEL_FLAGS=8 GV_TRAIN_RESERVE_SMS=8
run "dynamic_chunks,sampler_max_ctas=64,reserve_sms=8" GV_KERNEL_FLAGS=8 GV_SAMPLER_MAX_CTAS=64 GV_TRAIN_RESERVE_SMS=8
run "blocks_per_sm=3" GV_TRAIN_BLOCKS_PER_SM=3
run "dynamic_chunks,blocks_per_sm=3" GV_KERNEL_FLAGS=8 GV_TRAIN_BLOCKS_PER_SM=3
run "chunk_batches=32" GV_CHUNK_BATCHES=32
run "chunk_batches=8" GV_CHUNK_BATCHES=8
run "replicated_sampling" GV_REPLICATED_SAMPLING=1
run "sampler_max_ctas=64" GV_SAMPLER_MAX_CTAS=64
run "reserve_sms=8" GV_TRAIN_RESERVE_SMS=8
run "reserve_sms=16" GV_TRAIN_RESERVE_SMS=16
for hot in 256 512 1024 4096; do
    run "hot_rows=$hot" GV_HOT_ROWS=$hot
done
run "blocks_per_sm=3,hot_rows=512" GV_TRAIN_BLOCKS_PER_SM=3 GV_HOT_ROWS=512
python - "$OUT" <<'PY'
import json, sys
print("%-52s %12s %10s %8s" % ("tag", "edges/s", "ms/step", "frac"))
for line in open(sys.argv[1]):
    row = json.loads(line)
    r = row["result"]
    if r:
        print("%-52s %12.4g %10.2f %8.3f" % (row["tag"], r["value"], r["ms_per_step"], r["frac"]))
    else:
        print("%-52s failed" % row["tag"])
PY
