#!/bin/bash
# First GPU call of the next round: everything that round 1 could only execute under the CUDA emulation, on real
# hardware, plus the profiles that are still missing.  Run from the repo root on the GPU box:
#     gpurun --timeout 2400 -- 'bash tools/round2_gpu.sh'
# Outputs go to gpurun_out/ (merged back); copy what should be judged into profiles/ afterwards.
set -u
mkdir -p gpurun_out
OUT=gpurun_out

echo "== 1. GPU test-suite (validated files first, emulation-only files next, full size last)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt

echo "== 2. headline bench (N = 1)"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
tail -c 1500 $OUT/bench_n1.json

echo "== 3. knowledge-graph golden vectors from the reference ON THE DEVICE (device rounding; compare with the"
echo "      emulation-recorded fixtures in tests/golden/kg_*.npz)"
timeout 600 python oracle/make_golden_kg.py $OUT/golden_kg > $OUT/golden_kg.log 2>&1
tail -3 $OUT/golden_kg.log

echo "== 4. knowledge-graph throughput (RotatE d=2048, FB15k-237 shape) and its kernel under ncu"
timeout 600 python tools/bench_kg.py --episodes 8 --warmup 2 > $OUT/bench_kg.json 2> $OUT/bench_kg.err
tail -c 1200 $OUT/bench_kg.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kg_train_kernel -c 1 \
    -o $OUT/kg_train_kernel python tools/bench_kg.py --episodes 1 --warmup 0 > $OUT/ncu_kg.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/kg_launches.csv \
    python tools/bench_kg.py --episodes 2 --warmup 0 > /dev/null 2>&1

echo "== 5. launch list + full capture of the headline kernel (same commands as profiles/r01_*)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > $OUT/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:train_sgd_kernel -c 1 \
    -o $OUT/r02_train_kernel python bench.py --steps 1 --warmup 1 --no-e2e > $OUT/ncu_train.log 2>&1
ls -la $OUT
