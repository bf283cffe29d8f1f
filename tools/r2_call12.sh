#!/bin/bash
# round 2, GPU call 12 (1 GPU, the round's last 15 GPU-minutes): the knowledge-graph kernel after its instruction diet
# -- parity on real MUFU units, A/B of the kg_flags variants -- then the default bench (train() phases on stderr)
set -u
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_zz_kg_kernels.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/c12_kg_kernel_tests.txt
for F in 0 1 4 2; do
  GV_KG_FLAGS=$F timeout 100 python bench.py --workload rotate_fb15k237 --steps 6 --warmup 2 \
      > gpurun_out/c12_rotate_flags$F.json 2> gpurun_out/c12_rotate_flags$F.err
done
GV_LOG=2 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
( timeout 200 python -m pytest tests/test_host_runtime.py tests/test_gpu_zzz_full_size.py -x -q -k "engine or partition_initialisation" 2>&1 | tail -4 ) > gpurun_out/c12_engine_tests.txt
( timeout 200 python -m pytest tests/test_gpu_zz_kg_solver.py tests/test_gpu_solver.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/c12_solver_tests.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:kg_train_kernel -s 2 -c 1 -f -o gpurun_out/r02b_kg_train \
    python bench.py --workload rotate_fb15k237 --steps 1 --warmup 1 > gpurun_out/c12_ncu_kg.log 2>&1
cat gpurun_out/c12_kg_kernel_tests.txt gpurun_out/c12_engine_tests.txt gpurun_out/c12_solver_tests.txt
for F in 0 1 4 2; do python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/c12_rotate_flags$F.json").read().strip().splitlines()[-1])
    print("kg_flags $F:", "%.3e" % r["value"], "triplets/s, roofline frac %.3f" % r["roofline"]["frac"])
except Exception as e:
    print("kg_flags $F: no result", e)
PY
done
tail -c 600 gpurun_out/c12_bench.json
