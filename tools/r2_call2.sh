#!/bin/bash
# round 2, GPU call 2 (1 GPU): the one-warp-per-sample kernel -- parity vs the reference and speed (P=1 and inside a P=8 block)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/c2_tests.txt
S=per_sample,per_sample_chunk4,per_sample_chunk16,r1_shipped
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 4 --reference-repeat 4 --settings $S --out gpurun_out/parity2_blogcatalog.jsonl > gpurun_out/c2_blog.log 2>&1
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 3 --settings $S --out gpurun_out/parity2_youtube.jsonl > gpurun_out/c2_youtube.log 2>&1
timeout 600 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 1 --reference-repeat 0 --partitions 8 --settings $S \
    --out gpurun_out/parity2_youtube_p8.jsonl > gpurun_out/c2_youtube_p8.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
grep summary gpurun_out/parity2_*.jsonl | cut -c1-600
