#!/bin/bash
# round 2, GPU call 6 (1 GPU): resident-groups kernel -- parity vs the reference as a function of the resident block count, speed
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "resident or timeline" 2>&1 | tail -3 > gpurun_out/c6_tests.txt
S=shipped,rg_592,rg_592_noprefetch,rg_560,rg_520,rg_480,rg_444,rg_400
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 2 --reference-repeat 2 --settings $S --out gpurun_out/parity6_youtube.jsonl > gpurun_out/c6_youtube.log 2>&1
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 3 --reference-repeat 3 --settings $S --out gpurun_out/parity6_blogcatalog.jsonl > gpurun_out/c6_blog.log 2>&1
timeout 600 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 1 --reference-repeat 0 --partitions 8 --settings shipped,rg_592,rg_480 \
    --out gpurun_out/parity6_youtube_p8.jsonl > gpurun_out/c6_youtube_p8.log 2>&1
GV_KERNEL_FLAGS=1024 GV_SAMPLE_PREFETCH_BLOCKS=1 timeout 900 python bench.py --workload friendster_lite --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c6_bench_friendster_lite_rg.json 2> gpurun_out/c6_bench_friendster_lite_rg.err
grep summary gpurun_out/parity6_*.jsonl | cut -c1-400
