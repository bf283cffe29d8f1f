#!/bin/bash
# round 2, GPU call 1 (1 GPU): box facts, new kernel flags on hardware, float-parity sweep vs the reference
set -u
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; } > gpurun_out/box.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/c1_kernels.txt
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 3 --out gpurun_out/parity_blogcatalog.jsonl > gpurun_out/c1_blog.log 2>&1
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 2 --out gpurun_out/parity_youtube.jsonl > gpurun_out/c1_youtube.log 2>&1
timeout 600 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 1 --reference-repeat 0 --partitions 8 \
    --settings r1_shipped,l2_only,hot1024,hot8192,all_l1_chunk16,all_l1_wb_chunk16,interleaved_l2,interleaved_hot128,interleaved_hot1024,interleaved_all_l1_wb_chunk16,interleaved_all_l1_wb_chunk1 \
    --out gpurun_out/parity_youtube_p8.jsonl > gpurun_out/c1_youtube_p8.log 2>&1
grep summary gpurun_out/parity_*.jsonl | cut -c1-400
