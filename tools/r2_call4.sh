#!/bin/bash
# round 2, GPU call 4 (1 GPU): index prefetch variants of the one-warp-per-sample kernel; new GPU tests; ncu of the kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzz_blogcatalog.py tests/test_gpu_x_pybind.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/c4_tests.txt
S=per_sample,ps_pf592,ps_pf1184,ps_pf2368,ps_pf1184_serial
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 3 --reference-repeat 3 --settings $S --out gpurun_out/parity4_blogcatalog.jsonl > gpurun_out/c4_blog.log 2>&1
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 2 --reference-repeat 2 --settings $S --out gpurun_out/parity4_youtube.jsonl > gpurun_out/c4_youtube.log 2>&1
timeout 600 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 1 --reference-repeat 0 --partitions 8 --settings per_sample,ps_pf1184 \
    --out gpurun_out/parity4_youtube_p8.jsonl > gpurun_out/c4_youtube_p8.log 2>&1
# full ncu capture of one warm launch of the train kernel, at P = 1 and inside a block of P = 8
for P in 1 8; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_sample_per_warp -s 300 -c 1 -f -o gpurun_out/r02_train_p$P \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --partitions $P > gpurun_out/c4_ncu_p$P.log 2>&1
done
grep summary gpurun_out/parity4_*.jsonl | cut -c1-400
