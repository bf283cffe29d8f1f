#!/bin/bash
# round 2, GPU call 3 (1 GPU): which concurrency of the one-warp-per-sample kernel reproduces the reference's race statistics
set -u
mkdir -p gpurun_out
S=per_sample,ps_w60,ps_w54,ps_w48,ps_w44,ps_serial_w64,ps_serial_w54,ps_serial_w48,ps_l2only_w64,ps_l2only_w48
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 3 --reference-repeat 4 --settings $S --out gpurun_out/parity3_blogcatalog.jsonl > gpurun_out/c3_blog.log 2>&1
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 2 --reference-repeat 3 --settings $S --out gpurun_out/parity3_youtube.jsonl > gpurun_out/c3_youtube.log 2>&1
# durations of the reference's train kernel and of ours, per batch of 1e5 samples (serialised, cold cache)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:train -c 400 --csv --log-file gpurun_out/c3_train_durations.csv \
    python tools/parity_sweep.py --workload youtube --epochs 3 --repeat 1 --reference-repeat 1 --settings per_sample > gpurun_out/c3_ncu.log 2>&1
grep summary gpurun_out/parity3_*.jsonl | cut -c1-400
