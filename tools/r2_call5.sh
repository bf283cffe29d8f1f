#!/bin/bash
# round 2, GPU call 5 (1 GPU): timeline variant vs the reference; the new bench.py end to end (both arms); other workloads
set -u
mkdir -p gpurun_out
S=shipped,ps_timeline
timeout 900 python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --repeat 4 --reference-repeat 4 --settings $S --out gpurun_out/parity5_blogcatalog.jsonl > gpurun_out/c5_blog.log 2>&1
timeout 1200 python tools/parity_sweep.py --workload youtube --epochs 100 --repeat 3 --reference-repeat 3 --settings $S --out gpurun_out/parity5_youtube.jsonl > gpurun_out/c5_youtube.log 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/c5_bench_reference.json 2> gpurun_out/c5_bench_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
timeout 600 python bench.py --workload rotate_fb15k237 --steps 6 --warmup 2 > gpurun_out/c5_bench_rotate.json 2> gpurun_out/c5_bench_rotate.err
timeout 900 python bench.py --workload node2vec_youtube --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/c5_bench_node2vec.json 2> gpurun_out/c5_bench_node2vec.err
timeout 900 python bench.py --workload friendster_lite --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c5_bench_friendster_lite.json 2> gpurun_out/c5_bench_friendster_lite.err
grep summary gpurun_out/parity5_*.jsonl | cut -c1-400
tail -c 600 gpurun_out/c5_bench*.err
