#!/bin/bash
# round 2, GPU call 10 (1 GPU): GPU test files touched this round, the bench (both arms), ncu launch list + full captures at HEAD
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py tests/test_gpu_x_solver_more.py tests/test_gpu_x_pybind.py \
    tests/test_gpu_y_fill.py tests/test_gpu_yy_later_kernels.py tests/test_gpu_zzz_blogcatalog.py tests/test_gpu_zzz_full_size.py \
    tests/test_gpu_zzzzz_parity.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/c10_tests.txt
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/c10_bench_reference.json 2> gpurun_out/c10_bench_reference.err
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c10_ncu_launches.log 2>&1
for P in 1 8; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:train_sample_per_warp -s 300 -c 1 -f -o gpurun_out/r02_train_p$P \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --partitions $P > gpurun_out/c10_ncu_p$P.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kg_train_kernel -s 2 -c 1 -f -o gpurun_out/r02_kg_train \
    python bench.py --workload rotate_fb15k237 --steps 1 --warmup 1 > gpurun_out/c10_ncu_kg.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"fill_direct_tiled|random_walk_kernel|rng_generate_segmented|sample_negatives" -c 4 -f -o gpurun_out/r02_sampler_kernels \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c10_ncu_sampler.log 2>&1
cat gpurun_out/c10_tests.txt | tail -5
