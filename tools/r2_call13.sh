#!/bin/bash
# round 2, GPU call 13: why tests/test_gpu_solver.py::test_hogwild_training_matches_the_reference_statistically failed in call 12
# (three runs, full assertion text), and the new graph loader timed on the box
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 120 python -m pytest tests/test_gpu_solver.py -x -q -m gpu -k hogwild -s 2>&1 | grep -E "^ours|passed|failed|Error|assert|^E " | head -12
done > gpurun_out/c13_hogwild.txt 2>&1
GV_LOG=2 timeout 120 python - > gpurun_out/c13_load.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import bench
import graphvite_b200 as gv
path = bench.graph_file("youtube")
for attempt in range(3):
    graph = gv.graph.Graph()
    start = time.time()
    graph.load(path)
    print("Graph.load %.3f s (%d vertices, %d lines)" % (time.time() - start, graph.num_vertex, graph.num_edge), flush=True)
PY
cat gpurun_out/c13_hogwild.txt gpurun_out/c13_load.txt
# the default bench with train() phases (GV_LOG=2) after the set-up changes: staged CSR upload, device-built edge table
GV_LOG=2 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
grep "gv\]" gpurun_out/c13_bench.err | tail -14; tail -c 700 gpurun_out/c13_bench.json
