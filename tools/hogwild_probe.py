import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import graphvite_b200 as gv
from graphvite_b200 import _lib, datasets
import bench
u, v = datasets.power_law_edges(20000, 200000, seed=5)
path = "/tmp/mid.txt"
datasets.write_edge_list(path, u, v)
train = dict(num_epoch=100, augmentation_step=2, random_walk_length=10, random_walk_batch_size=20)
graph = gv.graph.Graph(); graph.load(path)
for flags in (0, 512, 0, 512):
    _lib.lib.gv_cuda_set_tunable(b"kernel_flags", flags)
    _lib.lib.gv_reset_global_engine(5489)
    solver = gv.solver.GraphSolver(128, device_ids=[0])
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_negative=1, batch_size=10000, episode_size=50)
    solver.train("LINE", **train)
    print("ours flags", flags, float(np.linalg.norm(solver.vertex_embeddings)), float(np.linalg.norm(solver.context_embeddings)), flush=True)
    solver.close()
ref = bench.load_reference()
for i in range(3):
    rgraph = ref.graph.Graph_j(); rgraph.load(path, True, False)
    rsolver = ref.solver.GraphSolver_128_f_j([0], 4, 0)
    rsolver.build(rgraph, ref.optimizer.SGD(0.025, 0.005), 0, 1, 10000, 50)
    rsolver.train(model="LINE", log_frequency=1 << 30, **train)
    print("reference", float(np.linalg.norm(np.array(rsolver.vertex_embeddings))), float(np.linalg.norm(np.array(rsolver.context_embeddings))), flush=True)
