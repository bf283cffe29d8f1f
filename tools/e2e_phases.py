"""Phase timings of GraphSolver.train() at the Youtube benchmark size (GV_LOG=2 prints them)."""
import os, sys, time
os.environ.setdefault("GV_LOG", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import graphvite_b200 as gv

cfg = bench.WORKLOADS["youtube"]
path = bench.graph_file("youtube")
for attempt in range(2):
    t = time.time()
    graph = gv.graph.Graph(); graph.load(path)
    print("graph load %.3f" % (time.time() - t), flush=True); t = time.time()
    solver = gv.solver.GraphSolver(128, device_ids=[0])
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), batch_size=100000, episode_size=500)
    print("ctor+build %.3f" % (time.time() - t), flush=True); t = time.time()
    solver.train(**bench.train_kwargs(cfg, 100))
    print("train %.3f  stats %s" % (time.time() - t, solver.stats), flush=True)
    del solver
