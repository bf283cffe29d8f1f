"""Throughput of the knowledge-graph path at config #4 of BASELINE.json (RotatE d=2048 on an FB15k-237-shaped
graph: 14 541 entities, 237 relations, 272 115 triplets; Adam, k = 64, B = 1e5, episode_size = 1, margin 9,
adversarial temperature 2 -- config/knowledge_graph/rotate_fb15k-237.yaml).  Not the headline metric (that is
bench.py); this is the measurement tool for `kg_train_kernel`:

    python tools/bench_kg.py [--dim 2048] [--episodes 8] [--warmup 2] [--model RotatE] [--negatives 64]

prints one JSON line: positives/s over whole episodes (sampler overlapped), the train kernel's CUDA-event time,
and its algorithmic bytes per positive against the measured HBM peak (the positive head /
tail / relation rows stay in registers, negative rows are read once for the normaliser and once more
read-modify-written with their moments).  Under ncu: add --episodes 1 --warmup 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def algorithmic_bytes(dim, k, num_moment, model, adversarial):
    row = dim * 4
    relation_row = row // 2 if model == "RotatE" else row
    states = 1 + num_moment
    pass1 = k * row if adversarial else 0                  # negative rows for the normaliser
    pass2 = k * states * 2 * row                           # negative rows (+ moments), read and written
    cached = states * 2 * (2 * row + relation_row)         # positive head, tail and the relation, once each way
    return pass1 + pass2 + cached + 12 + 8 * k + 4         # + the triplet, the negatives' randoms, the loss


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--dim", type=int, default=2048)
    parser.add_argument("--model", default="RotatE")
    parser.add_argument("--negatives", type=int, default=64)
    parser.add_argument("--batch", type=int, default=100000)
    parser.add_argument("--episodes", type=int, default=8)
    parser.add_argument("--warmup", type=int, default=2)
    parser.add_argument("--optimizer", default="Adam")
    args = parser.parse_args()

    import graphvite_b200 as gv
    from graphvite_b200 import _lib, datasets
    path = "/tmp/gv_b200_fb15k237.txt"
    if not os.path.exists(path):
        datasets.synthetic_knowledge_graph_file("fb15k-237", path)
    graph = gv.graph.KnowledgeGraph()
    graph.load(path)
    solver = gv.solver.KnowledgeGraphSolver(args.dim, device_ids=[0])
    optimizer = getattr(gv.optimizer, args.optimizer)(2e-6, 0)
    solver.build(graph, optimizer, num_negative=args.negatives, batch_size=args.batch, episode_size=1)
    lib, handle = _lib.lib, solver._handle
    temperature, margin = 2.0, 9.0
    _lib.check(lib.gv_kg_solver_train_begin(handle, args.model.encode(), 1000, 0, 1.0, margin, 2e-3, 2000, 1,
                                            temperature, 100))
    for _ in range(args.warmup):
        assert lib.gv_kg_solver_train_episode(handle) == 1, _lib.last_error()
    before = solver.stats
    start = time.time()
    for _ in range(args.episodes):
        assert lib.gv_kg_solver_train_episode(handle) == 1, _lib.last_error()
    seconds = time.time() - start
    after = solver.stats
    _lib.check(lib.gv_kg_solver_train_end(handle))
    positives = after["positives"] - before["positives"]
    kernel_seconds = after["kernel_seconds"] - before["kernel_seconds"]
    num_moment = {"SGD": 0, "Adam": 2}.get(args.optimizer, 1)
    per_positive = algorithmic_bytes(args.dim, args.negatives, num_moment, args.model, temperature > 0)
    peak = 6650.0
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        peak = float(json.load(open(peaks))["hbm_gbs"])
    achieved = positives * per_positive / max(kernel_seconds, 1e-12) / 1e9
    print(json.dumps({
        "metric": "positive triplets/s, %s d=%d on an FB15k-237-shaped graph" % (args.model, args.dim),
        "value": positives / seconds, "unit": "triplets/s", "episodes": args.episodes,
        "ms_per_batch": seconds / max(1, positives / args.batch) * 1e3,
        "kernel": {"name": "gv::device::kg_train_kernel", "seconds": kernel_seconds,
                   "positives_per_s": positives / max(kernel_seconds, 1e-12),
                   "algorithmic_bytes_per_positive": per_positive, "achieved_gbs": achieved, "hbm_peak_gbs": peak,
                   "frac_of_hbm_peak": achieved / peak,
                   "note": "entity blocks + moments of one GPU are ~357 MB at P = 1: the negatives' rows hit L2 "
                           "(126 MB) only partly; compare with dram__bytes from ncu"},
        "config": {"entities": graph.num_vertex, "relations": graph.num_relation, "triplets": graph.num_edge,
                   "num_negative": args.negatives, "batch_size": args.batch, "optimizer": args.optimizer,
                   "num_partition": solver.num_partition},
        "entity_norm": float((solver.entity_embeddings ** 2).sum() ** 0.5)}), flush=True)


if __name__ == "__main__":
    main()
