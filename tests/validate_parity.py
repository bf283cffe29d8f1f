#!/usr/bin/env python
"""Float parity at scale: train the same configuration with the reference (oracle/_ref/libgraphvite.so,
unmodified, through its pybind API) and with graphvite_b200, then compare embedding L2 norms,
link-prediction AUC on a held-out split (semantics of Dataset.link_prediction_split) and the
logged loss.  Runs on a GPU box:  python tests/validate_parity.py --workload blogcatalog --epochs 400
TEST / MEASUREMENT TOOLING (it executes oracle/_ref); not part of the product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def make_split(name, seed=20260922):
    from graphvite_b200 import datasets
    num_vertex, num_edge = datasets.SHAPES[name]
    u, v = datasets.power_law_edges(num_vertex, num_edge, seed=seed, max_degree=29000 if name == "youtube" else None)
    train_mask, tests = datasets.link_prediction_split(u, v, (100, 1, 1))
    path = "/tmp/gv_b200_%s_train.txt" % name
    if not os.path.exists(path):
        import pandas
        pandas.DataFrame({"u": u[train_mask], "v": v[train_mask]}).to_csv(path, sep="\t", header=False, index=False)
    return path, tests[0]


def auc_of(predict, name2id, test):
    from graphvite_b200.application import link_prediction_auc
    h, t, y = test
    keep = [i for i in range(len(h)) if str(h[i]) in name2id and str(t[i]) in name2id]
    pairs = np.array([[name2id[str(h[i])], name2id[str(t[i])]] for i in keep], dtype=np.uint32)
    return link_prediction_auc(predict(pairs), y[keep])


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--workload", default="blogcatalog")
    parser.add_argument("--epochs", type=int, default=400)
    parser.add_argument("--repeat", type=int, default=2)
    parser.add_argument("--skip-reference", action="store_true")
    args = parser.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    path, test = make_split(cfg["graph"])
    results = []

    import graphvite_b200 as gv
    for run in range(args.repeat):
        gv._clib.gv_reset_global_engine(5489)
        graph = gv.graph.Graph()
        graph.load(path)
        solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[0])
        solver.build(graph, gv.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), num_negative=cfg["num_negative"],
                     batch_size=cfg["batch_size"], episode_size=cfg["episode_size"])
        start = time.time()
        solver.train(**bench.train_kwargs(cfg, args.epochs))
        seconds = time.time() - start
        vertex, context = solver.vertex_embeddings, solver.context_embeddings
        scores = lambda pairs: np.einsum("ij,ij->i", vertex[pairs[:, 0]], context[pairs[:, 1]])
        results.append({"impl": "graphvite_b200", "run": run, "seconds": seconds, "batch_id": solver.batch_id,
                        "vertex_norm": float(np.linalg.norm(vertex)), "context_norm": float(np.linalg.norm(context)),
                        "auc": auc_of(scores, graph.name2id, test),
                        "auc_predict_kernel": auc_of(solver.predict, graph.name2id, test),
                        "loss": [float(x) for x in solver.logged_loss[-3:]]})
        print(json.dumps(results[-1]), flush=True)
        del solver

    if not args.skip_reference:
        ref = bench.load_reference()
        for run in range(args.repeat):
            graph = ref.graph.Graph_j()
            graph.load(path, True, False)
            solver = getattr(ref.solver, "GraphSolver_%d_f_j" % cfg["dim"])([0], 0, 0)
            solver.build(graph, ref.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), 0, cfg["num_negative"],
                         cfg["batch_size"], cfg["episode_size"])
            start = time.time()
            solver.train(model=cfg["model"], num_epoch=args.epochs, augmentation_step=cfg["augmentation_step"],
                         random_walk_length=cfg["random_walk_length"],
                         random_walk_batch_size=cfg["random_walk_batch_size"], negative_weight=cfg["negative_weight"],
                         log_frequency=1 << 30)
            seconds = time.time() - start
            vertex, context = np.array(solver.vertex_embeddings), np.array(solver.context_embeddings)
            scores = lambda pairs: np.einsum("ij,ij->i", vertex[pairs[:, 0]], context[pairs[:, 1]])
            results.append({"impl": "reference", "run": run, "seconds": seconds,
                            "vertex_norm": float(np.linalg.norm(vertex)),
                            "context_norm": float(np.linalg.norm(context)),
                            "auc": auc_of(scores, graph.name2id, test)})
            print(json.dumps(results[-1]), flush=True)
            del solver
    ours = [r for r in results if r["impl"] == "graphvite_b200"]
    theirs = [r for r in results if r["impl"] == "reference"]
    if ours and theirs:
        summary = {"workload": args.workload, "epochs": args.epochs}
        for key in ("vertex_norm", "context_norm", "auc"):
            a, b = np.mean([r[key] for r in ours]), np.mean([r[key] for r in theirs])
            summary[key] = {"graphvite_b200": a, "reference": b, "relative_difference": (a - b) / b,
                            "spread_ours": float(np.ptp([r[key] for r in ours])),
                            "spread_reference": float(np.ptp([r[key] for r in theirs]))}
        print(json.dumps({"summary": summary}), flush=True)


if __name__ == "__main__":
    main()
