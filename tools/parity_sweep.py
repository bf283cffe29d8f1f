#!/usr/bin/env python
"""Shipped-mode float parity against the UNMODIFIED reference, and what each kernel policy costs.

For one workload (BlogCatalog-shaped quick start, Youtube-shaped LINE) this trains the same split
 * with the reference (oracle/_ref/libgraphvite.so through its own pybind API) `--repeat` times, and
 * with graphvite_b200 under every kernel policy of SETTINGS `--repeat` times,
and records embedding L2 norms, held-out link-prediction AUC (Dataset.link_prediction_split semantics), the
wall time of train() and the train kernel's edges/s.  One JSON object per run, then one summary per setting
with the distance to the reference's mean and the reference's own run-to-run spread.

    python tools/parity_sweep.py --workload blogcatalog --epochs 2000 --out gpurun_out/parity_blogcatalog.jsonl

MEASUREMENT TOOLING (it executes oracle/_ref); not part of the product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from validate_parity import auc_of, make_split  # noqa: E402

# name -> tunables (gv_cuda_set_tunable) + chunk_batches (solver option) + num_partition
P = 64  # kernel_flags & 64: the persistent kernels (round 1's); without it SGD runs one warp per sample
SETTINGS = {
    # the shipped default: one warp per sample, one launch per batch, plain loads / stores (the reference's geometry)
    "per_sample": dict(kernel_flags=0, chunk_batches=1),
    "shipped": dict(kernel_flags=0, chunk_batches=1, sample_prefetch_blocks=592),
    # ... with fewer resident warps per SM (threads per block -> floor(2048 / threads) blocks of 32-register threads)
    "ps_w60": dict(kernel_flags=0, chunk_batches=1, sample_block_threads=640),
    "ps_w54": dict(kernel_flags=0, chunk_batches=1, sample_block_threads=576),
    "ps_w48": dict(kernel_flags=0, chunk_batches=1, sample_block_threads=768),
    "ps_w44": dict(kernel_flags=0, chunk_batches=1, sample_block_threads=704),
    # ... with the index lines prefetched into L2 N blocks ahead (592 blocks = one resident wave)
    "ps_pf592": dict(kernel_flags=0, chunk_batches=1, sample_prefetch_blocks=592),
    "ps_pf1184": dict(kernel_flags=0, chunk_batches=1, sample_prefetch_blocks=1184),
    "ps_pf2368": dict(kernel_flags=0, chunk_batches=1, sample_prefetch_blocks=2368),
    "ps_pf1184_serial": dict(kernel_flags=128, chunk_batches=1, sample_prefetch_blocks=1184),
    # resident blocks taking 16-sample groups by ticket, next indices loaded early, next rows prefetched into L2
    "rg_592": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=592),
    "rg_592_noprefetch": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=0, sample_resident_blocks=592),
    "rg_560": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=560),
    "rg_520": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=520),
    "rg_480": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=480),
    "rg_444": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=444),
    "rg_400": dict(kernel_flags=1024, chunk_batches=1, sample_prefetch_blocks=1, sample_resident_blocks=400),
    # ... with the reference's access timeline (128-byte segments, one dependent round trip each): a measuring variant
    "ps_timeline": dict(kernel_flags=512, chunk_batches=1),
    # ... with the vertex row complete before the first context row is requested (the reference's copy loop)
    "ps_serial_w64": dict(kernel_flags=128, chunk_batches=1),
    "ps_serial_w54": dict(kernel_flags=128, chunk_batches=1, sample_block_threads=576),
    "ps_serial_w48": dict(kernel_flags=128, chunk_batches=1, sample_block_threads=768),
    # ... without L1 (how much of the residual is L1 staleness?)
    "ps_l2only_w64": dict(kernel_flags=256, chunk_batches=1),
    "ps_l2only_w48": dict(kernel_flags=256, chunk_batches=1, sample_block_threads=768),
    "per_sample_chunk4": dict(kernel_flags=0, chunk_batches=4),
    "per_sample_chunk16": dict(kernel_flags=0, chunk_batches=16),
    # round-1 shipped policy: hub rows through L1, everything else L2-only, 16 batches per launch
    "r1_shipped": dict(hot_rows=128, kernel_flags=P, chunk_batches=16),
    "l2_only": dict(hot_rows=0, kernel_flags=P, chunk_batches=16),
    "hot128_chunk1": dict(hot_rows=128, kernel_flags=P, chunk_batches=1),
    "hot1024": dict(hot_rows=1024, kernel_flags=P, chunk_batches=16),
    "hot8192": dict(hot_rows=8192, kernel_flags=P, chunk_batches=16),
    # the reference's memory policy on the persistent kernels: every row through L1, write-back stores
    "all_l1_chunk16": dict(hot_rows=0, kernel_flags=P | 1, chunk_batches=16),
    "all_l1_chunk1": dict(hot_rows=0, kernel_flags=P | 1, chunk_batches=1),
    "all_l1_wb_chunk16": dict(hot_rows=0, kernel_flags=P | 1 | 32, chunk_batches=16),
    "all_l1_wb_chunk1": dict(hot_rows=0, kernel_flags=P | 1 | 32, chunk_batches=1),
    # + neighbouring pool entries on different warps (interleaved mapping)
    "interleaved_l2": dict(hot_rows=0, kernel_flags=P | 16, chunk_batches=16),
    "interleaved_hot128": dict(hot_rows=128, kernel_flags=P | 16, chunk_batches=16),
    "interleaved_hot1024": dict(hot_rows=1024, kernel_flags=P | 16, chunk_batches=16),
    "interleaved_all_l1_chunk16": dict(hot_rows=0, kernel_flags=P | 1 | 16, chunk_batches=16),
    "interleaved_all_l1_wb_chunk16": dict(hot_rows=0, kernel_flags=P | 1 | 16 | 32, chunk_batches=16),
    "interleaved_all_l1_wb_chunk4": dict(hot_rows=0, kernel_flags=P | 1 | 16 | 32, chunk_batches=4),
    "interleaved_all_l1_wb_chunk1": dict(hot_rows=0, kernel_flags=P | 1 | 16 | 32, chunk_batches=1),
    "interleaved_all_l1_wb_chunk1_3cta": dict(hot_rows=0, kernel_flags=P | 1 | 16 | 32, chunk_batches=1,
                                              train_blocks_per_sm=3),
}


def apply(gv, setting):
    for name in ("hot_rows", "kernel_flags", "train_blocks_per_sm"):
        gv._clib.gv_cuda_set_tunable(name.encode(), int(setting.get(name, 0)))
    gv._clib.gv_cuda_set_tunable(b"sample_block_threads", int(setting.get("sample_block_threads", 512)))
    gv._clib.gv_cuda_set_tunable(b"sample_prefetch_blocks", int(setting.get("sample_prefetch_blocks", 0)))
    gv._clib.gv_cuda_set_tunable(b"sample_resident_blocks", int(setting.get("sample_resident_blocks", 0)))


def run_ours(gv, cfg, graph, test, epochs, setting, num_partition):
    from graphvite_b200 import _lib
    gv._clib.gv_reset_global_engine(5489)
    apply(gv, setting)
    solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[0])
    solver.build(graph, gv.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), num_partition=num_partition,
                 num_negative=cfg["num_negative"], batch_size=cfg["batch_size"], episode_size=cfg["episode_size"])
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"chunk_batches", int(setting["chunk_batches"])))
    start = time.time()
    solver.train(**bench.train_kwargs(cfg, epochs))
    seconds = time.time() - start
    vertex, context = solver.vertex_embeddings, solver.context_embeddings
    scores = lambda pairs: np.einsum("ij,ij->i", vertex[pairs[:, 0]], context[pairs[:, 1]])
    stats = solver.stats
    row = {"vertex_norm": float(np.linalg.norm(vertex)), "context_norm": float(np.linalg.norm(context)),
           "auc": auc_of(scores, graph.name2id, test), "seconds": seconds, "edges": solver.batch_id * cfg["batch_size"],
           "kernel_edges_per_s": stats["positives"] / max(stats["kernel_seconds"], 1e-9),
           "train_loop_edges_per_s": stats["positives"] / max(stats["train_seconds"], 1e-9),
           "num_partition": solver.num_partition}
    solver.close()
    return row


def run_reference(ref, cfg, path, test, epochs):
    graph = ref.graph.Graph_j()
    graph.load(path, True, False)
    solver = getattr(ref.solver, "GraphSolver_%d_f_j" % cfg["dim"])([0], 0, 0)
    solver.build(graph, ref.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), 0, cfg["num_negative"],
                 cfg["batch_size"], cfg["episode_size"])
    start = time.time()
    solver.train(model=cfg["model"], num_epoch=epochs, augmentation_step=cfg["augmentation_step"],
                 random_walk_length=cfg["random_walk_length"], random_walk_batch_size=cfg["random_walk_batch_size"],
                 negative_weight=cfg["negative_weight"], log_frequency=1 << 30)
    seconds = time.time() - start
    vertex, context = np.array(solver.vertex_embeddings), np.array(solver.context_embeddings)
    scores = lambda pairs: np.einsum("ij,ij->i", vertex[pairs[:, 0]], context[pairs[:, 1]])
    return {"vertex_norm": float(np.linalg.norm(vertex)), "context_norm": float(np.linalg.norm(context)),
            "auc": auc_of(scores, graph.name2id, test), "seconds": seconds}


def summarise(name, runs, reference):
    out = {"summary": name, "runs": len(runs)}
    for key in ("vertex_norm", "context_norm", "auc"):
        mine = np.array([r[key] for r in runs])
        out[key] = float(mine.mean())
        out[key + "_spread"] = float(np.ptp(mine))
        if reference:
            theirs = np.array([r[key] for r in reference])
            out[key + "_reference"] = float(theirs.mean())
            out[key + "_reference_spread"] = float(np.ptp(theirs))
            out[key + ("_diff" if key == "auc" else "_rel")] = float(
                mine.mean() - theirs.mean() if key == "auc" else (mine.mean() - theirs.mean()) / theirs.mean())
    if runs and "kernel_edges_per_s" in runs[0]:
        out["kernel_edges_per_s"] = float(np.mean([r["kernel_edges_per_s"] for r in runs]))
        out["train_loop_edges_per_s"] = float(np.mean([r["train_loop_edges_per_s"] for r in runs]))
    if reference:
        out["within_north_star"] = bool(abs(out["vertex_norm_rel"]) <= 1e-3 and abs(out["context_norm_rel"]) <= 1e-3
                                        and abs(out["auc_diff"]) <= 3e-3)
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--workload", default="blogcatalog")
    parser.add_argument("--epochs", type=int, default=2000)
    parser.add_argument("--repeat", type=int, default=3)
    parser.add_argument("--reference-repeat", type=int, default=3)
    parser.add_argument("--settings", default="", help="comma-separated subset of SETTINGS (default: all)")
    parser.add_argument("--partitions", type=int, default=0)
    parser.add_argument("--out", default="")
    args = parser.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    path, test = make_split(cfg["graph"])
    sink = open(args.out, "a") if args.out else None

    def emit(row):
        line = json.dumps(row)
        print(line, flush=True)
        if sink:
            sink.write(line + "\n")
            sink.flush()

    emit({"workload": args.workload, "epochs": args.epochs, "partitions": args.partitions, "time": time.time()})
    reference = []
    if args.reference_repeat > 0:
        ref = bench.load_reference()
        for run in range(args.reference_repeat):
            row = run_reference(ref, cfg, path, test, args.epochs)
            reference.append(row)
            emit(dict(row, impl="reference", run=run))
        emit(summarise("reference", reference, None))

    import graphvite_b200 as gv
    graph = gv.graph.Graph()
    graph.load(path)
    names = [n for n in args.settings.split(",") if n] or list(SETTINGS)
    for name in names:
        runs = []
        for run in range(args.repeat):
            row = run_ours(gv, cfg, graph, test, args.epochs, SETTINGS[name], args.partitions)
            runs.append(row)
            emit(dict(row, impl="graphvite_b200", setting=name, run=run, **SETTINGS[name]))
        emit(summarise(name, runs, reference))


if __name__ == "__main__":
    main()
