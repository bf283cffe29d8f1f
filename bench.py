#!/usr/bin/env python
"""bench.py -- edges/sec of LINE d=128 on a Youtube-shaped synthetic graph (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU-sampler + CUDA path

A "step" is one sub-episode: every GPU trains one (head block, tail block) of the sample pool,
episode_size * batch_size = 5e7 positive edges per GPU at the reference's Youtube configuration
(config/graph/line_youtube.yaml), while the samplers refill the other pool.  For N > 1 the script
is launched by torchrun, one rank per GPU; the value is the whole job's edges per second with the
time taken as the max over ranks of the device time (CUDA events on the solver's work stream).
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# config/graph/line_youtube.yaml
YOUTUBE = dict(dim=128, lr=0.025, weight_decay=0.005, num_negative=1, batch_size=100000, episode_size=500,
               model="LINE", negative_weight=5, augmentation_step=5, random_walk_length=40,
               random_walk_batch_size=100)
WORKLOADS = {
    "youtube": dict(YOUTUBE, graph="youtube"),
    "blogcatalog": dict(YOUTUBE, graph="blogcatalog", augmentation_step=2),  # config/demo/quick_start.yaml
    "toy": dict(YOUTUBE, graph="toy", batch_size=1000, episode_size=20, augmentation_step=2, random_walk_length=10),
}


def bytes_per_edge(dim, k):
    """Algorithmic HBM bytes per positive edge for SGD: vertex row R+W,
    k+1 context rows R+W, the {tail, head} pair, k negative ids, the loss."""
    return 8 * dim * (k + 2) + 8 + 4 * k + 4


def graph_file(name):
    """Synthetic edge list with the published |V| / |E| (no network: real datasets are unavailable)."""
    from graphvite_b200 import datasets
    path = "/tmp/gv_b200_%s.txt" % name
    if not os.path.exists(path):
        num_vertex, num_edge = datasets.SHAPES[name]
        u, v = datasets.power_law_edges(num_vertex, num_edge, max_degree=29000 if name == "youtube" else None)
        import pandas
        pandas.DataFrame({"u": u, "v": v}).to_csv(path + ".tmp", sep="\t", header=False, index=False)
        os.replace(path + ".tmp", path)
    return path


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.process = [], None
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.process = None

    def _read(self):
        for line in self.process.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.process:
            return None
        self.process.terminate()
        self.thread.join(timeout=2)
        clocks = [int(r[0]) for r in self.rows if len(r) == 6 and r[0].isdigit()]
        if not clocks:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) == 6 for i in range(4) if r[2 + i] == "Active"})
        maxima = [int(r[1]) for r in self.rows if len(r) == 6 and r[1].isdigit()]
        return {"sm_mhz": int(np.median(clocks)), "sm_max_mhz": max(maxima) if maxima else None, "reasons": reasons,
                "samples": len(clocks)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def make_solver(cfg, path, rank, world, local_rank, num_partition=0):
    import graphvite_b200 as gv
    graph = gv.graph.Graph()
    graph.load(path, as_undirected=True)
    solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[local_rank], rank=rank, world_size=world)
    solver.build(graph, gv.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), num_partition=num_partition,
                 num_negative=cfg["num_negative"], batch_size=cfg["batch_size"], episode_size=cfg["episode_size"])
    return gv, graph, solver


def train_kwargs(cfg, num_epoch):
    return dict(model=cfg["model"], num_epoch=num_epoch, augmentation_step=cfg["augmentation_step"],
                random_walk_length=cfg["random_walk_length"], random_walk_batch_size=cfg["random_walk_batch_size"],
                negative_weight=cfg["negative_weight"])


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from graphvite_b200 import _lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, "--gpus must equal the number of launched ranks (torchrun --nproc-per-node)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    origin = time.time()

    def progress(what):  # stderr only: if a multi-GPU run stalls, the log says in which phase
        if rank == 0 or world > 1 and os.environ.get("GV_BENCH_VERBOSE"):
            sys.stderr.write("[bench %7.2f s] rank %d: %s\n" % (time.time() - origin, rank, what))
            sys.stderr.flush()

    if rank == 0:
        path = graph_file(cfg["graph"])
    barrier()
    path = graph_file(cfg["graph"])
    progress("graph file ready, building the solver")
    gv, graph, solver = make_solver(cfg, path, rank, world, local_rank, args.partitions)
    progress("solver built (%d partitions)" % solver.num_partition)
    lib, handle = _lib.lib, solver._handle
    edges_per_step = cfg["episode_size"] * cfg["batch_size"]  # per GPU
    kw = train_kwargs(cfg, 4000)  # config/graph/line_youtube.yaml; far more epochs than we will run
    _lib.check(lib.gv_solver_train_begin(handle, kw["model"].encode(), kw["num_epoch"], 0, kw["augmentation_step"],
                                         kw["random_walk_length"], kw["random_walk_batch_size"], 0, 1.0, 1.0, 1,
                                         0.75, float(kw["negative_weight"]), 1000))
    progress("train_begin done (tables, first pool, embeddings resident)")
    for _ in range(args.warmup):
        assert lib.gv_solver_train_step(handle) == 1, _lib.last_error()
    barrier()
    progress("%d warm-up steps done" % args.warmup)
    before = solver.stats
    sampler = ClockSampler(local_rank) if rank == 0 else None
    lib.gv_solver_device_timer(handle, 0)
    for _ in range(args.steps):
        assert lib.gv_solver_train_step(handle) == 1, _lib.last_error()
    seconds = lib.gv_solver_device_timer(handle, 1)
    barrier()
    progress("%d timed steps done" % args.steps)
    clocks = sampler.stop() if sampler else None
    after = solver.stats
    if world > 1:
        t = torch.tensor([seconds], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    _lib.check(lib.gv_solver_train_end(handle))
    norms = [float(np.linalg.norm(solver.vertex_embeddings)), float(np.linalg.norm(solver.context_embeddings))]
    value = args.steps * edges_per_step * world / seconds

    # roofline of the dominant kernel (the train kernel): algorithmic bytes / CUDA-event time of its launches
    kernel_seconds = after["kernel_seconds"] - before["kernel_seconds"]
    positives = after["positives"] - before["positives"]
    per_edge = bytes_per_edge(cfg["dim"], cfg["num_negative"])
    achieved = positives * per_edge / kernel_seconds / 1e9
    peak, peak_kind = measured_peak()
    launches = int(after["launches"] - before["launches"])
    # DRAM bytes of one launch of the same kernel from the committed `ncu --set full` capture
    traffic, traffic_note = None, "no ncu capture committed"
    capture = os.path.join(ROOT, "profiles", "r01_train_kernel.json")
    if os.path.exists(capture) and cfg["dim"] == 128 and cfg["num_negative"] == 1:
        info = json.load(open(capture))
        traffic = info["dram_bytes_per_launch"]
        traffic_note = ("dram__bytes_read+write.sum per launch of %d edges (%.0f B/edge vs %d algorithmic), %s" %
                        (info["edges_per_launch"], info["dram_bytes_per_edge"], per_edge, info["source"]))

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": seconds / args.steps * 1e3, "roofline_gbs": achieved,
                              "frac": achieved / peak, "gpu_launches": launches, "note": "profiling run, no e2e"}),
                  flush=True)
        solver.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # end to end through the public API: host graph + host embeddings in, host embeddings out
    solver.close()  # collective teardown (IPC importers close before exporters free)
    del solver
    progress("steady-state solver closed, building the end-to-end solver")
    gv2, graph2, solver2 = make_solver(cfg, path, rank, world, local_rank)
    per_episode = edges_per_step * world * (solver2.num_partition // world) ** 2 * world  # edges per episode
    episodes = max(1, int(round(args.steps * edges_per_step * world / per_episode)))
    num_epoch = max(1, int(np.ceil(episodes * per_episode / graph2.num_edge)) - 1)
    barrier()
    start = time.time()
    solver2.train(**train_kwargs(cfg, num_epoch))
    barrier()
    e2e_seconds = time.time() - start
    progress("end-to-end train() done")
    if world > 1:
        t = torch.tensor([e2e_seconds], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_seconds = float(t.item())
    e2e_edges = solver2.batch_id * cfg["batch_size"]
    e2e_steps = max(1, e2e_edges // (edges_per_step * world))
    matrix_bytes = graph2.num_vertex * cfg["dim"] * 4
    directed = 2 * graph2.num_edge
    h2d = 2 * matrix_bytes + directed * (4 + 4 + 4 + 8 + 8) + graph2.num_vertex * 24
    d2h = 2 * matrix_bytes
    e2e_stats = solver2.stats

    result = {
        "metric": "edges/sec on Youtube LINE d=128", "value": value, "unit": "edges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "graphvite_b200",
        "config": {"workload": "LINE d=%d on %s-shaped power-law graph (|V|=%d, %d edge lines), SGD lr=%g wd=%g, "
                               "k=%d, B=%d, episode_size=%d, augmentation_step=%d, walk length %d" %
                               (cfg["dim"], cfg["graph"], graph2.num_vertex, graph2.num_edge, cfg["lr"],
                                cfg["weight_decay"], cfg["num_negative"], cfg["batch_size"], cfg["episode_size"],
                                cfg["augmentation_step"], cfg["random_walk_length"]),
                   "num_partition": solver2.num_partition, "parallelism": "2-D block partition, %d GPU(s)" % world,
                   "l2": "working set (2 x %.0f MB embeddings + %.0f MB pool block) >> 126 MB L2, no flush needed" %
                         (matrix_bytes / 1e6, edges_per_step * 8 / 1e6)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_kind,
                     "bytes_per_edge": per_edge, "kernel": "gv::device::train_sgd_kernel<%d, %d>" % (cfg["dim"], cfg["num_negative"]),
                     "algorithmic_bytes_per_launch": per_edge * 16 * cfg["batch_size"],
                     "kernel_edges_per_s": positives / kernel_seconds},
        "e2e": {"value": e2e_edges / e2e_seconds, "unit": "edges/s", "h2d_bytes_per_step": h2d / e2e_steps,
                "d2h_bytes_per_step": d2h / e2e_steps, "seconds": e2e_seconds, "edges": e2e_edges,
                "sampler_seconds": e2e_stats["sample_seconds"], "train_seconds": e2e_stats["train_seconds"]},
        "gpu_launches": launches,
        "clocks": clocks,
        "embedding_norms": norms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            result["cpu_baseline"] = reference_baseline(cfg, path, 1, max(2, min(args.steps, 4)))
        except Exception as error:  # the reference build is absent: report, do not hide
            result["cpu_baseline"] = {"value": None, "unit": "edges/s", "cores": os.cpu_count(), "kind": "reference",
                                      "sample": "unavailable: %s" % error}
    if rank == 0:
        print(json.dumps(result), flush=True)
    solver2.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference (oracle/_ref/libgraphvite.so) through its own pybind API
# --------------------------------------------------------------------------------------------------
def load_reference():
    path = os.path.join(ROOT, "oracle", "_ref", "libgraphvite.so")
    if not os.path.exists(path):
        raise RuntimeError("oracle/_ref/libgraphvite.so is not built (make -C oracle ref)")
    spec = importlib.util.spec_from_file_location("libgraphvite", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    module.init_logging(module.ERROR, "", False)
    return module


def reference_train(cfg, path, num_gpu, steps):
    """Times GraphSolver.train() of the reference (its own definition of training time,
    python/graphvite/application/application.py:99-105) for about `steps` sub-episodes per GPU."""
    ref = load_reference()
    graph = ref.graph.Graph_j()
    graph.load(path, True, False)
    solver = getattr(ref.solver, "GraphSolver_%d_f_j" % cfg["dim"])(list(range(num_gpu)), 0, 0)
    solver.build(graph, ref.optimizer.SGD(cfg["lr"], cfg["weight_decay"]), 0, cfg["num_negative"],
                 cfg["batch_size"], cfg["episode_size"])
    P, E, B = solver.num_partition, solver.episode_size, solver.batch_size
    per_episode = P * P * E * B
    episodes = max(1, int(round(steps * num_gpu * E * B / per_episode)))
    num_epoch = max(1, int(np.ceil(episodes * per_episode / graph.num_edge)) - 1)
    kwargs = dict(model=cfg["model"], num_epoch=num_epoch, resume=False,
                  augmentation_step=cfg["augmentation_step"], random_walk_length=cfg["random_walk_length"],
                  random_walk_batch_size=cfg["random_walk_batch_size"], negative_weight=cfg["negative_weight"],
                  log_frequency=1 << 30)
    start = time.time()
    solver.train(**kwargs)
    seconds = time.time() - start
    num_batch = num_epoch * graph.num_edge // B
    per_episode_batches = per_episode // B
    trained = -(-num_batch // per_episode_batches) * per_episode_batches * B
    norms = [float(np.linalg.norm(solver.vertex_embeddings)), float(np.linalg.norm(solver.context_embeddings))]
    return {"edges": trained, "seconds": seconds, "num_sampler": solver.num_sampler, "num_worker": solver.num_worker,
            "num_partition": P, "norms": norms, "graph": graph, "steps": trained // (E * B * num_gpu)}


def reference_baseline(cfg, path, num_gpu, steps):
    r = reference_train(cfg, path, num_gpu, steps)
    return {"value": r["edges"] / r["seconds"], "unit": "edges/s", "cores": r["num_sampler"] + r["num_worker"],
            "kind": "reference",
            "sample": "reference GraphSolver.train(): %d edges (%d sub-episodes) in %.2f s wall, %d CPU sampler "
                      "threads + %d GPU worker thread(s), host has %d logical cores" %
                      (r["edges"], r["steps"], r["seconds"], r["num_sampler"], r["num_worker"], os.cpu_count())}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return  # the reference is one process driving all GPUs from threads
    try:
        path = graph_file(cfg["graph"])
        reference_train(cfg, path, args.gpus, 1)  # warm-up: CUDA context, page-in, first allocations
        r = reference_train(cfg, path, args.gpus, args.steps)
    except Exception as error:
        print(json.dumps({"impl": "reference", "unavailable": str(error).splitlines()[0][:200]}), flush=True)
        return
    value = r["edges"] / r["seconds"]
    graph = r["graph"]
    cores = r["num_sampler"] + r["num_worker"]
    sample = ("GraphSolver.train() wall time for %d edges (%d sub-episodes per GPU), %d CPU sampler threads + %d "
              "worker thread(s)" % (r["edges"], r["steps"], r["num_sampler"], r["num_worker"]))
    print(json.dumps({
        "metric": "edges/sec on Youtube LINE d=128", "value": value, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": int(r["steps"]), "warmup": 1, "ms_per_step": r["seconds"] / max(1, r["steps"]) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "LINE d=%d on %s-shaped power-law graph (|V|=%d, %d edge lines), SGD lr=%g wd=%g, "
                               "k=%d, B=%d, episode_size=%d, augmentation_step=%d, walk length %d" %
                               (cfg["dim"], cfg["graph"], graph.num_vertex, graph.num_edge, cfg["lr"],
                                cfg["weight_decay"], cfg["num_negative"], cfg["batch_size"], cfg["episode_size"],
                                cfg["augmentation_step"], cfg["random_walk_length"]),
                   "num_partition": r["num_partition"]},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "embedding_norms": r["norms"],
    }), flush=True)


def start_watchdog(seconds):
    """A hung collective must not hang the box: give up loudly after `seconds`."""
    def expire():
        sys.stderr.write("bench.py: no result after %d s, aborting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    timer = threading.Timer(seconds, expire)
    timer.daemon = True
    timer.start()
    return timer


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="graphvite_b200", choices=["graphvite_b200", "reference"])
    parser.add_argument("--workload", default="youtube", choices=sorted(WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    parser.add_argument("--partitions", type=int, default=0, help="num_partition (0 = auto; diagnosis only)")
    parser.add_argument("--watchdog", type=int, default=1500, help="abort after this many seconds (0 = never)")
    args = parser.parse_args()
    if args.watchdog > 0:
        start_watchdog(args.watchdog)
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    try:
        run_ours(args, cfg)
    except BaseException:
        # leave at once: a rank that unwinds normally would first wait for its sampler thread, which may be waiting
        # for the peers -- torchrun only stops the other ranks once this process is gone
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
