"""One rank of a world_size-N solver on the toy inputs, checked against the oracle emulating N workers.
Launched by tests/test_gpu_w_multi.py (real GPUs, NCCL) and by tests/test_emulated_multi_rank.py (the CUDA
emulation of tests/emu with GV_EMULATE=1: "device" memory is host memory, so the very same block exchange and
relation all-reduce run over gloo on a machine without GPUs).

    GV_TEST_SOLVER      graph | kg
    GV_TEST_MODEL       graph only: LINE | DeepWalk | node2vec
    GV_TEST_PARTITIONS  number of partitions
    GV_TEST_OPTIMIZER   kg only: SGD | Adam ...
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMULATED = os.environ.get("GV_EMULATE") == "1"
if EMULATED:
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu", "_pkg"))
sys.path.insert(1, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

import oracle_kg_lib as K  # noqa: E402
import oracle_lib as O  # noqa: E402
import graphvite_b200 as gv  # noqa: E402
from graphvite_b200 import _lib, distributed  # noqa: E402


def run_graph(rank, world, local, num_partition):
    toy = os.path.join(ROOT, "tests", "golden", "toy_graph.txt")
    cfg = dict(dim=32, k=2, B=300, E=2, S=2, epochs=6, aug=2, L=6, wb=10)
    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.Graph()
    graph.load(toy)
    solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[local], num_sampler_per_worker=cfg["S"], rank=rank,
                                   world_size=world)
    if EMULATED:
        distributed.attach(solver, None)  # host buffers over gloo
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"train_num_warps", 1))
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_partition, cfg["k"], cfg["B"], cfg["E"])
    assert solver.num_partition == num_partition and solver.num_worker == world
    ograph = O.OracleGraph(toy)
    osolver = O.OracleSolver(ograph, cfg["dim"], world, cfg["S"])
    osolver.build("SGD", num_partition, cfg["k"], cfg["B"], cfg["E"])
    model = os.environ.get("GV_TEST_MODEL", "LINE")  # node2vec: per-edge tables sharded over the ranks (needs IPC)
    p, q = (0.5, 2.0) if model == "node2vec" else (1.0, 1.0)
    args = (model.encode(), cfg["epochs"], 0, cfg["aug"], cfg["L"], cfg["wb"], 0, p, q, 1, 0.75, 5.0, 1000)
    _lib.check(_lib.lib.gv_solver_train_begin(solver._handle, *args))
    osolver.train_begin(model, cfg["epochs"], False, cfg["aug"], cfg["L"], cfg["wb"], 0, p, q)
    size = cfg["B"] * cfg["E"]

    def check_pools(side):
        for h in range(num_partition):
            for t in range(num_partition):
                out = np.zeros((size, 2), dtype=np.uint32)
                count = _lib.lib.gv_solver_pool(solver._handle, side, h, t, out.ctypes.data)
                if t % world == rank:
                    assert count == size
                    np.testing.assert_array_equal(out, osolver.pool(side, h, t), err_msg="pool %d (%d,%d)" % (side, h, t))
                else:
                    assert count == 0

    check_pools(1)
    episodes = 0
    while True:
        status = _lib.lib.gv_solver_train_episode(solver._handle)
        assert status >= 0, _lib.last_error()
        more = osolver.train_episode()
        assert (status == 1) == more
        if not more:
            break
        episodes += 1
        check_pools(osolver.info()["pool_id"] ^ 1)
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))
    assert episodes >= 1 and solver.batch_id == osolver.info()["batch_id"]
    np.testing.assert_allclose(solver.vertex_embeddings, osolver.embeddings(0), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.context_embeddings, osolver.embeddings(1), rtol=1e-3, atol=1e-5)
    solver.close()
    return episodes


def run_kg(rank, world, local, num_partition):
    toy = os.path.join(ROOT, "tests", "golden", "toy_kg.txt")
    optimizer = os.environ.get("GV_TEST_OPTIMIZER", "SGD")
    cfg = dict(dim=32, k=3, B=60, E=2, S=2, epochs=3, sb=37)
    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.KnowledgeGraph()
    graph.load(toy)
    solver = gv.solver.KnowledgeGraphSolver(cfg["dim"], device_ids=[local], num_sampler_per_worker=cfg["S"], rank=rank,
                                            world_size=world)
    if EMULATED:
        distributed.attach_knowledge_graph(solver, None)
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"train_num_groups", 1))
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"capture_negatives", 1))
    otype, lr, wd, a, b, eps = O.OPTIMIZERS[optimizer]
    kwargs = {"SGD": {}, "Momentum": dict(momentum=a), "AdaGrad": dict(epsilon=eps),
              "RMSprop": dict(alpha=a, epsilon=eps), "Adam": dict(beta1=a, beta2=b, epsilon=eps)}[optimizer]
    solver.build(graph, getattr(gv.optimizer, optimizer)(lr, wd, **kwargs), num_partition, cfg["k"], cfg["B"], cfg["E"])
    assert solver.num_partition == num_partition and solver.num_worker == world and solver.shuffle_partition == 0

    ograph = K.OracleKnowledgeGraph(toy)
    osolver = K.OracleKGSolver(ograph, cfg["dim"], world, cfg["S"])
    # the product's documented multi-worker semantics: no tail rotation, deltas of a step summed before the next load
    osolver.set_emulation(shuffle_override=0, synchronous_relation=True)
    osolver.build(optimizer, num_partition, cfg["k"], cfg["B"], cfg["E"])
    np.testing.assert_array_equal(distributed.kg_schedule(num_partition, world), osolver.schedule(world))

    kw = dict(model="RotatE", num_epoch=cfg["epochs"], resume=False, relation_lr_multiplier=1.0, margin=6.0,
              l3_regularization=2e-3, sample_batch_size=cfg["sb"], positive_reuse=1, adversarial_temperature=2.0,
              log_frequency=1000)
    _lib.check(_lib.lib.gv_kg_solver_train_begin(
        solver._handle, kw["model"].encode(), kw["num_epoch"], 0, kw["relation_lr_multiplier"], kw["margin"],
        kw["l3_regularization"], kw["sample_batch_size"], kw["positive_reuse"], kw["adversarial_temperature"],
        kw["log_frequency"]))
    osolver.train_begin(**kw)
    size = cfg["B"] * cfg["E"]

    def check_pools(side):  # every rank samples all blocks itself
        for h in range(num_partition):
            for t in range(num_partition):
                out = np.zeros((size, 3), dtype=np.uint32)
                assert _lib.lib.gv_kg_solver_pool(solver._handle, side, h, t, out.ctypes.data) == size
                np.testing.assert_array_equal(out, osolver.pool(side, h, t), err_msg="pool %d (%d,%d)" % (side, h, t))

    check_pools(1)
    episodes = 0
    while True:
        status = _lib.lib.gv_kg_solver_train_episode(solver._handle)
        assert status >= 0, _lib.last_error()
        more = osolver.train_episode()
        assert (status == 1) == more
        if not more:
            break
        episodes += 1
        check_pools(osolver.info()["pool_id"] ^ 1)
    _lib.check(_lib.lib.gv_kg_solver_train_end(solver._handle))
    assert episodes >= 1 and solver.batch_id == osolver.info()["batch_id"]
    # every rank ends with the complete matrices
    np.testing.assert_allclose(solver.entity_embeddings, osolver.entity_embeddings, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.relation_embeddings, osolver.relation_embeddings, rtol=1e-3, atol=1e-5)
    solver.close()
    return episodes


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if EMULATED else int(os.environ.get("LOCAL_RANK", rank))
    if EMULATED:
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = os.environ.get("GV_TEST_SOLVER", "graph")
    num_partition = int(os.environ.get("GV_TEST_PARTITIONS", world if which == "graph" else 2 * world))
    episodes = (run_graph if which == "graph" else run_kg)(rank, world, local, num_partition)
    dist.barrier()
    print("rank %d ok: %s, %d episodes, %d partitions" % (rank, which, episodes, num_partition), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
