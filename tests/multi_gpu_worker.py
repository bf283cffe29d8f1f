"""Worker of tests/test_gpu_w_multi.py: one rank of a world_size-N GraphSolver on the toy graph in
single-warp (sequential) mode, checked against the oracle emulating N workers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import graphvite_b200 as gv  # noqa: E402
from graphvite_b200 import _lib  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    toy = os.path.join(ROOT, "tests", "golden", "toy_graph.txt")
    num_partition = int(os.environ.get("GV_TEST_PARTITIONS", world))
    cfg = dict(dim=32, k=2, B=300, E=2, S=2, epochs=6, aug=2, L=6, wb=10)

    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.Graph()
    graph.load(toy)
    solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[local], num_sampler_per_worker=cfg["S"], rank=rank,
                                   world_size=world)
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"train_num_warps", 1))
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_partition, cfg["k"], cfg["B"], cfg["E"])
    assert solver.num_partition == num_partition and solver.num_worker == world
    print("rank %d sampling: %s" % (rank, "replicated" if os.environ.get("GV_REPLICATED_SAMPLING") else "partitioned"))

    ograph = O.OracleGraph(toy)
    osolver = O.OracleSolver(ograph, cfg["dim"], world, cfg["S"])
    osolver.build("SGD", num_partition, cfg["k"], cfg["B"], cfg["E"])

    args = (b"LINE", cfg["epochs"], 0, cfg["aug"], cfg["L"], cfg["wb"], 0, 1.0, 1.0, 1, 0.75, 5.0, 1000)
    _lib.check(_lib.lib.gv_solver_train_begin(solver._handle, *args))
    osolver.train_begin("LINE", cfg["epochs"], False, cfg["aug"], cfg["L"], cfg["wb"])
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
                    assert count == 0  # blocks of other ranks' tail partitions are not stored here

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
    # every rank ends with the complete matrices
    np.testing.assert_allclose(solver.vertex_embeddings, osolver.embeddings(0), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.context_embeddings, osolver.embeddings(1), rtol=1e-3, atol=1e-5)
    solver.close()
    dist.barrier()
    print("rank %d ok: %d episodes, %d partitions" % (rank, episodes, num_partition), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
