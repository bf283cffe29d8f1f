"""More end-to-end solver cases against the oracle: positive_reuse > 1, a second train() call on
the same build (resident sampler tables are reused), custom learning-rate schedules."""
import numpy as np
import pytest

from test_gpu_solver import make_oracle, make_product, product_pool

pytestmark = pytest.mark.gpu


def test_positive_reuse_and_second_train_call(toy_graph_file):
    """positive_reuse > 1 (every pool block is trained twice with fresh negatives) and the reference's
    quirk that `shuffle_base` follows the model of the PREVIOUS train() call (graph.cuh:785 runs before
    solver.h:592): a second DeepWalk call is forced to shuffle_base 1, the first is not."""
    cfg = dict(dim=32, P=2, k=1, B=300, E=4, S=1, model="DeepWalk", epochs=2, aug=2, L=6, wb=10, optimizer="SGD")
    gv, _lib, graph, solver = make_product(cfg, toy_graph_file, single_warp=True)
    ograph, osolver = make_oracle(cfg, toy_graph_file)
    for call, resume in enumerate([False, True]):
        solver.train("DeepWalk", cfg["epochs"], resume, cfg["aug"], cfg["L"], cfg["wb"], positive_reuse=2)
        osolver.train(model="DeepWalk", num_epoch=cfg["epochs"], resume=resume, augmentation_step=cfg["aug"],
                      random_walk_length=cfg["L"], random_walk_batch_size=cfg["wb"], positive_reuse=2)
        info = osolver.info()
        assert solver.shuffle_base == info["shuffle_base"] == (2 if call == 0 else 1)
        assert solver.batch_id == info["batch_id"]
        for h in range(2):
            for t in range(2):
                np.testing.assert_array_equal(product_pool(_lib, solver, 0, h, t, 1200), osolver.pool(0, h, t))
        np.testing.assert_allclose(solver.vertex_embeddings, osolver.embeddings(0), rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(solver.context_embeddings, osolver.embeddings(1), rtol=1e-3, atol=1e-5)


def test_custom_schedule_callback(toy_graph_file):
    """LRSchedule(callable) (core/optimizer.h:56-58): a Python schedule function reaches the kernel"""
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    results = []
    for optimizer in (gv.optimizer.SGD(0.05, 0.001, schedule=lambda batch_id, num_batch: 0.5),
                      gv.optimizer.SGD(0.025, 0.001, schedule="constant")):
        _lib.lib.gv_reset_global_engine(5489)
        graph = gv.graph.Graph()
        graph.load(toy_graph_file)
        solver = gv.solver.GraphSolver(32, device_ids=[0])
        _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"train_num_warps", 1))
        solver.build(graph, optimizer, num_negative=1, batch_size=500, episode_size=4)
        solver.train("LINE", 4, augmentation_step=2, random_walk_length=5, random_walk_batch_size=10)
        results.append(np.array(solver.vertex_embeddings))
    np.testing.assert_allclose(results[0], results[1], rtol=1e-5, atol=1e-7)
    assert np.abs(results[0]).max() > 0


@pytest.mark.parametrize("case", ["deepwalk_p1", "node2vec_p2"])
def test_scheduling_experiments_do_not_change_results(case, toy_graph_file):
    """The co-scheduling knobs of DESIGN.md section 8 -- walk kernels on a capped grid (grid-stride over the walks),
    train warps that take their chunks by ticket -- only change who computes what: pools, negatives and (single
    warp) embeddings must still equal the oracle's."""
    import test_gpu_solver as base
    from graphvite_b200 import _lib
    assert _lib.lib.gv_cuda_set_tunable(b"sampler_max_ctas", 1) == 0   # 256 threads walk all walks of a launch
    assert _lib.lib.gv_cuda_set_tunable(b"kernel_flags", 8) == 0
    try:
        base.test_solver_matches_oracle_step_by_step(case, toy_graph_file)
    finally:
        assert _lib.lib.gv_cuda_set_tunable(b"sampler_max_ctas", 0) == 0
        assert _lib.lib.gv_cuda_set_tunable(b"kernel_flags", 0) == 0
