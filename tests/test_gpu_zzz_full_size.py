"""Properties at the benchmark's full graph size (Youtube-shaped: |V| = 1 138 499, 4 945 382 edge lines,
d = 128, B = 1e5), where replaying the training with the sequential oracle is out of reach but
cheaper invariants still pin the path:
  * partition, engine initialisation and the first sample pool are bit-identical to the oracle's
    (a short episode keeps the oracle's sequential sampling to a few seconds);
  * a load -> write-back round trip through the device blocks is the identity;
  * sampled pairs lie inside their block, edge-mode pairs are edges of the graph, negatives are in
    range and favour hubs; the same engine seed reproduces the pools (checksum of checksums);
  * 2e8 Hogwild updates at the reference's Youtube configuration keep every value finite and lower
    the logged loss; batch accounting follows core/solver.h:611,629.
"""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_VERTEX, NUM_EDGE, BATCH = 1138499, 4945382, 100000


@pytest.fixture(scope="module")
def youtube():
    import bench
    import graphvite_b200 as gv
    path = bench.graph_file("youtube")
    graph = gv.graph.Graph()
    graph.load(path)
    assert (graph.num_vertex, graph.num_edge) == (NUM_VERTEX, NUM_EDGE)
    return path, graph


def make_solver(graph, episode_size):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    _lib.lib.gv_reset_global_engine(5489)
    solver = gv.solver.GraphSolver(128, device_ids=[0])
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_negative=1, batch_size=BATCH, episode_size=episode_size)
    return solver


def begin(solver, augmentation_step):
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_solver_train_begin(solver._handle, b"LINE", 4000, 0, augmentation_step, 40, 100, 0, 1.0,
                                              1.0, 1, 0.75, 5.0, 1000))


def end(solver):
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))


def first_pool(solver, size):
    from graphvite_b200 import _lib
    out = np.zeros((size, 2), dtype=np.uint32)
    assert _lib.lib.gv_solver_pool(solver._handle, 1, 0, 0, out.ctypes.data) == size
    return out


def locations(solver):
    from graphvite_b200 import _lib
    part_of, local_of = np.zeros(NUM_VERTEX, dtype=np.uint32), np.zeros(NUM_VERTEX, dtype=np.uint32)
    _lib.lib.gv_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data)
    return part_of, local_of


def test_partition_initialisation_and_first_pool_match_the_oracle(youtube):
    path, graph = youtube
    episode = 10  # 1e6 pairs: the oracle samples them sequentially in about a second
    solver = make_solver(graph, episode)
    begin(solver, augmentation_step=5)
    pairs = first_pool(solver, episode * BATCH)
    end(solver)  # no step was trained: upload -> device blocks -> write-back must be the identity

    ograph = O.OracleGraph(path)
    assert (ograph.num_vertex, ograph.num_edge, ograph.num_directed_edge) == (NUM_VERTEX, NUM_EDGE, 2 * NUM_EDGE)
    osolver = O.OracleSolver(ograph, 128, 1, 1)
    osolver.build("SGD", 0, 1, BATCH, episode)
    osolver.train_begin("LINE", 4000, False, 5, 40, 100)

    part_of, local_of = locations(solver)
    opart, olocal = osolver.locations()
    np.testing.assert_array_equal(local_of, olocal)   # degree order, including libstdc++'s order of ties
    assert part_of.max() == 0 and opart.max() == 0
    weights = ograph.vertex_weights()
    by_row = weights[np.argsort(local_of)]
    assert (by_row[:-1] >= by_row[1:]).all()           # rows are in non-increasing degree order

    np.testing.assert_array_equal(solver.vertex_embeddings, osolver.embeddings(0))  # 1.46e8 engine draws
    assert not np.any(solver.context_embeddings)
    np.testing.assert_array_equal(pairs, osolver.pool(1, 0, 0))                     # 1e6 sampled pairs


def test_pools_are_valid_reproducible_and_edge_mode_pairs_are_edges(youtube):
    from graphvite_b200 import _lib
    path, graph = youtube
    episode = 20
    size = episode * BATCH
    checksums = []
    for _ in range(2):
        solver = make_solver(graph, episode)
        begin(solver, augmentation_step=5)
        pairs = first_pool(solver, size)
        end(solver)
        assert pairs.max() < NUM_VERTEX               # {tail_local, head_local} inside the only block
        words = np.ascontiguousarray(pairs).view(np.uint64).ravel()
        checksums.append((int(pairs[:, 0].astype(np.uint64).sum()), int(pairs[:, 1].astype(np.uint64).sum()),
                          int(np.bitwise_xor.reduce(words))))
        del solver
    assert checksums[0] == checksums[1]                # same engine seed -> the same pool

    # augmentation_step = 1 selects the plain edge sampler (instance/graph.cuh:681-682): pairs are edges
    solver = make_solver(graph, episode)
    begin(solver, augmentation_step=1)
    pairs = first_pool(solver, size)[:500000].astype(np.int64)
    end(solver)
    _, local_of = locations(solver)
    global_of = np.zeros(NUM_VERTEX, dtype=np.int64)
    global_of[local_of] = np.arange(NUM_VERTEX)
    m = _lib.lib.gv_graph_flatten(graph._handle, None, None, None, None, None)
    assert m == 2 * NUM_EDGE
    u, v = np.zeros(m, dtype=np.uint32), np.zeros(m, dtype=np.uint32)
    _lib.lib.gv_graph_flatten(graph._handle, u.ctypes.data, v.ctypes.data, None, None, None)
    edge_keys = np.unique(u.astype(np.int64) * NUM_VERTEX + v.astype(np.int64))
    sampled = global_of[pairs[:, 1]] * NUM_VERTEX + global_of[pairs[:, 0]]  # head * |V| + tail
    assert np.isin(sampled, edge_keys).all()


def test_training_at_full_size_stays_finite_and_learns(youtube):
    from graphvite_b200 import _lib
    path, graph = youtube
    solver = make_solver(graph, 500)
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"capture_negatives", 1))
    solver.train("LINE", num_epoch=40, augmentation_step=5, random_walk_length=40, random_walk_batch_size=100,
                 log_frequency=100)
    assert solver.num_batch == 40 * NUM_EDGE // BATCH == 1978
    assert solver.batch_id == 2000                     # whole episodes of 500 batches (core/solver.h:629)
    losses = solver.logged_loss
    assert len(losses) == 20 and losses[0] == 0        # the first log point shows the still-empty loss buffer
    assert np.isfinite(losses).all() and losses[-1] < losses[1] < 0.7
    for view in (solver.vertex_embeddings, solver.context_embeddings):
        assert np.isfinite(view).all() and 0 < np.abs(view).max() < 50
    negatives = np.zeros(BATCH, dtype=np.uint32)
    assert _lib.lib.gv_solver_last_negatives(solver._handle, negatives.ctypes.data) == BATCH
    assert negatives.max() < NUM_VERTEX
    counts = np.bincount(negatives, minlength=NUM_VERTEX)
    assert counts[:1000].sum() > counts[-1000:].sum()  # degree^0.75 over degree-ordered rows: hubs dominate
    stats = solver.stats
    assert stats["positives"] == 2000 * BATCH and stats["kernel_seconds"] > 0
