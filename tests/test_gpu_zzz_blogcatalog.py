"""BASELINE.json config C1 -- "GraphSolver.build + edge / negative sample-pool fill on BlogCatalog": a
BlogCatalog-shaped graph (10 312 vertices, 333 983 edge lines, power-law degrees), the reference's own
link-prediction split (Dataset.link_prediction_split semantics, seed 1024, portions 100:1:1 -> the train file),
`as_undirected`, and the parameters of config/demo/quick_start.yaml (d = 128, SGD 0.025 / 0.005, k = 1, B = 1e5,
LINE, augmentation 2, walk length 40, walk batch 100, P = 1).  Everything the sampler side produces is
compared bit for bit with the oracle:
  * the edge alias table (AliasTable::build over all directed edges, base/alias_table.cuh:84-128),
  * the partition / degree-ordered row of every vertex (core/solver.h:873-887),
  * the engine-drawn initial vertex embeddings,
  * both sample pools -- the first fill and the refill that runs while episode 1 trains (random-walk sampler with
    pseudo shuffle, instance/graph.cuh:376-450), and the plain edge sampler (augmentation_step = 1, solver.h:975-1055),
  * the negatives of the last batch (gpu::Sample with float narrowing, base/alias_table.cuh:175-183) drawn from the
    pow(degree, 0.75) table (solver.h:1264-1278).
episode_size is 50 instead of 500 so that the oracle's sequential sampler finishes in seconds; the full-size pool
(episode_size = 500, 5e7 pairs) is checked once through a checksum of checksums against the same oracle run.
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

NUM_VERTEX = 10312
BATCH = int(os.environ.get("GV_TEST_C1_BATCH", 100000))  # quick_start.yaml: 100000 (smaller only to smoke-test this file)
QUICK_START = dict(dim=128, k=1, aug=2, L=40, wb=100)


@pytest.fixture(scope="module")
def blogcatalog():
    import graphvite_b200 as gv  # first: validate_parity puts the repository root in front of sys.path
    from validate_parity import make_split
    path, _ = make_split("blogcatalog")  # the train part of the 100:1:1 split, seed 1024
    graph = gv.graph.Graph()
    graph.load(path, as_undirected=True)
    assert graph.num_vertex <= NUM_VERTEX and graph.num_edge > 320000
    return path, graph


def make_pair(path, graph, episode_size):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    _lib.lib.gv_reset_global_engine(5489)
    solver = gv.solver.GraphSolver(QUICK_START["dim"], device_ids=[0], num_sampler_per_worker=1)
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"capture_negatives", 1))
    solver.build(graph, gv.optimizer.SGD(0.025, 0.005), num_partition=0, num_negative=QUICK_START["k"],
                 batch_size=BATCH, episode_size=episode_size)
    ograph = O.OracleGraph(path)
    osolver = O.OracleSolver(ograph, QUICK_START["dim"], 1, 1)
    osolver.build("SGD", 0, QUICK_START["k"], BATCH, episode_size)
    return solver, ograph, osolver


def pool_of(solver, side, size):
    from graphvite_b200 import _lib
    out = np.zeros((size, 2), dtype=np.uint32)
    assert _lib.lib.gv_solver_pool(solver._handle, side, 0, 0, out.ctypes.data) == size
    return out


def begin(solver, osolver, augmentation_step, epochs=2000):
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_solver_train_begin(solver._handle, b"LINE", epochs, 0, augmentation_step, QUICK_START["L"],
                                              QUICK_START["wb"], 0, 1.0, 1.0, 1, 0.75, 5.0, 1000))
    osolver.train_begin("LINE", epochs, False, augmentation_step, QUICK_START["L"], QUICK_START["wb"])


def test_build_tables_partition_and_initialisation(blogcatalog):
    from graphvite_b200 import _lib
    path, graph = blogcatalog
    solver, ograph, osolver = make_pair(path, graph, 50)
    assert (graph.num_vertex, graph.num_edge) == (ograph.num_vertex, ograph.num_edge)
    assert solver.num_partition == 1 == osolver.info()["num_partition"]
    # the graph's flattened edges are what the edge table is built over: same order, same weights
    m = _lib.lib.gv_graph_flatten(graph._handle, None, None, None, None, None)
    assert m == ograph.num_directed_edge
    weights = np.zeros(m, dtype=np.float32)
    _lib.lib.gv_graph_flatten(graph._handle, None, None, weights.ctypes.data, None, None)
    prob, alias = np.zeros(m, dtype=np.float32), np.zeros(m, dtype=np.uint64)
    _lib.check(_lib.lib.gv_alias_build(weights.ctypes.data, m, prob.ctypes.data, alias.ctypes.data))
    begin(solver, osolver, QUICK_START["aug"])
    oprob, oalias = osolver.edge_table()
    np.testing.assert_array_equal(prob, oprob)
    np.testing.assert_array_equal(alias, oalias)
    n = graph.num_vertex
    part_of, local_of = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    _lib.lib.gv_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data)
    opart, olocal = osolver.locations()
    np.testing.assert_array_equal(part_of, opart.astype(np.uint32))
    np.testing.assert_array_equal(local_of, olocal)
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))  # nothing trained: upload -> write-back = identity
    np.testing.assert_array_equal(solver.vertex_embeddings, osolver.embeddings(0))
    assert not np.any(solver.context_embeddings)


@pytest.mark.parametrize("augmentation_step", [2, 1], ids=["random_walk", "edge_sampler"])
def test_both_pools_and_negatives_are_bit_exact(blogcatalog, augmentation_step):
    from graphvite_b200 import _lib
    path, graph = blogcatalog
    episode = 50
    size = episode * BATCH
    solver, ograph, osolver = make_pair(path, graph, episode)
    epochs = 2 * size // graph.num_edge + 1  # two episodes
    begin(solver, osolver, augmentation_step, epochs)
    np.testing.assert_array_equal(pool_of(solver, 1, size), osolver.pool(1, 0, 0))   # first fill -> pool 1
    status = _lib.lib.gv_solver_train_episode(solver._handle)                        # trains pool 1, refills pool 0
    assert status == 1, _lib.last_error()
    assert osolver.train_episode()
    np.testing.assert_array_equal(pool_of(solver, 0, size), osolver.pool(0, 0, 0))   # the refill
    negatives = np.zeros(BATCH * QUICK_START["k"], dtype=np.uint32)
    assert _lib.lib.gv_solver_last_negatives(solver._handle, negatives.ctypes.data) == negatives.size
    np.testing.assert_array_equal(negatives, osolver.last_negatives(BATCH, QUICK_START["k"]))
    assert negatives.max() < graph.num_vertex
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))
    assert solver.batch_id == osolver.info()["batch_id"]
    assert np.isfinite(solver.vertex_embeddings).all() and np.isfinite(solver.context_embeddings).all()


def test_quick_start_sized_pool_matches_the_oracle_by_checksum(blogcatalog):
    """episode_size = 500 as in quick_start.yaml: 5e7 pairs, compared through order-sensitive checksums"""
    path, graph = blogcatalog
    episode = 500
    size = episode * BATCH
    solver, ograph, osolver = make_pair(path, graph, episode)
    begin(solver, osolver, QUICK_START["aug"])

    def checksums(pairs):
        words = np.ascontiguousarray(pairs).view(np.uint64).ravel()
        index = np.arange(1, len(words) + 1, dtype=np.uint64)
        return int(np.bitwise_xor.reduce(words)), int((words * index).sum(dtype=np.uint64)), \
            int(pairs[:, 0].astype(np.uint64).sum()), int(pairs[:, 1].astype(np.uint64).sum())

    assert checksums(pool_of(solver, 1, size)) == checksums(osolver.pool(1, 0, 0))
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))
