"""Properties of the knowledge-graph path at the size of BASELINE.json's configuration #4 (RotatE d = 2048 on an
FB15k-237-shaped graph: 14 541 entities, 237 relations, 272 115 triplets; Adam, k = 64, B = 1e5), where replaying
the training with the sequential oracle is out of reach:
  * entity / relation initialisation (2.98e7 engine draws + the phases) and the first sample pool are bit-identical
    to the oracle's (a short episode keeps the oracle's sequential sampling to seconds);
  * begin -> end without a trained step is the identity on the matrices;
  * every sampled {relation, tail, head} is a triplet of the graph, the same engine seed reproduces the pool
    (checksums), negatives are in range;
  * Hogwild training over the full grid at the reference's configuration keeps every value finite, lowers the logged
    loss and ranks true tails above corrupted ones; batch accounting follows core/solver.h:611,629.
Under the CUDA emulation (GV_EMULATE=1, tests/test_emulated_kernels.py) the very same code runs at a reduced shape,
which is how its logic was checked on a machine without a GPU.
"""
import os
import sys

import numpy as np
import pytest

import oracle_kg_lib as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMULATED = os.environ.get("GV_EMULATE") == "1"
if EMULATED:  # reduced shape: same code path, sizes an emulated run finishes in seconds
    SHAPE = dict(entities=600, relations=12, triplets=6000)
    DIM, NEGATIVES, BATCH, TRAIN_EPOCHS = 32, 4, 300, 20
else:
    SHAPE = dict(entities=14541, relations=237, triplets=272115)
    DIM, NEGATIVES, BATCH, TRAIN_EPOCHS = 2048, 64, 100000, 73  # 198 batches: one auto-sized episode of 200
MARGIN, TEMPERATURE = 9.0, 2.0


@pytest.fixture(scope="module")
def fb15k(tmp_path_factory):
    import graphvite_b200 as gv
    from graphvite_b200 import datasets
    path = str(tmp_path_factory.mktemp("kg") / "fb15k237_shaped.txt")
    if EMULATED:
        heads, relations, tails = datasets.power_law_triplets(SHAPE["entities"], SHAPE["relations"], SHAPE["triplets"],
                                                              seed=3)
        with open(path, "w") as out:
            for h, r, t in zip(heads, relations, tails):
                out.write("e%d\tr%d\te%d\n" % (h, r, t))
    else:
        datasets.synthetic_knowledge_graph_file("fb15k-237", path)
    graph = gv.graph.KnowledgeGraph()
    graph.load(path)
    if not EMULATED:
        assert (graph.num_vertex, graph.num_relation, graph.num_edge) == (14541, 237, 272115)
    return path, graph


def make_solver(graph, episode_size):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    _lib.lib.gv_reset_global_engine(5489)
    solver = gv.solver.KnowledgeGraphSolver(DIM, device_ids=[0])
    solver.build(graph, gv.optimizer.Adam(5e-5 if not EMULATED else 5e-3, 0), num_negative=NEGATIVES,
                 batch_size=BATCH, episode_size=episode_size)
    assert solver.num_partition == 1
    return solver


def begin(solver, num_epoch=1000):
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_kg_solver_train_begin(solver._handle, b"RotatE", num_epoch, 0, 1.0, MARGIN, 2e-3, 2000, 1,
                                                 TEMPERATURE, 1000))


def end(solver):
    from graphvite_b200 import _lib
    _lib.check(_lib.lib.gv_kg_solver_train_end(solver._handle))


def first_pool(solver, size):
    from graphvite_b200 import _lib
    out = np.zeros((size, 3), dtype=np.uint32)
    assert _lib.lib.gv_kg_solver_pool(solver._handle, 1, 0, 0, out.ctypes.data) == size
    return out


def locations(solver, num_vertex):
    from graphvite_b200 import _lib
    part_of, local_of = np.zeros(num_vertex, dtype=np.uint32), np.zeros(num_vertex, dtype=np.uint32)
    _lib.check(_lib.lib.gv_kg_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data))
    return part_of, local_of


def test_initialisation_and_first_pool_match_the_oracle(fb15k):
    path, graph = fb15k
    episode = 4
    solver = make_solver(graph, episode)
    begin(solver)
    triplets = first_pool(solver, episode * BATCH)
    end(solver)  # nothing was trained: upload -> device blocks -> write-back must be the identity

    ograph = K.OracleKnowledgeGraph(path)
    osolver = K.OracleKGSolver(ograph, DIM, 1, 1)
    osolver.build("Adam", 0, NEGATIVES, BATCH, episode)
    osolver.train_begin(model="RotatE", num_epoch=1000, resume=False, relation_lr_multiplier=1.0, margin=MARGIN,
                        l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1,
                        adversarial_temperature=TEMPERATURE, log_frequency=1000)
    part_of, local_of = locations(solver, graph.num_vertex)
    opart, olocal = osolver.locations()
    np.testing.assert_array_equal(local_of, olocal)
    assert part_of.max() == 0 and opart.max() == 0
    np.testing.assert_array_equal(solver.entity_embeddings, osolver.entity_embeddings)
    np.testing.assert_array_equal(solver.relation_embeddings, osolver.relation_embeddings)
    np.testing.assert_array_equal(triplets, osolver.pool(1, 0, 0))


def test_pools_are_triplets_of_the_graph_and_reproducible(fb15k):
    from graphvite_b200 import _lib
    path, graph = fb15k
    episode = 10
    size = episode * BATCH
    checksums, pools = [], []
    for _ in range(2):
        solver = make_solver(graph, episode)
        begin(solver)
        pool = first_pool(solver, size)
        end(solver)
        words = pool.astype(np.uint64)
        checksums.append((int(words[:, 0].sum()), int(words[:, 1].sum()), int(words[:, 2].sum()),
                          int(np.bitwise_xor.reduce(words[:, 0] * 1000003 + words[:, 1] * 10007 + words[:, 2]))))
        pools.append(pool)
        _, local_of = locations(solver, graph.num_vertex)
        del solver
    assert checksums[0] == checksums[1]  # same engine seed -> the same pool
    pool = pools[0].astype(np.int64)
    n, num_relation = graph.num_vertex, graph.num_relation
    assert pool[:, 0].max() < num_relation and pool[:, 1:].max() < n  # {relation, tail_local, head_local}
    global_of = np.zeros(n, dtype=np.int64)
    global_of[local_of] = np.arange(n)
    m = _lib.lib.gv_kgraph_flatten(graph._handle, None, None, None, None, None, None)
    assert m == graph.num_edge
    h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
    _lib.lib.gv_kgraph_flatten(graph._handle, h.ctypes.data, t.ctypes.data, r.ctypes.data, None, None, None)
    keys = np.unique((h.astype(np.int64) * n + t.astype(np.int64)) * num_relation + r.astype(np.int64))
    sampled = (global_of[pool[:, 2]] * n + global_of[pool[:, 1]]) * num_relation + pool[:, 0]
    assert np.isin(sampled, keys).all()
    # every triplet is drawn with the same probability: the sample covers most of the graph
    if size >= 4 * m:
        assert len(np.unique(sampled)) > 0.9 * len(keys)


def test_training_at_full_size_stays_finite_and_learns(fb15k):
    from graphvite_b200 import _lib
    path, graph = fb15k
    # episode_size = auto as in the reference's configuration: max(|V| * 50 / B, 2e7 / B) = 200 batches at P = 1
    solver = make_solver(graph, 40 if EMULATED else 0)
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"capture_negatives", 1))
    episode = solver.episode_size
    assert episode == (40 if EMULATED else 200)
    solver.train("RotatE", num_epoch=TRAIN_EPOCHS, margin=MARGIN, sample_batch_size=2000,
                 adversarial_temperature=TEMPERATURE, log_frequency=episode // 10)
    num_batch = TRAIN_EPOCHS * graph.num_edge // BATCH
    assert solver.num_batch == num_batch
    assert solver.batch_id == -(-num_batch // episode) * episode  # whole episodes (core/solver.h:629)
    losses = np.asarray(solver.logged_loss)
    assert len(losses) >= 3 and np.isfinite(losses).all()
    assert losses[-1] < losses[1]  # losses[0] is the still-empty loss buffer of the first log point
    for view in (solver.entity_embeddings, solver.relation_embeddings):
        assert np.isfinite(view).all() and np.abs(view).max() > 0
    negatives = np.zeros(BATCH * NEGATIVES, dtype=np.uint32)
    count = _lib.lib.gv_kg_solver_last_negatives(solver._handle, negatives.ctypes.data)
    # a negative id below the partition size corrupts the tail, the rest corrupt the head (knowledge_graph.cuh train)
    assert count == BATCH * NEGATIVES and negatives.max() < 2 * graph.num_vertex
    assert 0.4 < (negatives < graph.num_vertex).mean() < 0.6
    stats = solver.stats
    assert stats["positives"] == solver.batch_id * BATCH and stats["kernel_seconds"] > 0

    # true tails score above random corruptions after training (RotatE logit = margin - distance)
    m = graph.num_edge
    h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
    _lib.lib.gv_kgraph_flatten(graph._handle, h.ctypes.data, t.ctypes.data, r.ctypes.data, None, None, None)
    rng = np.random.RandomState(0)
    pick = rng.randint(0, m, 2000)
    true = solver.predict(np.stack([h[pick], t[pick], r[pick]], axis=1))
    corrupted = solver.predict(np.stack([h[pick], rng.randint(0, graph.num_vertex, 2000).astype(np.uint32), r[pick]],
                                        axis=1))
    assert np.isfinite(true).all() and np.isfinite(corrupted).all()
    assert (true > corrupted).mean() > 0.55  # chance is 0.5 +- 0.011 on 2000 pairs
