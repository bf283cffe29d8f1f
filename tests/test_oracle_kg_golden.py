"""The knowledge-graph oracle (and the product's loader) against golden vectors recorded from the
UNMODIFIED reference by oracle/make_golden_kg.py.

kg_graph_n*.npz were produced in the authoring container (the reference's graph loader is host code); the
kernel / predict / solver fixtures were recorded there too, from the reference's own kernels and solver executed
by the CUDA emulation of tests/emu (`make -C oracle ref_emu && python oracle/make_golden_kg.py --emulated DIR`,
no GPU needed; see the header of oracle/gv_oracle_kg.cpp for what that does and does not cover)."""
import glob
import os

import numpy as np
import pytest

import oracle_kg_lib as K

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOY = os.path.join(GOLDEN, "toy_kg.txt")


def golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("%s has not been generated yet (oracle/make_golden_kg.py needs a GPU box)" % name)
    return np.load(path)


@pytest.mark.parametrize("normalization", [0, 1])
def test_graph_loader(normalization):
    g = golden("kg_graph_n%d.npz" % normalization)
    oracle = K.OracleKnowledgeGraph(TOY, bool(normalization))
    assert (oracle.num_vertex, oracle.num_edge, oracle.num_relation) == tuple(int(x) for x in g["sizes"])
    for ours, theirs in zip(oracle.flat(), (g["h"], g["t"], g["r"], g["w"], g["vertex_weights"])):
        np.testing.assert_array_equal(ours, theirs)  # weights bit-exact, normalised or not


@pytest.mark.parametrize("normalization", [0, 1])
def test_product_graph_loader(normalization):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    g = golden("kg_graph_n%d.npz" % normalization)
    graph = gv.graph.KnowledgeGraph()
    graph.load(TOY, bool(normalization))
    assert (graph.num_vertex, graph.num_edge, graph.num_relation) == tuple(int(x) for x in g["sizes"])
    m = graph.num_edge
    h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
    w, vw = np.zeros(m, dtype=np.float32), np.zeros(graph.num_vertex, dtype=np.float32)
    _lib.lib.gv_kgraph_flatten(graph._handle, h.ctypes.data, t.ctypes.data, r.ctypes.data, w.ctypes.data, None,
                               vw.ctypes.data)
    for ours, theirs in zip((h, t, r, w, vw), (g["h"], g["t"], g["r"], g["w"], g["vertex_weights"])):
        np.testing.assert_array_equal(ours, theirs)


KERNEL_FILES = sorted(glob.glob(os.path.join(GOLDEN, "kg_kernel_*.npz")))


@pytest.mark.skipif(not KERNEL_FILES, reason="kg_kernel_*.npz not generated yet (needs a GPU box)")
@pytest.mark.parametrize("path", KERNEL_FILES, ids=[os.path.basename(p)[10:-4] for p in KERNEL_FILES])
def test_train_kernels_on_race_free_batches(path):
    """rows are disjoint between samples, so the reference's Hogwild launch equals the sequential oracle"""
    g = np.load(path)
    model, dim = os.path.basename(path).split("_")[2], int(os.path.basename(path).split("_")[3][1:])
    otype, lr, wd, a, b, eps, rlm, margin_or_l3, temperature = (float(x) for x in g["hyper"])
    head, tail, relation = g["before_head"].copy(), g["before_tail"].copy(), g["before_relation"].copy()
    names = ["hm1", "tm1", "rm1", "hm2", "tm2", "rm2"]
    moments = {name: g["before_" + name].copy() for name in names}
    num_moment = 0 if otype == 0 else (2 if otype == 4 else 1)
    L = K.lib()
    pointer = lambda name, order: K.ptr(moments[name]) if num_moment >= order else None
    batch, negatives = np.ascontiguousarray(g["batch"]), np.ascontiguousarray(g["negatives"])
    loss = np.zeros(len(batch), dtype=np.float32)
    K.check(L.og_kg_train_batch(model.encode(), dim, head.shape[0], K.ptr(head), K.ptr(tail), K.ptr(relation),
                                pointer("hm1", 1), pointer("tm1", 1), pointer("rm1", 1), pointer("hm2", 2),
                                pointer("tm2", 2), pointer("rm2", 2), K.ptr(batch), K.ptr(negatives), len(batch),
                                negatives.shape[1], int(otype), lr, wd, a, b, eps, rlm, margin_or_l3, temperature,
                                K.ptr(loss)))
    tolerance = dict(rtol=5e-4, atol=5e-6)
    np.testing.assert_allclose(loss, g["loss"], **tolerance)
    np.testing.assert_allclose(head, g["after_head"], **tolerance)
    np.testing.assert_allclose(tail, g["after_tail"], **tolerance)
    np.testing.assert_allclose(relation, g["after_relation"], **tolerance)
    for name in names[:3 * num_moment]:
        np.testing.assert_allclose(moments[name], g["after_" + name], **tolerance)


PREDICT_FILES = sorted(glob.glob(os.path.join(GOLDEN, "kg_predict_*.npz")))


@pytest.mark.skipif(not PREDICT_FILES, reason="kg_predict_*.npz not generated yet (needs a GPU box)")
@pytest.mark.parametrize("path", PREDICT_FILES, ids=[os.path.basename(p)[11:-4] for p in PREDICT_FILES])
def test_predict_kernel(path):
    g = np.load(path)
    model = os.path.basename(path).split("_")[2]
    expected = g["logits"]
    got = np.array([K.forward(model, g["entity"][h], g["entity"][t], g["relation"][r], float(g["margin"]))
                    for r, t, h in g["batch"]], dtype=np.float32)
    np.testing.assert_allclose(got, expected, rtol=5e-5, atol=5e-5)


SOLVER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "kg_solver_*.npz")))


@pytest.mark.skipif(not SOLVER_FILES, reason="kg_solver_*.npz not generated yet (needs a GPU box)")
@pytest.mark.parametrize("path", SOLVER_FILES, ids=[os.path.basename(p)[10:-4] for p in SOLVER_FILES])
def test_solver_runs(path):
    """integer state bit-exact (partition, both sample pools, last negatives, schedule, batch accounting);
    embeddings only statistically (the reference trains Hogwild, the oracle sequentially)"""
    import oracle_lib as O
    g = np.load(path)
    cfg = {key[4:]: g[key].item() for key in g.files if key.startswith("cfg_")}
    graph = K.OracleKnowledgeGraph(TOY)
    solver = K.OracleKGSolver(graph, cfg["dim"], 1, cfg["S"])
    solver.build(O.OPTIMIZERS[cfg["optimizer"]], cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    solver.train(model=cfg["model"], num_epoch=cfg["epochs"], relation_lr_multiplier=cfg["rlm"], margin=cfg["margin"],
                 l3_regularization=cfg["l3"], sample_batch_size=cfg["sbs"], positive_reuse=cfg["reuse"],
                 adversarial_temperature=cfg["temperature"], log_frequency=100)
    info, ref = solver.info(), g["info"]
    for key, index in (("num_partition", 0), ("episode_size", 1), ("batch_size", 2), ("num_batch", 3),
                       ("batch_id", 4), ("pool_id", 5), ("num_sampler", 6), ("assignment_offset", 7),
                       ("last_negative_count", 8), ("shuffle_partition", 9)):
        assert info[key] == ref[index], key
    part_of, local_of = solver.locations()
    np.testing.assert_array_equal(part_of, g["part_of"])
    np.testing.assert_array_equal(local_of, g["local_of"])
    P = info["num_partition"]
    for side in range(2):
        for h in range(P):
            for t in range(P):
                np.testing.assert_array_equal(solver.pool(side, h, t), g["pools"][side, h, t],
                                              err_msg="pool %d block (%d, %d)" % (side, h, t))
    np.testing.assert_array_equal(solver.last_negatives(), g["negatives"])
    np.testing.assert_array_equal(solver.schedule(1), g["schedule"])
    assert (g["negative_prob"] == 1).all() and (g["negative_alias"] == np.arange(len(g["negative_alias"]))).all()
    emulated = "emulated" in g.files and int(g["emulated"]) == 1
    tolerance = dict(rtol=1e-4, atol=1e-6)
    if emulated and P == 1:
        # Recorded from the reference under the CUDA emulation, where the samples of a batch are processed in
        # order (one warp per sample, warps and CTAs one after another): exactly the oracle's order, so the floats
        # must agree too -- the whole training run, not only its integer state.
        np.testing.assert_allclose(solver.entity_embeddings, g["entity_0"], **tolerance)
        np.testing.assert_allclose(solver.relation_embeddings, g["relation_0"], **tolerance)
        np.testing.assert_allclose(solver.last_loss(), g["loss"], **tolerance)
        np.testing.assert_allclose(solver.predict(g["triplets"]), g["logits"], rtol=1e-4, atol=1e-5)
        return
    # Hogwild on a GPU, or several partitions: with P > 1 the reference's partition cache can keep a second, stale
    # copy of an entity partition (a "tail hit" leaves the trained tail copy on the device while the head copy of
    # the same partition is reloaded from host memory, core/solver.h:1436-1476) and the later write-back wins,
    # which the oracle's single in-place matrix does not imitate by default: only magnitudes are compared ...
    for ours, name in ((solver.entity_embeddings, "entity_0"), (solver.relation_embeddings, "relation_0")):
        assert np.linalg.norm(ours) == pytest.approx(np.linalg.norm(g[name]), rel=0.05), name
    assert float(solver.last_loss().mean()) == pytest.approx(float(g["loss"].mean()), rel=0.2, abs=0.02)
    if emulated:
        # ... and with the oracle's restatement of that cache switched on, the P > 1 runs agree float for float too
        cached = K.OracleKGSolver(graph, cfg["dim"], 1, cfg["S"])
        cached.set_reference_cache(True)
        cached.build(O.OPTIMIZERS[cfg["optimizer"]], cfg["P"], cfg["k"], cfg["B"], cfg["E"])
        cached.train(model=cfg["model"], num_epoch=cfg["epochs"], relation_lr_multiplier=cfg["rlm"],
                     margin=cfg["margin"], l3_regularization=cfg["l3"], sample_batch_size=cfg["sbs"],
                     positive_reuse=cfg["reuse"], adversarial_temperature=cfg["temperature"], log_frequency=100)
        np.testing.assert_allclose(cached.entity_embeddings, g["entity_0"], **tolerance)
        np.testing.assert_allclose(cached.relation_embeddings, g["relation_0"], **tolerance)
        np.testing.assert_allclose(cached.last_loss(), g["loss"], **tolerance)
        np.testing.assert_allclose(cached.predict(g["triplets"]), g["logits"], rtol=1e-4, atol=1e-5)
        for order in (1, 2):
            if "entity_%d" % order in g.files:
                np.testing.assert_allclose(cached.matrix(0, order), g["entity_%d" % order], **tolerance)
