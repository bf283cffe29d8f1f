"""GPU parity of the knowledge-graph host runtime (gv_kg_solver_*) against the oracle's KGSolver
(oracle/gv_oracle_kg.cpp): partition, sample pools and negative indices bit-exact for the default engine
seed; entity / relation embeddings, logged loss and predict within float tolerance (rtol 1e-3) in the
one-group (sequential) mode.  The same file runs on the CPU under tests/emu (test_emulated_kernels.py).

The oracle's solver half is parity-unpinned until oracle/make_golden_kg.py has recorded the reference on a
GPU (see its header); what this file pins is product == oracle."""
import os

import numpy as np
import pytest

import oracle_kg_lib as K
import oracle_lib as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOY_KG = os.path.join(ROOT, "tests", "golden", "toy_kg.txt")

# dim, P, k, B (batch), E (episode), S (samplers), model, epochs, optimizer, sample batch, extra train kwargs
CASES = {
    "rotate_adam_p1": dict(dim=32, P=1, k=4, B=100, E=3, S=1, model="RotatE", epochs=1, optimizer="Adam", sb=64,
                           train=dict(margin=6.0, adversarial_temperature=2.0)),
    "transe_sgd_p1_s2": dict(dim=32, P=1, k=3, B=120, E=2, S=2, model="TransE", epochs=1, optimizer="SGD", sb=50,
                             train=dict(margin=8.0, adversarial_temperature=0.0)),
    "distmult_adagrad_p2": dict(dim=64, P=2, k=2, B=80, E=2, S=1, model="DistMult", epochs=2, optimizer="AdaGrad",
                                sb=40, train=dict(l3_regularization=1e-3, adversarial_temperature=1.0)),
    "complex_sgd_p4": dict(dim=32, P=4, k=2, B=60, E=2, S=3, model="ComplEx", epochs=2, optimizer="SGD", sb=32,
                           train=dict(l3_regularization=2e-3, relation_lr_multiplier=0.5)),
    "simple_momentum_p2": dict(dim=32, P=2, k=3, B=90, E=1, S=2, model="SimplE", epochs=2, optimizer="Momentum",
                               sb=77, train=dict(adversarial_temperature=0.5, positive_reuse=2)),
    "quate_adam_p2": dict(dim=64, P=2, k=3, B=75, E=2, S=1, model="QuatE", epochs=2, optimizer="Adam", sb=33,
                          train=dict(l3_regularization=1e-3, adversarial_temperature=1.5)),
    "rotate_rmsprop_p2": dict(dim=96, P=2, k=2, B=70, E=2, S=1, model="RotatE", epochs=2, optimizer="RMSprop",
                              sb=25, train=dict(margin=9.0, log_frequency=3)),
}


def optimizer_of(gv, name):
    otype, lr, wd, a, b, eps = O.OPTIMIZERS[name]
    kwargs = {"SGD": {}, "Momentum": dict(momentum=a), "AdaGrad": dict(epsilon=eps),
              "RMSprop": dict(alpha=a, epsilon=eps), "Adam": dict(beta1=a, beta2=b, epsilon=eps)}[name]
    return getattr(gv.optimizer, name)(lr, wd, **kwargs)


def make_pair(cfg):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.KnowledgeGraph()
    graph.load(TOY_KG)
    solver = gv.solver.KnowledgeGraphSolver(cfg["dim"], device_ids=[0], num_sampler_per_worker=cfg["S"])
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"capture_negatives", 1))
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"train_num_groups", 1))
    solver.build(graph, optimizer_of(gv, cfg["optimizer"]), cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    ograph = K.OracleKnowledgeGraph(TOY_KG)
    osolver = K.OracleKGSolver(ograph, cfg["dim"], 1, cfg["S"])
    osolver.build(cfg["optimizer"], cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    return gv, _lib, graph, solver, ograph, osolver


def train_kwargs(cfg):
    kwargs = dict(model=cfg["model"], num_epoch=cfg["epochs"], resume=False, relation_lr_multiplier=1.0, margin=12.0,
                  l3_regularization=2e-3, sample_batch_size=cfg["sb"], positive_reuse=1, adversarial_temperature=2.0,
                  log_frequency=2)
    kwargs.update(cfg["train"])
    return kwargs


def product_begin(_lib, solver, kw):
    _lib.check(_lib.lib.gv_kg_solver_train_begin(
        solver._handle, kw["model"].encode(), kw["num_epoch"], int(kw["resume"]), kw["relation_lr_multiplier"],
        kw["margin"], kw["l3_regularization"], kw["sample_batch_size"], kw["positive_reuse"],
        kw["adversarial_temperature"], kw["log_frequency"]))


def product_pool(_lib, solver, side, head, tail, size):
    out = np.zeros((size, 3), dtype=np.uint32)
    assert _lib.lib.gv_kg_solver_pool(solver._handle, side, head, tail, out.ctypes.data) == size
    return out


@pytest.mark.parametrize("case", list(CASES))
def test_kg_solver_matches_oracle_step_by_step(case):
    cfg = CASES[case]
    gv, _lib, graph, solver, ograph, osolver = make_pair(cfg)
    n = graph.num_vertex
    part_of, local_of = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    _lib.lib.gv_kg_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data)
    opart, olocal = osolver.locations()
    np.testing.assert_array_equal(part_of, opart.astype(np.uint32))
    np.testing.assert_array_equal(local_of, olocal)

    kw = train_kwargs(cfg)
    product_begin(_lib, solver, kw)
    osolver.train_begin(**kw)
    info = osolver.info()
    for key in ("num_partition", "episode_size", "batch_size", "num_batch", "shuffle_partition"):
        assert getattr(solver, key) == info[key], key
    # identical initial embeddings: the same engine, the same draws
    size = info["episode_size"] * info["batch_size"]
    P = info["num_partition"]

    def check_pools(side):
        for h in range(P):
            for t in range(P):
                np.testing.assert_array_equal(product_pool(_lib, solver, side, h, t, size), osolver.pool(side, h, t),
                                              err_msg="pool %d block (%d, %d)" % (side, h, t))

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
        oinfo = osolver.info()
        check_pools(oinfo["pool_id"] ^ 1)
        assert solver.assignment_offset == oinfo["assignment_offset"]
        negatives = np.zeros(cfg["B"] * cfg["k"], dtype=np.uint32)
        assert _lib.lib.gv_kg_solver_last_negatives(solver._handle, negatives.ctypes.data) == negatives.size
        np.testing.assert_array_equal(negatives, osolver.last_negatives())
    assert episodes >= 1
    _lib.check(_lib.lib.gv_kg_solver_train_end(solver._handle))
    assert solver.batch_id == osolver.info()["batch_id"]
    np.testing.assert_allclose(solver.entity_embeddings, osolver.entity_embeddings, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.relation_embeddings, osolver.relation_embeddings, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.logged_loss, osolver.logged_loss(), rtol=1e-3, atol=1e-6)
    triplets = np.stack([np.random.RandomState(3).randint(0, n, 400), np.random.RandomState(4).randint(0, n, 400),
                         np.random.RandomState(5).randint(0, graph.num_relation, 400)], axis=1).astype(np.uint32)
    np.testing.assert_allclose(solver.predict(triplets), osolver.predict(triplets), rtol=1e-3, atol=1e-4)


def test_second_train_call_resumes_and_keeps_the_workers_relation_moments():
    """train(); train(resume=True); train(resume=False) on one build(): the entity moments follow `resume`, the
    worker-resident relation moments and the loss buffer survive every call (core/solver.h:1326,1378-1385)."""
    cfg = CASES["rotate_adam_p1"]
    gv, _lib, graph, solver, ograph, osolver = make_pair(cfg)
    kw = train_kwargs(cfg)
    for resume in (False, True, False):
        kw["resume"] = resume
        product_begin(_lib, solver, kw)
        osolver.train_begin(**kw)
        while _lib.lib.gv_kg_solver_train_episode(solver._handle) == 1:
            assert osolver.train_episode()
        assert not osolver.train_episode()
        _lib.check(_lib.lib.gv_kg_solver_train_end(solver._handle))
        np.testing.assert_allclose(solver.entity_embeddings, osolver.entity_embeddings, rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(solver.relation_embeddings, osolver.relation_embeddings, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.logged_loss, osolver.logged_loss(), rtol=1e-3, atol=1e-6)


def test_full_grid_training_learns_the_toy_graph():
    """Every thread group of the device at once (Hogwild, like the reference's warps): true triplets must end up
    scoring far above corrupted ones."""
    import graphvite_b200 as gv
    graph = gv.graph.KnowledgeGraph()
    graph.load(TOY_KG)
    solver = gv.solver.KnowledgeGraphSolver(64, device_ids=[0])
    solver.build(graph, gv.optimizer.Adam(5e-3), num_negative=8, batch_size=256, episode_size=2)
    solver.train("RotatE", num_epoch=60, margin=6.0, sample_batch_size=100, log_frequency=50)
    ograph = K.OracleKnowledgeGraph(TOY_KG)
    h, t, r, _, _ = ograph.flat()
    true = np.stack([h, t, r], axis=1).astype(np.uint32)
    rng = np.random.RandomState(0)
    corrupted = true.copy()
    corrupted[:, 1] = rng.randint(0, graph.num_vertex, len(true))
    positive, negative = solver.predict(true), solver.predict(corrupted)
    assert np.isfinite(positive).all() and np.isfinite(negative).all()
    assert positive.mean() > negative.mean() + 1.0
    assert "KnowledgeGraphSolver<64" in repr(solver) and "tied weights: yes" in repr(solver)


def test_error_paths():
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    graph = gv.graph.KnowledgeGraph()
    graph.load(TOY_KG)
    with pytest.raises(ValueError):
        gv.solver.KnowledgeGraphSolver(100)
    solver = gv.solver.KnowledgeGraphSolver(32, device_ids=[0])
    with pytest.raises(_lib.GVError, match="must be built"):
        solver.train("RotatE", num_epoch=1)
    solver.build(graph, num_negative=2, batch_size=50, episode_size=1)
    assert solver.optimizer.type in ("Default", "Adam")
    with pytest.raises(_lib.GVError, match="multiple of 2"):
        solver.build(graph, num_partition=3, num_negative=2, batch_size=50, episode_size=1)
    solver.build(graph, num_negative=2, batch_size=50, episode_size=1)
    with pytest.raises(_lib.GVError, match="Invalid model"):
        solver.train("LINE", num_epoch=1)
    with pytest.raises(_lib.GVError, match="shape"):
        solver.predict(np.zeros((4, 2), dtype=np.uint32))


def test_config_driven_knowledge_graph_run(tmp_path):
    """The reference's config/knowledge_graph/*.yaml layout through the `run` entry: train on the toy graph,
    filtered link prediction (MR / MRR / HITS@k), entity prediction, save_model / load_model by name."""
    import yaml
    from graphvite_b200 import application, cmd
    lines = [l for l in open(TOY_KG) if l.strip() and not l.startswith("#")]
    test_file, model_file = tmp_path / "test.txt", tmp_path / "model.pkl"
    test_file.write_text("".join(lines[:40]))
    config = {
        "application": "knowledge graph",
        "resource": {"gpus": [0], "cpu_per_gpu": "auto", "dim": 32},
        "graph": {"file_name": TOY_KG},
        "build": {"optimizer": {"type": "Adam", "lr": 5e-3, "weight_decay": 0}, "num_partition": "auto",
                  "num_negative": 8, "batch_size": 256, "episode_size": 2},
        "train": {"model": "RotatE", "num_epoch": 40, "margin": 6, "sample_batch_size": 100,
                  "adversarial_temperature": 2, "log_frequency": 100},
        "evaluate": {"task": "link prediction", "file_name": str(test_file), "filter_files": [TOY_KG]},
        "save": {"file_name": str(model_file)},
    }
    config_file = tmp_path / "rotate_toy.yaml"
    config_file.write_text(yaml.safe_dump(config))
    app, results = cmd.run_main(cmd.get_parser().parse_args(["run", str(config_file)]))
    metrics = results[0]
    assert set(metrics) == {"MR", "MRR", "HITS@1", "HITS@3", "HITS@10"}
    num_entity = app.graph.num_vertex
    assert 1 <= metrics["MR"] < num_entity / 4 and metrics["MRR"] > 4.0 / num_entity  # far better than chance
    assert metrics["HITS@1"] <= metrics["HITS@3"] <= metrics["HITS@10"] <= 1

    tokens = lines[0].split()
    recalls = app.entity_prediction(H=[tokens[0]], R=[tokens[1]], target="tail", k=5)
    assert len(recalls) == 1 and len(recalls[0]) == 5
    assert all(name in app.graph.entity2id for name, _ in recalls[0])
    assert recalls[0][0][1] >= recalls[0][-1][1]

    other = application.Application("knowledge graph", 32, gpus=[0])
    other.load(file_name=TOY_KG)
    other.build(num_negative=8, batch_size=256, episode_size=2)
    other.load_model(str(model_file))
    np.testing.assert_array_equal(other.solver.entity_embeddings, app.solver.entity_embeddings)
    np.testing.assert_array_equal(other.solver.relation_embeddings, app.solver.relation_embeddings)


# ---- the product against the REFERENCE's own solver (tests/golden/kg_solver_*.npz, recorded by
# oracle/make_golden_kg.py from the unmodified reference; its samples of a batch are processed in order there,
# like the one-group mode here) ---------------------------------------------------------------------------------
import glob

SOLVER_FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "kg_solver_*.npz")))


@pytest.mark.skipif(not SOLVER_FILES, reason="no kg_solver_*.npz fixtures")
@pytest.mark.parametrize("path", SOLVER_FILES, ids=[os.path.basename(p)[10:-4] for p in SOLVER_FILES])
def test_kg_solver_matches_the_reference_solver(path):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    g = np.load(path)
    cfg = {key[4:]: g[key].item() for key in g.files if key.startswith("cfg_")}
    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.KnowledgeGraph()
    graph.load(TOY_KG)
    solver = gv.solver.KnowledgeGraphSolver(cfg["dim"], device_ids=[0], num_sampler_per_worker=cfg["S"])
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"capture_negatives", 1))
    _lib.check(_lib.lib.gv_kg_solver_set_option(solver._handle, b"train_num_groups", 1))
    solver.build(graph, optimizer_of(gv, cfg["optimizer"]), cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    solver.train(cfg["model"], num_epoch=cfg["epochs"], relation_lr_multiplier=cfg["rlm"], margin=cfg["margin"],
                 l3_regularization=cfg["l3"], sample_batch_size=cfg["sbs"], positive_reuse=cfg["reuse"],
                 adversarial_temperature=cfg["temperature"], log_frequency=100)
    info = g["info"]
    P, E, B = int(info[0]), int(info[1]), int(info[2])
    assert (solver.num_partition, solver.episode_size, solver.batch_size) == (P, E, B)
    assert (solver.num_batch, solver.batch_id, solver.pool_id) == (int(info[3]), int(info[4]), int(info[5]))
    assert (solver.assignment_offset, solver.shuffle_partition) == (int(info[7]), int(info[9]))
    part_of, local_of = np.zeros(graph.num_vertex, dtype=np.uint32), np.zeros(graph.num_vertex, dtype=np.uint32)
    _lib.lib.gv_kg_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data)
    np.testing.assert_array_equal(part_of, g["part_of"].astype(np.uint32))
    np.testing.assert_array_equal(local_of, g["local_of"])
    for side in range(2):  # both sample pools as the reference left them: every sampled triplet, bit for bit
        for h in range(P):
            for t in range(P):
                np.testing.assert_array_equal(product_pool(_lib, solver, side, h, t, E * B), g["pools"][side, h, t],
                                              err_msg="pool %d block (%d, %d)" % (side, h, t))
    negatives = np.zeros(B * cfg["k"], dtype=np.uint32)
    assert _lib.lib.gv_kg_solver_last_negatives(solver._handle, negatives.ctypes.data) == negatives.size
    np.testing.assert_array_equal(negatives, g["negatives"])
    if P == 1:
        # the north-star tolerance: 1e-3 relative on the L2 norms of the final embeddings ...
        for ours, name in ((solver.entity_embeddings, "entity_0"), (solver.relation_embeddings, "relation_0")):
            assert np.linalg.norm(ours) == pytest.approx(np.linalg.norm(g[name]), rel=1e-3), name
        # ... and element by element (the fixtures carry the host's libm / no-FMA rounding, the device its own)
        np.testing.assert_allclose(solver.entity_embeddings, g["entity_0"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(solver.relation_embeddings, g["relation_0"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(solver.predict(g["triplets"]), g["logits"], rtol=2e-3, atol=2e-4)
    else:
        # With several partitions the reference's partition cache can hold a second, stale copy of an entity
        # partition (a "tail hit" keeps the trained tail copy on the device while the head copy of the same
        # partition is reloaded from host memory, core/solver.h:1436-1476) and the later write-back overwrites the
        # earlier: updates are lost.  Blocks have one owner here, so only the magnitude is comparable.
        for ours, name in ((solver.entity_embeddings, "entity_0"), (solver.relation_embeddings, "relation_0")):
            assert np.linalg.norm(ours) == pytest.approx(np.linalg.norm(g[name]), rel=0.05), name
