"""GPU parity of the host runtime + device samplers (gv_solver_*) against the oracle solver:
sample pools and negative indices bit-exact for the default engine seed, embeddings within
float tolerance in the single-warp (sequential) mode, and norms / loss in the Hogwild mode."""
import ctypes

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

CASES = {
    "line_p1": dict(dim=32, P=1, k=1, B=500, E=4, S=1, model="LINE", epochs=4, aug=2, L=5, wb=10, optimizer="SGD"),
    "line_p2_s3": dict(dim=32, P=2, k=3, B=400, E=3, S=3, model="LINE", epochs=8, aug=3, L=7, wb=8,
                       optimizer="SGD"),
    "deepwalk_p1": dict(dim=128, P=1, k=2, B=300, E=5, S=2, model="DeepWalk", epochs=3, aug=4, L=9, wb=6,
                        optimizer="SGD"),
    "edge_p2": dict(dim=32, P=2, k=1, B=500, E=2, S=2, model="LINE", epochs=3, aug=1, L=5, wb=10, optimizer="SGD"),
    "line_p3_adam": dict(dim=32, P=3, k=2, B=300, E=2, S=1, model="LINE", epochs=4, aug=2, L=6, wb=10,
                         optimizer="Adam"),
    "line_odd_walks": dict(dim=64, P=2, k=1, B=450, E=2, S=1, model="LINE", epochs=2, aug=2, L=3, wb=7,
                           optimizer="Momentum"),
    "node2vec_p2": dict(dim=32, P=2, k=1, B=400, E=3, S=2, model="node2vec", epochs=5, aug=3, L=8, wb=10,
                        optimizer="SGD", p=0.5, q=2.0),
    "node2vec_p1": dict(dim=32, P=1, k=2, B=500, E=2, S=1, model="node2vec", epochs=3, aug=2, L=5, wb=10,
                        optimizer="SGD", p=4.0, q=0.25),
}


def make_product(cfg, toy_graph_file, single_warp):
    import graphvite_b200 as gv
    from graphvite_b200 import _lib
    _lib.lib.gv_reset_global_engine(5489)
    graph = gv.graph.Graph()
    graph.load(toy_graph_file)
    solver = gv.solver.GraphSolver(cfg["dim"], device_ids=[0], num_sampler_per_worker=cfg["S"])
    _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"capture_negatives", 1))
    if single_warp:
        _lib.check(_lib.lib.gv_solver_set_option(solver._handle, b"train_num_warps", 1))
    name = cfg["optimizer"]
    otype, lr, wd, a, b, eps = O.OPTIMIZERS[name]
    kwargs = {"SGD": {}, "Momentum": dict(momentum=a), "Adam": dict(beta1=a, beta2=b, epsilon=eps)}[name]
    optimizer = getattr(gv.optimizer, name)(lr, wd, **kwargs)
    solver.build(graph, optimizer, cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    return gv, _lib, graph, solver


def make_oracle(cfg, toy_graph_file):
    graph = O.OracleGraph(toy_graph_file)
    solver = O.OracleSolver(graph, cfg["dim"], 1, cfg["S"])
    solver.build(cfg["optimizer"], cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    return graph, solver


def product_pool(_lib, solver, side, head, tail, size):
    out = np.zeros((size, 2), dtype=np.uint32)
    count = _lib.lib.gv_solver_pool(solver._handle, side, head, tail, out.ctypes.data)
    assert count == size
    return out


def train_args(cfg):
    return (cfg["model"].encode(), cfg["epochs"], 0, cfg["aug"], cfg["L"], cfg["wb"], 0, cfg.get("p", 1.0),
            cfg.get("q", 1.0), 1, 0.75, 5.0, 1000)


@pytest.mark.parametrize("case", list(CASES))
def test_solver_matches_oracle_step_by_step(case, toy_graph_file):
    cfg = CASES[case]
    gv, _lib, graph, solver = make_product(cfg, toy_graph_file, single_warp=True)
    ograph, osolver = make_oracle(cfg, toy_graph_file)
    # partition: head_locations bit-exact
    n = graph.num_vertex
    part_of, local_of = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    _lib.lib.gv_solver_locations(solver._handle, part_of.ctypes.data, local_of.ctypes.data)
    opart, olocal = osolver.locations()
    np.testing.assert_array_equal(part_of, opart.astype(np.uint32))
    np.testing.assert_array_equal(local_of, olocal)

    _lib.check(_lib.lib.gv_solver_train_begin(solver._handle, *train_args(cfg)))
    osolver.train_begin(cfg["model"], cfg["epochs"], False, cfg["aug"], cfg["L"], cfg["wb"], 0, cfg.get("p", 1.0),
                        cfg.get("q", 1.0))
    info = osolver.info()
    for key in ("num_partition", "episode_size", "batch_size", "augmentation_step", "shuffle_base", "num_batch"):
        assert getattr(solver, key) == info[key], key
    size = info["episode_size"] * info["batch_size"]
    P = info["num_partition"]

    def check_pools(side):
        for h in range(P):
            for t in range(P):
                np.testing.assert_array_equal(product_pool(_lib, solver, side, h, t, size), osolver.pool(side, h, t),
                                              err_msg="pool %d block (%d, %d)" % (side, h, t))

    check_pools(1)  # the first fill goes to pool_id ^ 1 = 1
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
        negatives = np.zeros(cfg["B"] * cfg["k"], dtype=np.uint32)
        assert _lib.lib.gv_solver_last_negatives(solver._handle, negatives.ctypes.data) == negatives.size
        np.testing.assert_array_equal(negatives, osolver.last_negatives(cfg["B"], cfg["k"]))
    assert episodes >= 1
    _lib.check(_lib.lib.gv_solver_train_end(solver._handle))
    assert solver.batch_id == osolver.info()["batch_id"]
    # sequential mode: embeddings equal up to fp32 summation order
    np.testing.assert_allclose(solver.vertex_embeddings, osolver.embeddings(0), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.context_embeddings, osolver.embeddings(1), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(solver.logged_loss, osolver.logged_loss(), rtol=1e-3, atol=1e-6)
    # predict on the written-back matrices
    pairs = np.random.RandomState(3).randint(0, n, (500, 2)).astype(np.uint32)
    np.testing.assert_allclose(solver.predict(pairs), osolver.predict(pairs), rtol=1e-3, atol=1e-5)


def test_resume_and_numpy_views(toy_graph_file):
    cfg = CASES["line_p1"]
    gv, _lib, graph, solver = make_product(cfg, toy_graph_file, single_warp=True)
    solver.train("LINE", 2, False, 2, 5, 10)
    first = np.array(solver.vertex_embeddings)
    assert solver.batch_id == 8  # 6 batches asked for, whole episodes of 4 trained (core/solver.h:629)
    view = solver.vertex_embeddings
    view[0, :] = 0.25  # views alias solver memory (bind.h:90-106); resume starts from the edited matrix
    solver.train("LINE", 2, True, 2, 5, 10)
    assert solver.batch_id == 16  # num_batch = 8 + 6 = 14 -> two more episodes
    assert not np.allclose(first, solver.vertex_embeddings)
    assert solver.vertex_embeddings.shape == (graph.num_vertex, cfg["dim"])
    solver.clear()
    assert solver.vertex_embeddings.shape == (graph.num_vertex, cfg["dim"])


def test_errors_are_reported_not_fatal(toy_graph_file):
    import graphvite_b200 as gv
    graph = gv.graph.Graph()
    graph.load(toy_graph_file)
    solver = gv.solver.GraphSolver(32, device_ids=[0])
    with pytest.raises(gv.GVError):
        solver.train("LINE")  # not built
    solver.build(graph, batch_size=500, episode_size=3)
    with pytest.raises(gv.GVError):
        solver.train("TransE", 1)  # invalid model
    with pytest.raises(gv.GVError):
        solver.train("LINE", 1, augmentation_step=50, random_walk_length=40)  # walk shorter than augmentation
    with pytest.raises(ValueError):
        gv.solver.GraphSolver(100)


def test_config_driven_run(tmp_path, toy_graph_file):
    """`graphvite run config.yaml` semantics (cmd.py:140-163) on the toy graph: load, build, train,
    link prediction with a filter file, save_model / load_model round trip"""
    import yaml
    import graphvite_b200 as gv
    from graphvite_b200 import cmd
    edges = [line.split("#")[0].split()[:2] for line in open(toy_graph_file)]
    edges = [e for e in edges if len(e) == 2]
    rng = np.random.RandomState(1)
    test = tmp_path / "test.txt"
    with open(test, "w") as fout:
        for u, v in edges[:200]:
            fout.write("%s\t%s\t1\n" % (u, v))
        for _ in range(200):
            fout.write("n%d\tn%d\t0\n" % (rng.randint(300), rng.randint(300)))
    train = tmp_path / "train.txt"
    with open(train, "w") as fout:
        for u, v in edges[200:]:
            fout.write("%s %s\n" % (u, v))
    model = tmp_path / "model.pkl"
    cfg = {"application": "graph", "resource": {"gpus": [0], "cpu_per_gpu": 3, "dim": 32},
           "format": {"delimiters": " \t\r\n", "comment": "#"},
           "graph": {"file_name": str(train), "as_undirected": True},
           "build": {"optimizer": {"type": "SGD", "lr": 0.025, "weight_decay": 0.005}, "num_partition": "auto",
                     "num_negative": 1, "batch_size": 500, "episode_size": 4},
           "train": {"model": "LINE", "num_epoch": 300, "augmentation_step": 2, "random_walk_length": 5,
                     "random_walk_batch_size": 10},
           "evaluate": [{"task": "link prediction", "file_name": str(test), "filter_file": str(train)}],
           "save": {"file_name": str(model)}}
    path = tmp_path / "config.yaml"
    path.write_text(yaml.safe_dump(cfg))
    app, results = cmd.run_main(cmd.get_parser().parse_args(["run", str(path)]))
    assert app.solver.num_sampler == 2  # cpu_per_gpu - 1 (application.py:283-286)
    assert 0.55 < results[0]["AUC"] <= 1.0
    other = gv.application.Application("graph", 32, gpus=[0])
    other.load(file_name=str(train)).build(batch_size=500, episode_size=4).load_model(str(model))
    np.testing.assert_array_equal(other.solver.vertex_embeddings, app.solver.vertex_embeddings)
    assert abs(other.link_prediction(file_name=str(test), filter_file=str(train))["AUC"] - results[0]["AUC"]) < 1e-6
