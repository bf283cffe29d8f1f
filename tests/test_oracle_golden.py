"""Pins the oracle (oracle/gv_oracle.cpp) against golden vectors produced by the UNMODIFIED
reference on a B200 (oracle/make_golden.py -> tests/golden/*.npz).  CPU only.

Integer / index outputs must be bit-exact.  The train kernels are compared with
rtol 5e-4 / atol 5e-6: the oracle restates the reference's evaluation order but uses host libm
(expf / logf / sqrtf) and explicit fmaf where nvcc contracts on its own."""
import numpy as np
import pytest

import oracle_lib as O

SOLVER_CASES = ["line_p1", "line_p2_s3", "deepwalk_p1", "edge_p2", "line_p3_adam", "node2vec_p2", "node2vec_p1"]


def load(golden_dir, name):
    return np.load("%s/%s.npz" % (golden_dir, name))


def test_global_engine_seeds(golden_dir, toy_graph_file):
    """std::mt19937 + uniform_int_distribution<unsigned long long>: sampler seeds then worker seeds"""
    golden = load(golden_dir, "curand")["seeds"]
    solver = O.OracleSolver(O.OracleGraph(toy_graph_file), 32, num_worker=2, num_sampler_per_worker=2)
    sampler, worker = solver.seeds()
    np.testing.assert_array_equal(np.concatenate([sampler, worker]), golden)


def test_curand_host_generator_reproduces_device_stream(golden_dir):
    golden = load(golden_dir, "curand")
    small = O.curand_uniform_double(golden["seeds"][0], golden["small_chunks"])
    np.testing.assert_array_equal(small, golden["small"])
    big = O.curand_uniform_double(golden["big_seed"], [5000000, 5000000])
    np.testing.assert_array_equal(big[:8192], golden["big_head"])
    np.testing.assert_array_equal(big[5000000 - 2048:5000000 + 2048], golden["big_mid"])
    np.testing.assert_array_equal(big[-4096:], golden["big_tail"])
    np.testing.assert_allclose([big[:5000000].sum(), big[5000000:].sum()], golden["big_sum"], rtol=1e-12)
    assert big.min() > 0.0 and big.max() <= 1.0  # (0, 1]


def test_alias_table_build_and_sampling(golden_dir):
    golden = load(golden_dir, "alias")
    prob, alias = O.alias_build(golden["weights"])
    np.testing.assert_array_equal(prob, golden["prob"])
    np.testing.assert_array_equal(alias, golden["alias"])
    np.testing.assert_array_equal(O.alias_sample(prob, alias, golden["random"], gpu_path=False), golden["cpu_samples"])
    np.testing.assert_array_equal(O.alias_sample(prob, alias, golden["random"], gpu_path=True), golden["gpu_samples"])
    uprob, ualias = O.alias_build(np.ones(37, dtype=np.float32))
    np.testing.assert_array_equal(uprob, golden["uniform_prob"])
    np.testing.assert_array_equal(ualias, golden["uniform_alias"])


@pytest.mark.parametrize("undirected", [1, 0])
@pytest.mark.parametrize("normalization", [0, 1])
def test_graph_loading_and_flatten(golden_dir, toy_graph_file, undirected, normalization):
    golden = load(golden_dir, "graph_u%d_n%d" % (undirected, normalization))
    graph = O.OracleGraph(toy_graph_file, bool(undirected), bool(normalization))
    assert graph.num_vertex == golden["num_vertex"] and graph.num_edge == golden["num_edge"]
    u, v, w, offsets = graph.flat()
    np.testing.assert_array_equal(u, golden["u"])
    np.testing.assert_array_equal(v, golden["v"])
    np.testing.assert_array_equal(w, golden["w"])
    np.testing.assert_array_equal(graph.vertex_weights(), golden["vertex_weights"])


@pytest.mark.parametrize("case", SOLVER_CASES)
def test_solver_integer_state(golden_dir, toy_graph_file, case):
    """partition, both sample pools after the whole run, the last batch's negatives, the edge table"""
    golden = load(golden_dir, "solver_" + case)
    cfg = {key[4:]: golden[key].item() for key in golden.files if key.startswith("cfg_")}
    graph = O.OracleGraph(toy_graph_file)
    solver = O.OracleSolver(graph, cfg["dim"], 1, cfg["S"])
    solver.build(cfg["optimizer"], cfg["P"], cfg["k"], cfg["B"], cfg["E"])
    solver.train(model=cfg["model"], num_epoch=cfg["epochs"], augmentation_step=cfg["aug"],
                 random_walk_length=cfg["L"], random_walk_batch_size=cfg["wb"], p=cfg.get("p", 1.0),
                 q=cfg.get("q", 1.0))
    assert list(solver.info().values()) == golden["info"].tolist()
    part_of, local_of = solver.locations()
    np.testing.assert_array_equal(part_of, golden["part_of"])
    np.testing.assert_array_equal(local_of, golden["local_of"])
    P = cfg["P"]
    for side in range(2):
        for h in range(P):
            for t in range(P):
                np.testing.assert_array_equal(solver.pool(side, h, t), golden["pools"][side, h, t])
    np.testing.assert_array_equal(solver.last_negatives(cfg["B"], cfg["k"]), golden["negatives"])
    prob, alias = solver.edge_table()
    np.testing.assert_array_equal(prob, golden["edge_prob"])
    np.testing.assert_array_equal(alias, golden["edge_alias"])


@pytest.mark.parametrize("dim", [32, 128])
@pytest.mark.parametrize("opt", list(O.OPTIMIZERS))
def test_train_kernels_race_free(golden_dir, dim, opt):
    """gpu::graph::train / train_1_moment / train_2_moment on a batch with pairwise distinct rows"""
    golden = load(golden_dir, "kernel_d%d_%s" % (dim, opt))
    otype, lr, wd, a, b, eps, negative_weight = [float(x) for x in golden["hyper"]]
    num_moment = 0 if otype == 0 else (2 if otype == 4 else 1)
    vertex, context = golden["before_vertex"].copy(), golden["before_context"].copy()
    moments = [golden["before_" + key].copy() if i // 2 < num_moment else None
               for i, key in enumerate(("vm1", "cm1", "vm2", "cm2"))]
    loss = O.train_batch(dim, vertex, context, moments, golden["batch"], golden["negatives"],
                         (int(otype), lr, wd, a, b, eps), negative_weight)
    np.testing.assert_allclose(vertex, golden["after_vertex"], rtol=5e-4, atol=5e-6)
    np.testing.assert_allclose(context, golden["after_context"], rtol=5e-4, atol=5e-6)
    np.testing.assert_allclose(loss, golden["loss"], rtol=5e-4, atol=5e-6)
    for key, moment in zip(("vm1", "cm1", "vm2", "cm2"), moments):
        if moment is not None:
            np.testing.assert_allclose(moment, golden["after_" + key], rtol=5e-4, atol=5e-6)
    assert np.abs(golden["after_vertex"] - golden["before_vertex"]).max() > 1e-3  # the batch did something


def test_predict_matches_reference(golden_dir, toy_graph_file):
    golden = load(golden_dir, "solver_line_p1")
    logits = O.predict_batch(32, golden["vertex"], golden["context"], golden["pairs"][:, ::-1].copy())
    np.testing.assert_allclose(logits, golden["logits"], rtol=1e-4, atol=1e-6)


def test_schedule_and_lr():
    np.testing.assert_array_equal(O.schedule(1, 1), [[[0, 0]]])
    steps = O.schedule(4, 2)
    assert steps.shape == (8, 2, 2)
    for step in steps:  # orthogonal blocks: no head or tail partition is shared inside a step
        assert len(set(step[:, 0])) == 2 and len(set(step[:, 1])) == 2
    assert {(h, t) for step in steps for h, t in step} == {(h, t) for h in range(4) for t in range(4)}
    lib = O.lib()
    assert lib.og_lr(1, 0.025, 0, 100) == np.float32(0.025)
    assert abs(lib.og_lr(1, 0.025, 50, 100) - 0.0125) < 1e-9
    assert abs(lib.og_lr(1, 0.025, 100, 100) - 0.025 * 1e-4) < 1e-12  # floor at 1e-4
    assert lib.og_lr(0, 0.025, 77, 100) == np.float32(0.025)
