"""Generate the knowledge-graph golden vectors (tests/golden/kg_*.npz) by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  The reference's solver needs a CUDA device:
    gpurun -- 'python oracle/make_golden_kg.py gpurun_out/golden_kg'
then copy gpurun_out/golden_kg/*.npz into tests/golden/ and commit them.  Without a GPU,
    make -C oracle ref_emu && python oracle/make_golden_kg.py --emulated <out_dir>
runs the very same unmodified reference code under the CUDA emulation of tests/emu (see
oracle/emulate_reference.py): integer outputs (pools, negatives, partitions) are what a GPU produces; float
outputs carry the host's libm / no-FMA rounding instead of the device's (differences of a few ulp).  The reference is driven through
oracle/_ref/libref_harness_kg.so (oracle/ref_harness_kg.cu, built by `make -C oracle ref` in the authoring
container; /root/reference does not exist on the GPU box).  The toy knowledge graph the cases run on is
generated here deterministically and committed as tests/golden/toy_kg.txt.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
c = ctypes

OPTIMIZERS = {
    # name: (type, lr, weight_decay, a, b, epsilon) -- same table as oracle/make_golden.py
    "SGD": (0, 0.025, 0.005, 0.0, 0.0, 0.0),
    "Momentum": (1, 0.01, 0.001, 0.9, 0.0, 0.0),
    "AdaGrad": (2, 0.05, 0.001, 0.0, 0.0, 1e-10),
    "RMSprop": (3, 0.001, 0.001, 0.99, 0.0, 1e-8),
    "Adam": (4, 0.001, 0.001, 0.9, 0.999, 1e-8),
}
MODELS = ("TransE", "DistMult", "ComplEx", "SimplE", "RotatE", "QuatE")

# solver cases, one GPU; each runs on tests/golden/toy_kg.txt
SOLVER_CASES = {
    "rotate_p1_adam": dict(dim=32, P=1, k=4, B=200, E=3, S=1, model="RotatE", epochs=4, optimizer="Adam", margin=6.0,
                           l3=2e-3, temperature=2.0, sbs=50, rlm=1.0, reuse=1),
    "transe_p2_sgd": dict(dim=32, P=2, k=3, B=150, E=2, S=2, model="TransE", epochs=3, optimizer="SGD", margin=4.0,
                          l3=2e-3, temperature=0.0, sbs=40, rlm=1.0, reuse=1),
    "distmult_p1_adagrad": dict(dim=64, P=1, k=2, B=100, E=4, S=1, model="DistMult", epochs=3, optimizer="AdaGrad",
                                margin=12.0, l3=2e-3, temperature=1.0, sbs=64, rlm=2.0, reuse=1),
    "complex_p2_adam": dict(dim=32, P=2, k=2, B=120, E=2, S=1, model="ComplEx", epochs=4, optimizer="Adam",
                            margin=12.0, l3=1e-3, temperature=2.0, sbs=30, rlm=1.0, reuse=1),
    "simple_p1_momentum": dict(dim=32, P=1, k=3, B=150, E=2, S=3, model="SimplE", epochs=2, optimizer="Momentum",
                               margin=12.0, l3=2e-3, temperature=0.5, sbs=50, rlm=1.0, reuse=2),
    "quate_p1_adam": dict(dim=32, P=1, k=3, B=150, E=2, S=1, model="QuatE", epochs=3, optimizer="Adam", margin=12.0,
                          l3=2e-3, temperature=2.0, sbs=45, rlm=1.0, reuse=1),
    "rotate_p4_rmsprop": dict(dim=32, P=4, k=2, B=60, E=2, S=1, model="RotatE", epochs=6, optimizer="RMSprop",
                              margin=9.0, l3=2e-3, temperature=2.0, sbs=25, rlm=0.5, reuse=1),
}


EMULATED = "--emulated" in sys.argv  # the reference executed by tests/emu's CUDA emulation (`make -C oracle ref_emu`)


def load_harness():
    if EMULATED:
        # two-pass build (oracle/Makefile, ref_emu): launches swap the host-pass kernel for its device-pass twin
        os.environ["GV_EMU_DEVICE_LIBRARY"] = os.path.join(HERE, "_ref", "libref_harness_kg_emu_device.so")
        os.environ["GV_EMU_DEVICE_NAMESPACE"] = "graphvite=graphvite_device"
        lib = c.CDLL(os.path.join(HERE, "_ref", "libref_harness_kg_emu.so"), mode=c.RTLD_GLOBAL)
    else:
        lib = c.CDLL(os.path.join(HERE, "_ref", "libref_harness_kg.so"))
    V, I, F, U64, S = c.c_void_p, c.c_int, c.c_float, c.c_uint64, c.c_char_p
    lib.rk_graph_load.restype = V
    lib.rk_graph_load.argtypes = [S, I]
    lib.rk_graph_free.argtypes = [V]
    lib.rk_graph_sizes.argtypes = [V, V]
    lib.rk_graph_flat.argtypes = [V] * 6
    lib.rk_solver_new.restype = V
    lib.rk_solver_new.argtypes = [I, I, I, U64]
    lib.rk_solver_free.argtypes = [V]
    lib.rk_solver_build.argtypes = [V, V, I, I, F, F, F, F, F, I, I, I, I]
    lib.rk_solver_train.argtypes = [V, S, I, I, F, F, F, I, I, F, I]
    lib.rk_solver_info.argtypes = [V, V]
    lib.rk_solver_locations.argtypes = [V, V, V]
    lib.rk_solver_pool.argtypes = [V, I, I, I, V]
    lib.rk_solver_matrix.argtypes = [V, I, I, V]
    lib.rk_solver_last_negatives.argtypes = [V, V]
    lib.rk_solver_last_loss.argtypes = [V, V]
    lib.rk_solver_negative_table.argtypes = [V, V, V]
    lib.rk_solver_schedule.argtypes = [V, V]
    lib.rk_solver_predict.argtypes = [V, V, U64, V]
    lib.rk_kernel_train.argtypes = [S, I, I, F, F, F, F, F, U64, U64, U64, I] + [V] * 11 + [I, I, F, F, F, V]
    lib.rk_kernel_predict.argtypes = [S, I, U64, U64, V, V, V, I, F, V]
    return lib


def ptr(array):
    return array.ctypes.data if array is not None else None


def make_toy_kg(path):
    """Triplets with power-law heads, a few weighted lines, comments, a self loop and repeated triplets."""
    rng = np.random.RandomState(5)
    n, num_relation, m = 120, 7, 1400
    weights = np.arange(1, n + 1) ** -0.8
    cdf = np.cumsum(weights) / weights.sum()
    heads = np.searchsorted(cdf, rng.rand(m))
    relations = rng.randint(0, num_relation, m)
    tails = (heads * (relations + 2) + relations + rng.randint(0, 3, m)) % n
    lines = ["# toy knowledge graph for the golden vectors", ""]
    for i in range(m):
        h, r, t = int(heads[i]), int(relations[i]), int(tails[i])
        if i == 23:
            t = h
        if i % 9 == 0:
            lines.append("/m/%03d\t/r/%d\t/m/%03d\t%.2f" % (h, r, t, 0.5 + 1.5 * rng.rand()))
        elif i % 11 == 0:
            lines.append("/m/%03d /r/%d /m/%03d  # trailing comment" % (h, r, t))
        else:
            lines.append("/m/%03d /r/%d /m/%03d" % (h, r, t))
    with open(path, "w") as fout:
        fout.write("\n".join(lines) + "\n")


def write_graph(lib, out_dir, toy):
    for normalization in (0, 1):
        g = lib.rk_graph_load(toy.encode(), normalization)
        sizes = np.zeros(3, dtype=np.uint64)
        lib.rk_graph_sizes(g, ptr(sizes))
        n, m = int(sizes[0]), int(sizes[1])
        h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
        w, vw = np.zeros(m, dtype=np.float32), np.zeros(n, dtype=np.float32)
        lib.rk_graph_flat(g, ptr(h), ptr(t), ptr(r), ptr(w), ptr(vw))
        np.savez_compressed(os.path.join(out_dir, "kg_graph_n%d.npz" % normalization), sizes=sizes, h=h, t=t, r=r, w=w,
                            vertex_weights=vw)
        lib.rk_graph_free(g)


def write_solver_cases(lib, out_dir, toy, only):
    graph = lib.rk_graph_load(toy.encode(), 0)
    sizes = np.zeros(3, dtype=np.uint64)
    lib.rk_graph_sizes(graph, ptr(sizes))
    num_vertex, num_relation = int(sizes[0]), int(sizes[2])
    for name, cfg in SOLVER_CASES.items():
        if only and name not in only:
            continue
        lib.rk_reset_engine()
        solver = lib.rk_solver_new(cfg["dim"], 1, cfg["S"], 4 << 30)
        otype, lr, wd, a, b, eps = OPTIMIZERS[cfg["optimizer"]]
        lib.rk_solver_build(solver, graph, otype, 1, lr, wd, a, b, eps, cfg["P"], cfg["k"], cfg["B"], cfg["E"])
        lib.rk_solver_train(solver, cfg["model"].encode(), cfg["epochs"], 0, cfg["rlm"], cfg["margin"], cfg["l3"],
                            cfg["sbs"], cfg["reuse"], cfg["temperature"], 100)
        info = np.zeros(12, dtype=np.int32)
        lib.rk_solver_info(solver, ptr(info))
        P, E, B = int(info[0]), int(info[1]), int(info[2])
        part_of, local_of = np.zeros(num_vertex, dtype=np.int32), np.zeros(num_vertex, dtype=np.uint32)
        lib.rk_solver_locations(solver, ptr(part_of), ptr(local_of))
        pools = np.zeros((2, P, P, E * B, 3), dtype=np.uint32)
        for side in range(2):
            for h in range(P):
                for t in range(P):
                    lib.rk_solver_pool(solver, side, h, t, ptr(pools[side, h, t]))
        num_moment = {0: 0, 4: 2}.get(otype, 1)
        matrices = {}
        for which, rows, label in ((0, num_vertex, "entity"), (1, num_relation, "relation")):
            for order in range(num_moment + 1):
                out = np.zeros((rows, cfg["dim"]), dtype=np.float32)
                lib.rk_solver_matrix(solver, which, order, ptr(out))
                matrices["%s_%d" % (label, order)] = out
        negatives = np.zeros(B * cfg["k"], dtype=np.uint32)
        lib.rk_solver_last_negatives(solver, ptr(negatives))
        loss = np.zeros(B, dtype=np.float32)
        lib.rk_solver_last_loss(solver, ptr(loss))
        count = lib.rk_solver_negative_table(solver, None, None)
        prob, alias = np.zeros(count, dtype=np.float32), np.zeros(count, dtype=np.uint32)
        lib.rk_solver_negative_table(solver, ptr(prob), ptr(alias))
        schedule = np.zeros(4096, dtype=np.int32)
        steps = lib.rk_solver_schedule(solver, ptr(schedule))
        rng = np.random.RandomState(3)
        triplets = np.stack([rng.randint(0, num_vertex, 400), rng.randint(0, num_vertex, 400),
                             rng.randint(0, num_relation, 400)], axis=1).astype(np.uint32)
        logits = np.zeros(400, dtype=np.float32)
        lib.rk_solver_predict(solver, ptr(triplets), 400, ptr(logits))
        np.savez_compressed(os.path.join(out_dir, "kg_solver_%s.npz" % name), emulated=np.array(int(EMULATED)),
                            info=info, part_of=part_of,
                            local_of=local_of, pools=pools, negatives=negatives, loss=loss, negative_prob=prob,
                            negative_alias=alias, schedule=schedule[:steps * 2].reshape(steps, 1, 2), triplets=triplets,
                            logits=logits, **matrices, **{"cfg_" + k: np.array(v) for k, v in cfg.items()})
        lib.rk_solver_free(solver)
        print("kg solver case", name, "info", info.tolist(), flush=True)
    # the tied schedule for several (W, P); needs a solver object with that many workers, i.e. GPUs
    lib.rk_graph_free(graph)


def race_free_batch(rng, n, k, num_head, num_tail, num_relation):
    """every entity row and every relation row is named by at most one sample (targets included)"""
    head_rows = rng.permutation(num_head)
    tail_rows = rng.permutation(num_tail)
    relations = rng.permutation(num_relation)[:n]
    batch = np.zeros((n, 3), dtype=np.uint32)
    negatives = np.zeros((n, k), dtype=np.uint32)
    hi = ti = 0
    for i in range(n):
        batch[i] = (relations[i], tail_rows[ti], head_rows[hi])
        hi += 1
        ti += 1
        for s in range(k):
            if rng.rand() < 0.5:
                negatives[i, s] = head_rows[hi]
                hi += 1
            else:
                negatives[i, s] = num_head + tail_rows[ti]
                ti += 1
    assert hi <= num_head and ti <= num_tail
    return batch, negatives


def write_kernel_cases(lib, out_dir):
    for model in MODELS:
        margin_or_l3 = 6.0 if model in ("TransE", "RotatE") else 2e-3
        for oname, (otype, lr, wd, a, b, eps) in OPTIMIZERS.items():
            for dim, temperature in ((32, 1.5), (512, 0.0)):
                if dim == 512 and not (model == "RotatE" and oname in ("SGD", "Adam")):
                    continue
                rng = np.random.RandomState(1000 + 7 * dim + 13 * otype + len(model))
                n, k = 24, 3
                num_head, num_tail, num_relation = 130, 140, 40
                if dim == 512:  # keeps the fixture small
                    n, num_head, num_tail, num_relation = 12, 60, 70, 30
                scale = 0.3 if dim == 512 else 1.0
                head = ((rng.rand(num_head, dim) - 0.5) * scale).astype(np.float32)
                tail = ((rng.rand(num_tail, dim) - 0.5) * scale).astype(np.float32)
                relation = ((rng.rand(num_relation, dim) - 0.5) * 2.0).astype(np.float32)
                moments = [np.abs(rng.randn(rows, dim).astype(np.float32)) * 0.01
                           for rows in (num_head, num_tail, num_relation, num_head, num_tail, num_relation)]
                batch, negatives = race_free_batch(rng, n, k, num_head, num_tail, num_relation)
                before = dict(head=head.copy(), tail=tail.copy(), relation=relation.copy(), hm1=moments[0].copy(),
                              tm1=moments[1].copy(), rm1=moments[2].copy(), hm2=moments[3].copy(),
                              tm2=moments[4].copy(), rm2=moments[5].copy())
                loss = np.zeros(n, dtype=np.float32)
                status = lib.rk_kernel_train(model.encode(), dim, otype, lr, wd, a, b, eps, num_head, num_tail,
                                             num_relation, 0, ptr(head), ptr(tail), ptr(relation), ptr(moments[0]),
                                             ptr(moments[1]), ptr(moments[2]), ptr(moments[3]), ptr(moments[4]),
                                             ptr(moments[5]), ptr(batch), ptr(negatives), n, k, 1.5, margin_or_l3,
                                             temperature, ptr(loss))
                assert status == 0
                np.savez_compressed(os.path.join(out_dir, "kg_kernel_%s_d%d_%s.npz" % (model, dim, oname)),
                                    batch=batch, negatives=negatives, loss=loss, after_head=head, after_tail=tail,
                                    after_relation=relation, after_hm1=moments[0], after_tm1=moments[1],
                                    after_rm1=moments[2], after_hm2=moments[3], after_tm2=moments[4],
                                    after_rm2=moments[5],
                                    hyper=np.array([otype, lr, wd, a, b, eps, 1.5, margin_or_l3, temperature]),
                                    **{"before_" + key: value for key, value in before.items()})
        # predict kernel (one shared entity matrix, as in the solver after training)
        rng = np.random.RandomState(77 + len(model))
        for dim in (32, 512):
            entity = ((rng.rand(90, dim) - 0.5) * (0.3 if dim == 512 else 1.0)).astype(np.float32)
            relation = ((rng.rand(11, dim) - 0.5) * 2.0).astype(np.float32)
            batch = np.stack([rng.randint(0, 11, 200), rng.randint(0, 90, 200), rng.randint(0, 90, 200)],
                             axis=1).astype(np.uint32)  # {relation, tail, head}
            logits = np.zeros(200, dtype=np.float32)
            assert lib.rk_kernel_predict(model.encode(), dim, 90, 11, ptr(entity), ptr(relation), ptr(batch), 200,
                                         6.0, ptr(logits)) == 0
            np.savez_compressed(os.path.join(out_dir, "kg_predict_%s_d%d.npz" % (model, dim)), entity=entity,
                                relation=relation, batch=batch, logits=logits, margin=np.float32(6.0))


def main(out_dir, only=None):
    os.makedirs(out_dir, exist_ok=True)
    toy = os.path.join(GOLDEN, "toy_kg.txt")
    if not os.path.exists(toy):
        make_toy_kg(toy)
    if only == {"toy"}:
        return
    lib = load_harness()
    if not only:
        write_graph(lib, out_dir, toy)
        write_kernel_cases(lib, out_dir)
    write_solver_cases(lib, out_dir, toy, only)
    print("knowledge-graph golden vectors written to", out_dir)


if __name__ == "__main__":
    arguments = [a for a in sys.argv[1:] if a != "--emulated"]
    main(arguments[0] if arguments else os.path.join(ROOT, "gpurun_out", "golden_kg"), set(arguments[1:]) or None)
