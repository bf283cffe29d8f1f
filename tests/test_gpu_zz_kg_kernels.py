"""GPU parity of the knowledge-graph kernels (gv_cuda_kg_train_block / gv_cuda_kg_predict) against the
knowledge-graph oracle (green on B200s: GPUTEST_r01 and `profiles/r02_kg_kernel_tests_after_diet.txt`); the file sorts
behind the node-embedding suites under `pytest -x`.

Tolerances: ids bit-exact; floats follow the same algorithm with a different summation order (slices of
E contiguous floats per thread + butterfly / shared-memory sum instead of lane-strided + shfl_down) and
device libm, so rtol 1e-3 / atol 1e-5 after a few hundred dependent updates."""
import ctypes

import numpy as np
import pytest

import oracle_kg_lib as K
import oracle_lib as O

pytestmark = pytest.mark.gpu

MODEL_IDS = {"TransE": 0, "DistMult": 1, "ComplEx": 2, "SimplE": 3, "RotatE": 4, "QuatE": 5}
TOLERANCE = dict(rtol=1e-3, atol=1e-5)


def run_kg_train(model, dim, head, tail, relation, moments, batch, negatives, optimizer, num_head, rlm, margin_or_l3,
                 temperature, num_group, random=None, negative_count=0, lr=None, batch_size=None):
    """Runs gv_cuda_kg_train_block on copies; tail=None means one shared entity matrix.  moments =
    dict(name -> array) for hm1, tm1, rm1, hm2, tm2, rm2 (tm* ignored when shared)."""
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import dev, host, stream_pointer
    otype, olr, wd, a, b, eps = optimizer
    n = batch.shape[0]
    batch_size = batch_size or max(1, n)
    num_batch = (n + batch_size - 1) // batch_size
    if lr is None:
        lr = np.full(num_batch, olr, dtype=np.float32)
    d = {"head": dev(head), "relation": dev(relation)}
    d["tail"] = d["head"] if tail is None else dev(tail)
    for name in ("hm1", "rm1", "hm2", "rm2"):
        d[name] = dev(moments[name]) if moments and moments.get(name) is not None else None
    for order in ("1", "2"):
        if tail is None:
            d["tm" + order] = d["hm" + order]
        else:
            d["tm" + order] = dev(moments["tm" + order]) if moments and moments.get("tm" + order) is not None else None
    pointer = lambda t: t.data_ptr() if t is not None else None
    m = _lib.KgMatrices(dim, num_head, pointer(d["head"]), pointer(d["tail"]), pointer(d["relation"]), pointer(d["hm1"]),
                        pointer(d["tm1"]), pointer(d["rm1"]), pointer(d["hm2"]), pointer(d["tm2"]), pointer(d["rm2"]))
    d_batch = dev(batch, np.uint32)
    if negatives is not None:
        k = negatives.size // n if n else 0
        d_negatives, d_random, d_out = dev(negatives, np.uint32), None, None
    else:
        k = len(random) // (2 * n)
        d_negatives, d_random = None, dev(np.asarray(random, dtype=np.float64))
        d_out = torch.zeros(n * k, dtype=torch.int32, device=gpu_util.DEVICE)
    d_lr = dev(np.asarray(lr, dtype=np.float32))
    d_loss = torch.zeros(max(1, n), dtype=torch.float32, device=gpu_util.DEVICE)
    d_batch_loss = torch.zeros(num_batch, dtype=torch.float32, device=gpu_util.DEVICE)
    device_optimizer = _lib.DeviceOptimizer(otype, wd, a, b, eps)
    _lib.check(_lib.lib.gv_cuda_kg_train_block(
        ctypes.byref(m), MODEL_IDS[model], d_batch.data_ptr(), n, k, pointer(d_negatives), pointer(d_random),
        negative_count, pointer(d_out), ctypes.byref(device_optimizer), d_lr.data_ptr(), batch_size, rlm, margin_or_l3,
        temperature, d_loss.data_ptr(), d_batch_loss.data_ptr(), num_group, stream_pointer()))
    gpu_util.synchronize()
    result = {name: (tensor.cpu().numpy() if tensor is not None else None) for name, tensor in d.items()}
    result["loss"] = d_loss.cpu().numpy()[:n]
    result["batch_loss"] = d_batch_loss.cpu().numpy()
    if negatives is None:
        result["negatives"] = host(d_out, np.uint32)
    return result


def make_problem(dim, n, k, rows, num_relation, seed, moments, scale=1.0):
    rng = np.random.RandomState(seed)
    entity = ((rng.rand(rows, dim) - 0.5) * scale).astype(np.float32)
    relation = ((rng.rand(num_relation, dim) - 0.5) * 2.0).astype(np.float32)
    ms = None
    if moments:
        ms = {"hm1": np.abs(rng.randn(rows, dim)).astype(np.float32) * 0.01,
              "rm1": np.abs(rng.randn(num_relation, dim)).astype(np.float32) * 0.01,
              "hm2": np.abs(rng.randn(rows, dim)).astype(np.float32) * 0.01 if moments == 2 else None,
              "rm2": np.abs(rng.randn(num_relation, dim)).astype(np.float32) * 0.01 if moments == 2 else None}
    batch = np.stack([rng.randint(0, num_relation, n), rng.randint(0, rows, n), rng.randint(0, rows, n)],
                     axis=1).astype(np.uint32)
    negatives = rng.randint(0, 2 * rows, (n, k)).astype(np.uint32)  # < rows: head, otherwise tail
    if n > 8 and k >= 2:  # corner cases of one shared matrix
        negatives[1, 0] = batch[1, 2]                 # the positive head as a head corruption
        negatives[2, 1] = rows + batch[2, 1]          # the positive tail as a tail corruption
        negatives[3, 0] = batch[3, 1]                 # head corrupted INTO the positive tail row: aliasing target
        negatives[4, 1] = rows + batch[4, 2]          # tail corrupted into the positive head row: aliasing target
        batch[5, 1] = batch[5, 2]                     # a self loop: no row is cached
        negatives[6, 0] = negatives[6, 1] = 7         # the same negative twice
        batch[8] = batch[7]                           # the same triplet twice in a row
    return entity, relation, ms, batch, negatives


def num_moment_of(optimizer):
    return 0 if optimizer[0] == 0 else (2 if optimizer[0] == 4 else 1)


def oracle_shared(model, dim, entity, relation, ms, batch, negatives, optimizer, rlm, margin_or_l3, temperature):
    e, r = entity.copy(), relation.copy()
    m = None
    if ms:
        m = [ms["hm1"].copy(), ms["rm1"].copy(), ms["hm2"].copy() if ms["hm2"] is not None else None,
             ms["rm2"].copy() if ms["rm2"] is not None else None]
    loss = K.train_batch(model, dim, e, r, m, batch, negatives, optimizer, rlm, margin_or_l3, temperature)
    return e, r, m, loss


@pytest.mark.parametrize("dim", [32, 64, 256, 512, 2048])
@pytest.mark.parametrize("opt", list(O.OPTIMIZERS))
@pytest.mark.parametrize("model", K.MODELS)
def test_one_group_is_the_sequential_algorithm(model, opt, dim):
    """one thread group = the reference's per-sample order executed sequentially, on one shared entity
    matrix with every aliasing case (cached rows reused, head == tail targets, self loops, duplicates)"""
    if dim >= 512 and not (model == "RotatE" and opt in ("SGD", "Adam")):
        pytest.skip("large rows are covered for the flagship model only")
    optimizer = O.OPTIMIZERS[opt]
    nm = num_moment_of(optimizer)
    n, k, rows = 40, 5, 30
    entity, relation, ms, batch, negatives = make_problem(dim, n, k, rows, 4, 100 + dim, nm,
                                                          scale=0.2 if dim >= 512 else 1.0)
    margin_or_l3 = 6.0 if model in ("TransE", "RotatE") else 2e-3
    e, r, m, loss = oracle_shared(model, dim, entity, relation, ms, batch, negatives, optimizer, 1.5, margin_or_l3, 1.0)
    got = run_kg_train(model, dim, entity, None, relation, ms, batch, negatives, optimizer, rows, 1.5, margin_or_l3,
                       1.0, num_group=1)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)
    np.testing.assert_allclose(got["head"], e, **TOLERANCE)
    np.testing.assert_allclose(got["relation"], r, **TOLERANCE)
    if nm >= 1:
        np.testing.assert_allclose(got["hm1"], m[0], **TOLERANCE)
        np.testing.assert_allclose(got["rm1"], m[1], **TOLERANCE)
    if nm >= 2:
        np.testing.assert_allclose(got["hm2"], m[2], **TOLERANCE)
        np.testing.assert_allclose(got["rm2"], m[3], **TOLERANCE)
    np.testing.assert_allclose(got["batch_loss"], [loss.sum()], rtol=1e-3)


@pytest.mark.parametrize("temperature", [0.0, 2.0])
@pytest.mark.parametrize("model", K.MODELS)
def test_separate_head_and_tail_blocks(model, temperature):
    """two different partitions: negative ids >= num_head address the tail block"""
    rng = np.random.RandomState(9)
    dim, n, k, num_head, num_tail = 64, 30, 4, 17, 23
    head = ((rng.rand(num_head, dim) - 0.5)).astype(np.float32)
    tail = ((rng.rand(num_tail, dim) - 0.5)).astype(np.float32)
    relation = ((rng.rand(3, dim) - 0.5) * 2).astype(np.float32)
    batch = np.stack([rng.randint(0, 3, n), rng.randint(0, num_tail, n), rng.randint(0, num_head, n)],
                     axis=1).astype(np.uint32)
    negatives = rng.randint(0, num_head + num_tail, (n, k)).astype(np.uint32)
    optimizer = O.OPTIMIZERS["SGD"]
    margin_or_l3 = 5.0 if model in ("TransE", "RotatE") else 1e-3
    h, t, r = head.copy(), tail.copy(), relation.copy()
    loss = K.train_batch(model, dim, h, r, None, batch, negatives, optimizer, 1.0, margin_or_l3, temperature, tail=t,
                         num_head=num_head)
    got = run_kg_train(model, dim, head, tail, relation, None, batch, negatives, optimizer, num_head, 1.0, margin_or_l3,
                       temperature, num_group=1)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)
    np.testing.assert_allclose(got["head"], h, **TOLERANCE)
    np.testing.assert_allclose(got["tail"], t, **TOLERANCE)
    np.testing.assert_allclose(got["relation"], r, **TOLERANCE)


@pytest.mark.parametrize("dim,opt", [(32, "SGD"), (512, "Adam"), (2048, "Adam")])
def test_full_grid_on_a_race_free_batch(dim, opt):
    """every row is named by one sample only, so the Hogwild launch equals the sequential oracle"""
    rng = np.random.RandomState(dim)
    n, k = 48, 3
    rows, num_relation = n * (k + 2), n
    optimizer = O.OPTIMIZERS[opt]
    nm = num_moment_of(optimizer)
    entity, relation, ms, _, _ = make_problem(dim, 1, 1, rows, num_relation, dim + 1, nm, scale=0.2)
    order = rng.permutation(rows)
    batch = np.stack([rng.permutation(num_relation)[:n], order[:n], order[n:2 * n]], axis=1).astype(np.uint32)
    extra = order[2 * n:].reshape(n, k)
    flip = rng.rand(n, k) < 0.5
    negatives = np.where(flip, extra, rows + extra).astype(np.uint32)
    e, r, m, loss = oracle_shared("RotatE", dim, entity, relation, ms, batch, negatives, optimizer, 1.0, 6.0, 2.0)
    got = run_kg_train("RotatE", dim, entity, None, relation, ms, batch, negatives, optimizer, rows, 1.0, 6.0, 2.0,
                       num_group=0, batch_size=16)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)
    np.testing.assert_allclose(got["head"], e, **TOLERANCE)
    np.testing.assert_allclose(got["relation"], r, **TOLERANCE)
    np.testing.assert_allclose(got["batch_loss"], loss.reshape(3, 16).sum(axis=1), rtol=1e-3)


def test_fused_uniform_negative_sampling_is_bit_exact():
    """gpu::Sample over the all-ones table: index = Index(double(float(rand1)) * count), clamped"""
    rng = np.random.RandomState(3)
    dim, n, k, rows = 32, 64, 7, 41
    entity, relation, _, batch, _ = make_problem(dim, n, k, rows, 3, 5, 0)
    random = rng.rand(n * k * 2)
    random[0] = 1.0  # cuRAND doubles lie in (0, 1]: the clamp
    random[2] = np.nextafter(1.0, 0.0)
    count = 2 * rows
    expected = np.minimum((random[0::2].astype(np.float32).astype(np.float64) * count).astype(np.uint32), count - 1)
    got = run_kg_train("TransE", dim, entity, None, relation, None, batch, None, O.OPTIMIZERS["SGD"], rows, 1.0, 4.0,
                       0.0, num_group=1, random=random, negative_count=count)
    np.testing.assert_array_equal(got["negatives"], expected)
    e, r, _, loss = oracle_shared("TransE", dim, entity, relation, None, batch, expected.reshape(n, k),
                                  O.OPTIMIZERS["SGD"], 1.0, 4.0, 0.0)
    np.testing.assert_allclose(got["head"], e, **TOLERANCE)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)


@pytest.mark.parametrize("dim", [32, 96, 256, 1024, 2048])
@pytest.mark.parametrize("model", K.MODELS)
def test_predict(model, dim):
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import dev, stream_pointer
    rng = np.random.RandomState(dim)
    entity = ((rng.rand(50, dim) - 0.5) * (0.2 if dim >= 512 else 1.0)).astype(np.float32)
    relation = ((rng.rand(6, dim) - 0.5) * 2).astype(np.float32)
    batch = np.stack([rng.randint(0, 6, 300), rng.randint(0, 50, 300), rng.randint(0, 50, 300)], axis=1).astype(np.uint32)
    d_entity, d_relation, d_batch = dev(entity), dev(relation), dev(batch, np.uint32)
    logits = torch.zeros(300, dtype=torch.float32, device=gpu_util.DEVICE)
    m = _lib.KgMatrices(dim, 50, d_entity.data_ptr(), d_entity.data_ptr(), d_relation.data_ptr(), None, None, None, None,
                        None, None)
    _lib.check(_lib.lib.gv_cuda_kg_predict(ctypes.byref(m), MODEL_IDS[model], d_batch.data_ptr(), 300, 6.0,
                                           logits.data_ptr(), stream_pointer()))
    gpu_util.synchronize()
    expected = np.array([K.forward(model, entity[h], entity[t], relation[r], 6.0) for r, t, h in batch], dtype=np.float32)
    np.testing.assert_allclose(logits.cpu().numpy(), expected, rtol=1e-4, atol=1e-4)


def test_arguments_are_checked():
    from graphvite_b200 import _lib
    m = _lib.KgMatrices(33, 1, 1, 1, 1, None, None, None, None, None, None)
    optimizer = _lib.DeviceOptimizer(0, 0.0, 0.0, 0.0, 0.0)
    status = _lib.lib.gv_cuda_kg_train_block(ctypes.byref(m), 4, 1, 1, 0, None, None, 0, None, ctypes.byref(optimizer), 1,
                                             1, 1.0, 1.0, 0.0, None, None, 0, None)
    assert status == -1 and "even" in _lib.last_error()
    m = _lib.KgMatrices(32, 1, 1, 1, 1, None, None, None, None, None, None)
    optimizer = _lib.DeviceOptimizer(4, 0.0, 0.9, 0.999, 1e-8)
    status = _lib.lib.gv_cuda_kg_train_block(ctypes.byref(m), 4, 1, 1, 0, None, None, 0, None, ctypes.byref(optimizer), 1,
                                             1, 1.0, 1.0, 0.0, None, None, 0, None)
    assert status == -1 and "moment" in _lib.last_error()


# ---- the product against the REFERENCE's own kernels (tests/golden/kg_kernel_*.npz, recorded by
# oracle/make_golden_kg.py from the unmodified reference) -----------------------------------------------------
import glob
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KERNEL_FILES = sorted(glob.glob(os.path.join(GOLDEN, "kg_kernel_*.npz")))
PREDICT_FILES = sorted(glob.glob(os.path.join(GOLDEN, "kg_predict_*.npz")))


@pytest.mark.skipif(not KERNEL_FILES, reason="no kg_kernel_*.npz fixtures")
@pytest.mark.parametrize("path", KERNEL_FILES, ids=[os.path.basename(p)[10:-4] for p in KERNEL_FILES])
@pytest.mark.parametrize("num_group", [1, 0], ids=["one-group", "full-grid"])
def test_train_kernel_matches_the_reference_kernels(path, num_group):
    """race-free batches (no row is named twice), so the full Hogwild grid must give the same result as one group"""
    g = np.load(path)
    model, dim = os.path.basename(path).split("_")[2], int(os.path.basename(path).split("_")[3][1:])
    otype, lr, wd, a, b, eps, rlm, margin_or_l3, temperature = (float(x) for x in g["hyper"])
    num_moment = 0 if otype == 0 else (2 if otype == 4 else 1)
    moments = {name: (g["before_" + name] if num_moment >= int(name[2]) else None)
               for name in ("hm1", "tm1", "rm1", "hm2", "tm2", "rm2")}
    got = run_kg_train(model, dim, g["before_head"], g["before_tail"], g["before_relation"], moments, g["batch"],
                       g["negatives"], (int(otype), lr, wd, a, b, eps), g["before_head"].shape[0], rlm, margin_or_l3,
                       temperature, num_group)
    tolerance = dict(rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(got["loss"], g["loss"], **tolerance)
    for ours, theirs in (("head", "after_head"), ("tail", "after_tail"), ("relation", "after_relation")):
        np.testing.assert_allclose(got[ours], g[theirs], err_msg=ours, **tolerance)
    for name in ("hm1", "tm1", "rm1", "hm2", "tm2", "rm2"):
        if num_moment >= int(name[2]):
            np.testing.assert_allclose(got[name], g["after_" + name], err_msg=name, **tolerance)


@pytest.mark.skipif(not PREDICT_FILES, reason="no kg_predict_*.npz fixtures")
@pytest.mark.parametrize("path", PREDICT_FILES, ids=[os.path.basename(p)[11:-4] for p in PREDICT_FILES])
def test_predict_kernel_matches_the_reference_kernel(path):
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import dev, stream_pointer
    g = np.load(path)
    model, dim = os.path.basename(path).split("_")[2], g["entity"].shape[1]
    d_entity, d_relation, d_batch = dev(g["entity"]), dev(g["relation"]), dev(g["batch"], np.uint32)
    n = len(g["batch"])
    logits = torch.zeros(n, dtype=torch.float32, device=gpu_util.DEVICE)
    m = _lib.KgMatrices(dim, g["entity"].shape[0], d_entity.data_ptr(), d_entity.data_ptr(), d_relation.data_ptr(),
                        None, None, None, None, None, None)
    _lib.check(_lib.lib.gv_cuda_kg_predict(ctypes.byref(m), MODEL_IDS[model], d_batch.data_ptr(), n, float(g["margin"]),
                                           logits.data_ptr(), stream_pointer()))
    gpu_util.synchronize()
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("shared", [True, False], ids=["one-matrix", "two-matrices"])
@pytest.mark.parametrize("model,opt", [("RotatE", "Adam"), ("TransE", "SGD"), ("QuatE", "Momentum"), ("ComplEx", "SGD")])
def test_next_target_prefetch_never_uses_a_row_the_current_target_writes(model, opt, shared):
    """4 entity rows and 12 negatives per sample: consecutive targets keep naming the same row -- on the same side,
    and (one shared matrix) as head then tail of the same memory -- so the row requested ahead for target s + 1 is
    very often the one target s is about to update.  Sequential semantics must survive (one group vs the oracle)."""
    if not shared and opt != "SGD":
        pytest.skip("the oracle binding takes separate tail matrices without moments only")
    optimizer = O.OPTIMIZERS[opt]
    nm = num_moment_of(optimizer)
    dim, n, k, rows = 64, 30, 12, 4
    entity, relation, ms, batch, negatives = make_problem(dim, n, k, rows, 3, 77, nm)
    negatives[0, :6] = [1, 1, rows + 1, 1, rows + 1, rows + 1]  # same row: head, head, tail, head, tail, tail
    margin_or_l3 = 6.0 if model in ("TransE", "RotatE") else 2e-3
    if shared:
        e, r, m, loss = oracle_shared(model, dim, entity, relation, ms, batch, negatives, optimizer, 1.0, margin_or_l3, 2.0)
        got = run_kg_train(model, dim, entity, None, relation, ms, batch, negatives, optimizer, rows, 1.0,
                           margin_or_l3, 2.0, num_group=1)
        np.testing.assert_allclose(got["head"], e, **TOLERANCE)
    else:
        tail = ((np.random.RandomState(5).rand(rows, dim) - 0.5)).astype(np.float32)
        e, t, r = entity.copy(), tail.copy(), relation.copy()
        loss = K.train_batch(model, dim, e, r, None, batch, negatives, optimizer, 1.0, margin_or_l3, 2.0, tail=t,
                             num_head=rows)
        got = run_kg_train(model, dim, entity, tail, relation, None, batch, negatives, optimizer, rows, 1.0,
                           margin_or_l3, 2.0, num_group=1)
        np.testing.assert_allclose(got["head"], e, **TOLERANCE)
        np.testing.assert_allclose(got["tail"], t, **TOLERANCE)
    np.testing.assert_allclose(got["relation"], r, **TOLERANCE)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)


@pytest.mark.parametrize("k", [1, 3, 4, 7, 9])
@pytest.mark.parametrize("dim", [512, 2048])
def test_multi_warp_groups_with_every_batching_remainder(dim, k):
    """Groups of several warps (dim >= 512) reduce through shared memory; the normaliser pass handles 4 targets per
    barrier, so k = 1, 3, 4, 7, 9 covers a lone partial batch, full batches and both parities of the double-buffered
    scratch at the hand-over from the normaliser pass to the update pass."""
    optimizer = O.OPTIMIZERS["Adam"]
    n, rows = 12, 20
    entity, relation, ms, batch, negatives = make_problem(dim, n, k, rows, 3, 300 + k, 2, scale=0.2)
    e, r, m, loss = oracle_shared("RotatE", dim, entity, relation, ms, batch, negatives, optimizer, 1.0, 6.0, 1.5)
    got = run_kg_train("RotatE", dim, entity, None, relation, ms, batch, negatives, optimizer, rows, 1.0, 6.0, 1.5,
                       num_group=1)
    np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)
    np.testing.assert_allclose(got["head"], e, **TOLERANCE)
    np.testing.assert_allclose(got["relation"], r, **TOLERANCE)
    np.testing.assert_allclose(got["hm2"], m[2], **TOLERANCE)


@pytest.mark.parametrize("model,opt,dim", [("RotatE", "Adam", 256), ("RotatE", "SGD", 256), ("QuatE", "AdaGrad", 256),
                                           ("TransE", "RMSprop", 256), ("RotatE", "Adam", 2048)])
def test_kernel_variants_agree(model, opt, dim):
    """`kg_flags` selects instantiations of the train kernel: bit 0 = IEEE sqrt / division / sincosf() (the default uses
    one MUFU instruction each and gv_sincos), bit 1 = no L2 prefetch of the negative rows, bit 2 = 4 floats per thread
    (a 512-thread group at d = 2048).  All follow the oracle, and the math variants agree with each other far inside
    the tolerance (<= 2 ulp per operation)"""
    from graphvite_b200 import _lib
    optimizer = O.OPTIMIZERS[opt]
    nm = num_moment_of(optimizer)
    n, k, rows = 40, 5, 30
    entity, relation, ms, batch, negatives = make_problem(dim, n, k, rows, 4, 77, nm, scale=0.2 if dim >= 512 else 1.0)
    margin_or_l3 = 6.0 if model in ("TransE", "RotatE") else 2e-3
    e, r, m, loss = oracle_shared(model, dim, entity, relation, ms, batch, negatives, optimizer, 1.5, margin_or_l3, 1.0)
    results = {}
    try:
        for flags in (0, 1, 2, 4):
            _lib.check(_lib.lib.gv_cuda_set_tunable(b"kg_flags", flags))
            assert _lib.lib.gv_cuda_get_tunable(b"kg_flags") == flags
            got = run_kg_train(model, dim, entity, None, relation, ms, batch, negatives, optimizer, rows, 1.5,
                               margin_or_l3, 1.0, num_group=1)
            np.testing.assert_allclose(got["loss"], loss, **TOLERANCE)
            np.testing.assert_allclose(got["head"], e, **TOLERANCE)
            np.testing.assert_allclose(got["relation"], r, **TOLERANCE)
            results[flags] = got
    finally:
        _lib.lib.gv_cuda_set_tunable(b"kg_flags", 0)
    np.testing.assert_allclose(results[0]["head"], results[1]["head"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(results[0]["relation"], results[1]["relation"], rtol=2e-4, atol=2e-6)
    np.testing.assert_array_equal(results[0]["head"], results[2]["head"])  # a prefetch is not a read
