"""GPU parity of the device layer (gv_cuda_*) against the oracle.  Tolerances: integer / index
results are bit-exact; fp32 results follow the same algorithm with a different summation order
(float4 lanes + butterfly instead of lane-strided + shfl_down tree), so they agree to ~1e-6;
we assert rtol 2e-4 / atol 2e-6 after up to a few hundred dependent updates."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-4, 2e-6
DIMS = [32, 64, 96, 128, 256, 512]


def make_problem(dim, n, k, rows_v, rows_c, seed, unique=False, moments=0):
    rng = np.random.RandomState(seed)
    vertex = ((rng.rand(rows_v, dim) - 0.5) * 0.6).astype(np.float32)
    context = ((rng.rand(rows_c, dim) - 0.5) * 0.6).astype(np.float32)
    ms = [np.abs(rng.randn(*shape)).astype(np.float32) * 0.01
          for shape in ((rows_v, dim), (rows_c, dim), (rows_v, dim), (rows_c, dim))]
    ms = [m if i // 2 < moments else None for i, m in enumerate(ms)]
    if unique:
        heads = rng.permutation(rows_v)[:n]
        tails = rng.permutation(rows_c)[:n * (k + 1)].reshape(n, k + 1)
    else:
        heads = rng.randint(0, rows_v, n)
        tails = rng.randint(0, rows_c, (n, k + 1))
        if n > 8 and k >= 1:  # force the corner cases: duplicate rows inside a sample, repeated heads
            tails[3, 0] = tails[3, k]
            tails[5, :] = tails[5, 0]
            heads[7] = heads[6]
    batch = np.stack([tails[:, k], heads], axis=1).astype(np.uint32)
    negatives = np.ascontiguousarray(tails[:, :k]).astype(np.uint32)
    return vertex, context, ms, batch, negatives


def oracle_run(dim, vertex, context, ms, batch, negatives, optimizer, negative_weight, lr, batch_size):
    v, c = vertex.copy(), context.copy()
    m = [x.copy() if x is not None else None for x in ms]
    losses = []
    for b, start in enumerate(range(0, batch.shape[0], batch_size)):
        sl = slice(start, start + batch_size)
        losses.append(O.train_batch(dim, v, c, m, batch[sl], negatives[sl], optimizer, negative_weight, lr=float(lr[b])))
    return v, c, m, np.concatenate(losses)


def compare(got, v, c, m, loss):
    np.testing.assert_allclose(got["vertex"], v, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got["context"], c, rtol=RTOL, atol=ATOL)
    for name, expected in zip(["vm1", "cm1", "vm2", "cm2"], m):
        if expected is not None:
            np.testing.assert_allclose(got[name], expected, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got["loss"], loss, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("opt", list(O.OPTIMIZERS))
def test_train_single_warp_with_collisions(dim, opt):
    """one warp = the reference's per-sample order executed sequentially, collisions included"""
    from gpu_util import run_train_block
    optimizer = O.OPTIMIZERS[opt]
    moments = 0 if opt == "SGD" else (2 if opt == "Adam" else 1)
    n, k, batch_size = 257, 3, 100
    vertex, context, ms, batch, negatives = make_problem(dim, n, k, 40, 50, seed=dim + len(opt), moments=moments)
    lr = np.array([optimizer[1], optimizer[1] * 0.5, optimizer[1] * 0.25], dtype=np.float32)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr, batch_size)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr=lr, batch_size=batch_size,
                          num_warps=1)
    compare(got, v, c, m, loss)
    expected_batch = np.array([loss[i:i + batch_size].sum() for i in range(0, n, batch_size)])
    np.testing.assert_allclose(got["batch_loss"], expected_batch, rtol=1e-4)


@pytest.mark.parametrize("dim", [32, 128, 512])
@pytest.mark.parametrize("opt", ["SGD", "Momentum", "Adam"])
@pytest.mark.parametrize("k", [0, 1, 5])
def test_train_full_grid_race_free(dim, opt, k):
    """persistent grid on a batch whose rows are all distinct: no races, so order does not matter"""
    from gpu_util import run_train_block
    optimizer = O.OPTIMIZERS[opt]
    moments = 0 if opt == "SGD" else (2 if opt == "Adam" else 1)
    n = 3000
    vertex, context, ms, batch, negatives = make_problem(dim, n, k, n, n * (k + 1), seed=k + dim, unique=True,
                                                         moments=moments)
    lr = np.full(3, optimizer[1], dtype=np.float32)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives.reshape(n, k), optimizer, 5.0, lr, 1000)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr=lr, batch_size=1000)
    compare(got, v, c, m, loss)


@pytest.fixture
def kernel_flags():
    """sets gv_cuda_set_tunable("kernel_flags") for one test and restores the library default afterwards"""
    from graphvite_b200 import _lib

    def set_flags(value):
        _lib.check(_lib.lib.gv_cuda_set_tunable(b"kernel_flags", int(value)))
    before = _lib.lib.gv_cuda_get_tunable(b"kernel_flags")
    yield set_flags
    set_flags(before)


@pytest.mark.parametrize("flags", [64 | 16, 64 | 1 | 16 | 32, 64 | 1 | 32, 64])
@pytest.mark.parametrize("opt,k,dim,num_warps", [("SGD", 1, 128, 0), ("SGD", 1, 128, 5), ("SGD", 3, 64, 7),
                                                 ("Adam", 2, 128, 3), ("Momentum", 5, 32, 0)])
def test_train_interleaved_mapping_race_free(kernel_flags, flags, opt, k, dim, num_warps):
    """The persistent kernels (flags & 64; without it SGD runs the one-warp-per-sample kernel, which every other SGD test
    of this file covers).  flags & 16: warp w trains the pool entries w, w + G, w + 2G ...;
    flags & 32: write-back stores; flags & 1: every row through L1.  On a batch whose rows are all distinct the
    mapping and the cache policy cannot change the result; 2 999 samples leave a ragged last visit."""
    from gpu_util import run_train_block
    kernel_flags(flags)
    optimizer = O.OPTIMIZERS[opt]
    moments = 0 if opt == "SGD" else (2 if opt == "Adam" else 1)
    n = 2999
    vertex, context, ms, batch, negatives = make_problem(dim, n, k, n, n * (k + 1), seed=k + dim + flags, unique=True,
                                                         moments=moments)
    lr = np.array([optimizer[1], optimizer[1] * 0.5, optimizer[1] * 0.25], dtype=np.float32)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives.reshape(n, k), optimizer, 5.0, lr, 1000)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr=lr, batch_size=1000,
                          num_warps=num_warps)
    compare(got, v, c, m, loss)
    expected_batch = np.array([loss[i:i + 1000].sum() for i in range(0, n, 1000)])
    np.testing.assert_allclose(got["batch_loss"], expected_batch, rtol=1e-4)


@pytest.mark.parametrize("flags", [64, 64 | 16, 64 | 1 | 16 | 32])
def test_train_interleaved_single_warp_is_sequential(kernel_flags, flags):
    """one warp alone visits the pool in order in either mapping: collisions resolve as in the reference's loop"""
    from gpu_util import run_train_block
    kernel_flags(flags)
    optimizer = O.OPTIMIZERS["SGD"]
    dim, n, k = 128, 257, 1
    vertex, context, ms, batch, negatives = make_problem(dim, n, k, 40, 50, seed=11)
    lr = np.array([optimizer[1], optimizer[1] * 0.5, optimizer[1] * 0.25], dtype=np.float32)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr, 100)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr=lr, batch_size=100, num_warps=1)
    compare(got, v, c, m, loss)


@pytest.mark.parametrize("dim,k,num_warps", [(128, 1, 0), (128, 1, 1), (64, 3, 5), (512, 2, 0)])
def test_train_reference_timeline_variant(kernel_flags, dim, k, num_warps):
    """kernel_flags & 512: the one-warp-per-sample kernel with the reference's 128-byte-segment access timeline
    (a measuring instrument for the Hogwild parity study): same arithmetic, so race-free batches and the
    single-warp sequential order match the oracle like every other kernel"""
    from gpu_util import run_train_block
    kernel_flags(512)
    optimizer = O.OPTIMIZERS["SGD"]
    if num_warps == 1:
        n = 257
        vertex, context, ms, batch, negatives = make_problem(dim, n, k, 40, 50, seed=dim + k)
    else:
        n = 2999
        vertex, context, ms, batch, negatives = make_problem(dim, n, k, n, n * (k + 1), seed=dim + k, unique=True)
    lr = np.array([optimizer[1], optimizer[1] * 0.5, optimizer[1] * 0.25], dtype=np.float32)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives.reshape(n, k), optimizer, 5.0, lr, 1000)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, lr=lr, batch_size=1000,
                          num_warps=num_warps)
    compare(got, v, c, m, loss)


@pytest.mark.parametrize("dim,k", [(128, 1), (64, 3), (32, 0), (256, 2), (96, 5)])
def test_train_resident_groups_variant(kernel_flags, dim, k):
    """kernel_flags & 1024: resident blocks take groups of 16 samples by ticket, the next sample's indices are loaded
    and its rows prefetched into L2 during the current one -- same arithmetic, every sample trained exactly once
    (2 999 samples: a ragged last group; three launches in a row re-arm the ticket counter)"""
    from gpu_util import run_train_block
    kernel_flags(1024)
    optimizer = O.OPTIMIZERS["SGD"]
    n = 2999
    for seed in range(3):
        vertex, context, ms, batch, negatives = make_problem(dim, n, k, n, n * (k + 1), seed=dim + k + seed, unique=True)
        lr = np.array([optimizer[1], optimizer[1] * 0.5, optimizer[1] * 0.25], dtype=np.float32)
        v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives.reshape(n, k), optimizer, 5.0, lr, 1000)
        got = run_train_block(dim, vertex, context, ms, batch, negatives.reshape(n, k) if k else negatives, optimizer, 5.0,
                              lr=lr, batch_size=1000)
        compare(got, v, c, m, loss)


def test_train_large_k_shared_memory_opt_in():
    """k = 200 needs > 48 KB of dynamic shared memory for the id staging"""
    from gpu_util import run_train_block
    optimizer = O.OPTIMIZERS["SGD"]
    dim, n, k = 64, 96, 200
    vertex, context, ms, batch, negatives = make_problem(dim, n, k, 30, 400, seed=5)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, [optimizer[1]], n)
    got = run_train_block(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, num_warps=1)
    compare(got, v, c, m, loss)


def test_empty_launch_is_a_no_op():
    from gpu_util import run_train_block
    optimizer = O.OPTIMIZERS["SGD"]
    vertex, context, ms, batch, negatives = make_problem(128, 4, 1, 8, 8, seed=1)
    got = run_train_block(128, vertex, context, ms, batch[:0], negatives[:0], optimizer, 5.0)
    np.testing.assert_array_equal(got["vertex"], vertex)
    np.testing.assert_array_equal(got["context"], context)


@pytest.mark.parametrize("count", [1, 37, 1000, 70001])
def test_sample_negatives_bit_exact(count):
    """gpu::Sample with its double->float narrowing, including rand1 that narrows to 1.0f"""
    from gpu_util import sample_negatives
    rng = np.random.RandomState(count)
    weights = (rng.pareto(1.2, count) + 0.01).astype(np.float32)
    prob, alias = O.alias_build(weights)
    random = 1.0 - rng.rand(2 * 5000)  # (0, 1] like cuRAND
    random[10] = 1.0
    random[12] = 1.0 - 1e-9  # narrows to 1.0f: the reference reads one past the table; both sides clamp
    expected = O.alias_sample(prob, alias, random, gpu_path=True)
    got = sample_negatives(prob, alias.astype(np.uint32), random)
    np.testing.assert_array_equal(got, expected.astype(np.uint32))


def test_fused_negative_sampling_equals_two_step():
    from gpu_util import run_train_block
    rng = np.random.RandomState(9)
    dim, n, k, rows = 128, 999, 2, 64
    optimizer = O.OPTIMIZERS["SGD"]
    vertex, context, ms, batch, _ = make_problem(dim, n, k, rows, rows, seed=3)
    weights = (rng.pareto(1.5, rows) + 0.05).astype(np.float32)
    prob, alias = O.alias_build(weights)
    random = 1.0 - rng.rand(n * k * 2)
    negatives = O.alias_sample(prob, alias, random, gpu_path=True).astype(np.uint32).reshape(n, k)
    v, c, m, loss = oracle_run(dim, vertex, context, ms, batch, negatives, optimizer, 5.0, [optimizer[1]], n)
    got = run_train_block(dim, vertex, context, ms, batch, None, optimizer, 5.0, num_warps=1, random=random,
                          negative_table=(prob, alias.astype(np.uint32)))
    np.testing.assert_array_equal(got["negatives"].reshape(n, k), negatives)
    compare(got, v, c, m, loss)


@pytest.mark.parametrize("dim", DIMS)
def test_predict(dim):
    from gpu_util import predict
    rng = np.random.RandomState(dim)
    vertex = rng.randn(100, dim).astype(np.float32)
    context = rng.randn(120, dim).astype(np.float32)
    batch = np.stack([rng.randint(0, 120, 777), rng.randint(0, 100, 777)], axis=1).astype(np.uint32)
    np.testing.assert_allclose(predict(dim, vertex, context, batch), O.predict_batch(dim, vertex, context, batch),
                               rtol=1e-5, atol=1e-5)


def test_move_rows_round_trip():
    import ctypes
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import dev, stream_pointer
    rng = np.random.RandomState(0)
    matrix = rng.randn(500, 96).astype(np.float32)
    ids = rng.permutation(500)[:123].astype(np.uint32)
    d_matrix, d_ids = dev(matrix), dev(ids)
    d_block = torch.zeros(123, 96, device=gpu_util.DEVICE)
    _lib.check(_lib.lib.gv_cuda_move_rows(d_block.data_ptr(), d_matrix.data_ptr(), d_ids.data_ptr(), 123, 96, 1,
                                          stream_pointer()))
    np.testing.assert_array_equal(d_block.cpu().numpy(), matrix[ids])
    d_back = torch.zeros(500, 96, device=gpu_util.DEVICE)
    _lib.check(_lib.lib.gv_cuda_move_rows(d_back.data_ptr(), d_block.data_ptr(), d_ids.data_ptr(), 123, 96, 0,
                                          stream_pointer()))
    expected = np.zeros_like(matrix)
    expected[ids] = matrix[ids]
    np.testing.assert_array_equal(d_back.cpu().numpy(), expected)


def test_rng_reproduces_curand_stream(golden_dir):
    """gv_rng (our XORWOW kernel) == cuRAND's device generator (golden from the reference harness) ==
    cuRAND's host generator (oracle), for any call split; snapshots rewind it exactly."""
    import ctypes
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import stream_pointer
    lib = _lib.lib
    golden = np.load(golden_dir + "/curand.npz")

    def generate(rng, n):
        out = torch.zeros(n, dtype=torch.float64, device=gpu_util.DEVICE)
        _lib.check(lib.gv_rng_generate(rng, out.data_ptr(), n, stream_pointer()))
        gpu_util.synchronize()
        return out.cpu().numpy()

    rng = lib.gv_rng_create(int(golden["seeds"][0]), stream_pointer())
    assert rng
    small = np.concatenate([generate(rng, int(n)) for n in golden["small_chunks"]])
    np.testing.assert_array_equal(small, golden["small"])  # device cuRAND, same call sizes
    assert lib.gv_rng_position(rng) == len(small)
    lib.gv_rng_destroy(rng)

    seed = int(golden["big_seed"])
    rng = lib.gv_rng_create(seed, stream_pointer())
    first = generate(rng, 5000000)
    snapshot = torch.zeros(lib.gv_rng_state_bytes(), dtype=torch.uint8, device=gpu_util.DEVICE)
    _lib.check(lib.gv_rng_save(rng, snapshot.data_ptr(), stream_pointer()))
    second = generate(rng, 5000000)
    np.testing.assert_array_equal(first[:8192], golden["big_head"])
    np.testing.assert_array_equal(np.r_[first[-2048:], second[:2048]], golden["big_mid"])
    np.testing.assert_array_equal(second[-4096:], golden["big_tail"])
    # odd splits of the same stretch, after rewinding
    _lib.check(lib.gv_rng_restore(rng, snapshot.data_ptr(), stream_pointer()))
    assert lib.gv_rng_position(rng) == 5000000
    pieces = [generate(rng, n) for n in (1, 4095, 4097, 123457, 5000000 - 1 - 4095 - 4097 - 123457)]
    np.testing.assert_array_equal(np.concatenate(pieces), second)
    lib.gv_rng_destroy(rng)
    # host cuRAND (the oracle's generator) for a fresh seed
    expected = O.curand_uniform_double(987654321, [100000])
    rng = lib.gv_rng_create(987654321, stream_pointer())
    np.testing.assert_array_equal(generate(rng, 100000), expected)
    assert expected.min() > 0 and expected.max() <= 1
    lib.gv_rng_destroy(rng)
