"""Kernel-level parity of the pool fill (gv_cuda_fill_pool, gv_cuda_fill_count / _scatter) against a
sequential restatement of the reference's append loop (instance/graph.cuh:427-447), bit-exact."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sequential_fill(chains, P, L, aug, shuffle_base, pool_size, start, end, fill, pools):
    """the reference loop: walk-major, j, then k; append to block [hp][tp] while its slice has room"""
    num_walk = chains.shape[1]
    last = -1
    for w in range(num_walk):
        for j in range(L):
            for k in range(1, aug + 1):
                if j + k > L:
                    break
                hp, hl = chains[j, w]
                tp, tl = chains[j + k, w]
                b = hp * P + tp
                offset = start + fill[b]
                fill[b] += 1
                if offset < end:
                    shuffled = offset % shuffle_base * (pool_size // shuffle_base) + offset // shuffle_base
                    pools[b][shuffled] = (tl, hl)
                    if offset + 1 == end:
                        last = max(last, w)
    return last


@pytest.mark.parametrize("P,L,aug,shuffle_base", [(1, 5, 2, 2), (2, 6, 3, 3), (3, 4, 4, 1), (4, 7, 2, 2), (8, 5, 5, 5),
                                                  (16, 3, 2, 1), (2, 1, 1, 1)])
def test_fill_pool_matches_sequential_append(P, L, aug, shuffle_base):
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import stream_pointer
    lib = _lib.lib
    rng = np.random.RandomState(P * 100 + L)
    num_block = P * P
    pool_size = 600 * shuffle_base
    start, end = 60 * shuffle_base, 60 * shuffle_base + 400
    expected = [np.full((pool_size, 2), 0xFFFFFFFF, dtype=np.uint32) for _ in range(num_block)]
    fill = np.zeros(num_block, dtype=np.int64)
    d_pools = [torch.full((pool_size, 2), -1, dtype=torch.int32, device=gpu_util.DEVICE) for _ in range(num_block)]
    if P > 1:
        d_pools[num_block - 1] = None  # a block owned by "another rank": counted, never written
    pointers = torch.tensor([p.data_ptr() if p is not None else 0 for p in d_pools], dtype=torch.int64, device=gpu_util.DEVICE)
    d_fill = torch.zeros(num_block, dtype=torch.int64, device=gpu_util.DEVICE)
    d_last = torch.zeros(1, dtype=torch.int64, device=gpu_util.DEVICE)
    params = _lib.FillParams(P, L, aug, shuffle_base, pool_size, start, end)
    first_walk, last_expected = 0, 0
    for call, num_walk in enumerate([257, 1, 1000, 33]):
        chains = np.zeros((L + 1, num_walk, 2), dtype=np.uint32)
        chains[:, :, 0] = rng.randint(0, P, (L + 1, num_walk))
        chains[:, :, 1] = rng.randint(0, 1000, (L + 1, num_walk))
        last = sequential_fill(chains, P, L, aug, shuffle_base, pool_size, start, end, fill, expected)
        if last >= 0:
            last_expected = max(last_expected, first_walk + last)
        d_chains = gpu_util.to_device(torch.from_numpy(chains.view(np.int32)))
        scratch = torch.zeros(lib.gv_cuda_fill_scratch_bytes(num_walk, P) + 16, dtype=torch.uint8, device=gpu_util.DEVICE)
        _lib.check(lib.gv_cuda_fill_pool(ctypes.byref(params), d_chains.data_ptr(), num_walk, first_walk,
                                         pointers.data_ptr(), d_fill.data_ptr(), d_last.data_ptr(),
                                         scratch.data_ptr(), stream_pointer()))
        gpu_util.synchronize()
        first_walk += num_walk
        np.testing.assert_array_equal(d_fill.cpu().numpy(), fill)
    for b in range(num_block - (1 if P > 1 else 0)):
        got = d_pools[b].cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got, expected[b], err_msg="block %d" % b)
    assert int(d_last.item()) == last_expected


@pytest.mark.parametrize("staged,shuffle_base,pool_size", [(False, 3, 900), (True, 3, 900), (True, 1, 900),
                                                          (True, 7, 910), (True, 455, 910)])
def test_partitioned_fill_equals_single_rank(staged, shuffle_base, pool_size):
    """count / exchange-by-hand / scatter over 3 'ranks' gives the pools of one sequential pass.  staged: the blocks a
    rank does not own (tail partition != rank) go through the local staging array and the coalescing forward kernel
    (gv_cuda_fill_scatter_staged), as they do when the pool lives on a peer GPU."""
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import stream_pointer
    lib = _lib.lib
    rng = np.random.RandomState(5)
    P, L, aug = 3, 6, 3
    num_block = P * P
    start, end = 30, 630
    num_walk = 700
    chains = np.zeros((L + 1, num_walk, 2), dtype=np.uint32)
    chains[:, :, 0] = rng.randint(0, P, (L + 1, num_walk))
    chains[:, :, 1] = rng.randint(0, 5000, (L + 1, num_walk))
    expected = [np.full((pool_size, 2), 0xFFFFFFFF, dtype=np.uint32) for _ in range(num_block)]
    fill = np.zeros(num_block, dtype=np.int64)
    sequential_fill(chains, P, L, aug, shuffle_base, pool_size, start, end, fill, expected)

    d_pools = [torch.full((pool_size, 2), -1, dtype=torch.int32, device=gpu_util.DEVICE) for _ in range(num_block)]
    pointers = torch.tensor([p.data_ptr() for p in d_pools], dtype=torch.int64, device=gpu_util.DEVICE)
    params = _lib.FillParams(P, L, aug, shuffle_base, pool_size, start, end)
    bounds = [0, 200, 201, 700]
    totals, state = [], []
    for r in range(3):
        lo, hi = bounds[r], bounds[r + 1]
        part = np.ascontiguousarray(chains[:, lo:hi])
        d_chains = gpu_util.to_device(torch.from_numpy(part.view(np.int32)))
        scratch = torch.zeros(lib.gv_cuda_fill_scratch_bytes(hi - lo, P) + 16, dtype=torch.uint8, device=gpu_util.DEVICE)
        d_totals = torch.zeros(num_block, dtype=torch.int64, device=gpu_util.DEVICE)
        _lib.check(lib.gv_cuda_fill_count(ctypes.byref(params), d_chains.data_ptr(), hi - lo, scratch.data_ptr(),
                                          d_totals.data_ptr(), stream_pointer()))
        gpu_util.synchronize()
        totals.append(d_totals.cpu().numpy())
        state.append((d_chains, scratch, lo, hi))
    np.testing.assert_array_equal(sum(totals), fill)
    d_last = torch.zeros(1, dtype=torch.int64, device=gpu_util.DEVICE)
    for r in range(3):
        d_chains, scratch, lo, hi = state[r]
        bases = gpu_util.to_device(torch.from_numpy(sum(totals[:r], np.zeros(num_block, dtype=np.int64))))
        if not staged:
            _lib.check(lib.gv_cuda_fill_scatter(ctypes.byref(params), d_chains.data_ptr(), hi - lo, lo,
                                                pointers.data_ptr(), bases.data_ptr(), d_last.data_ptr(),
                                                scratch.data_ptr(), stream_pointer()))
            continue
        remote = gpu_util.to_device(torch.tensor([int(b % P != r) for b in range(num_block)], dtype=torch.uint8))
        staging = torch.full((lib.gv_cuda_fill_staging_bytes(hi - lo, L, aug),), 0xEE, dtype=torch.uint8,
                             device=gpu_util.DEVICE)
        offsets = torch.zeros(num_block, dtype=torch.int64, device=gpu_util.DEVICE)
        d_totals = gpu_util.to_device(torch.from_numpy(totals[r]))
        _lib.check(lib.gv_cuda_fill_scatter_staged(ctypes.byref(params), d_chains.data_ptr(), hi - lo, lo,
                                                   pointers.data_ptr(), bases.data_ptr(), d_last.data_ptr(),
                                                   scratch.data_ptr(), remote.data_ptr(), d_totals.data_ptr(),
                                                   staging.data_ptr(), offsets.data_ptr(), stream_pointer()))
        gpu_util.synchronize()
    gpu_util.synchronize()
    for b in range(num_block):
        np.testing.assert_array_equal(d_pools[b].cpu().numpy().view(np.uint32), expected[b], err_msg="block %d" % b)


def test_vertex_alias_tables_built_on_the_device_are_bit_identical():
    """gv_cuda_vertex_tables_build (build_vertex_edge, instance/graph.cuh:645-653) against the oracle's AliasTable
    restatement: weighted vertices (Vose pairing through the rings), uniform vertices (no pairing), isolated
    vertices, one hub."""
    import torch
    import gpu_util
    import oracle_lib as O
    from graphvite_b200 import _lib
    from gpu_util import dev, stream_pointer
    rng = np.random.RandomState(11)
    degrees = rng.randint(0, 40, 300)
    degrees[7] = 3000  # hub
    degrees[[0, 50, 299]] = 0
    offsets = np.concatenate([[0], np.cumsum(degrees)]).astype(np.uint64)
    m = int(offsets[-1])
    weights = (rng.rand(m) * 4 + 0.01).astype(np.float32)
    for v in range(0, 300, 3):  # every third vertex is unweighted
        weights[int(offsets[v]):int(offsets[v + 1])] = 1.5
    weights[int(offsets[8]):int(offsets[9])] = rng.choice([0.5, 2.0], degrees[8]).astype(np.float32)
    d_offsets, d_weights = dev(offsets), dev(weights)
    d_tables = torch.zeros(m, dtype=torch.int64, device=gpu_util.DEVICE)
    d_little = torch.zeros(m, dtype=torch.int32, device=gpu_util.DEVICE)
    d_large = torch.zeros(m, dtype=torch.int32, device=gpu_util.DEVICE)
    graph = _lib.DeviceGraph()
    graph.num_vertex, graph.num_edge, graph.offsets = 300, m, d_offsets.data_ptr()
    _lib.check(_lib.lib.gv_cuda_vertex_tables_build(ctypes.byref(graph), d_weights.data_ptr(), d_tables.data_ptr(),
                                                    d_little.data_ptr(), d_large.data_ptr(), stream_pointer()))
    gpu_util.synchronize()
    tables = d_tables.cpu().numpy().view([("prob", np.float32), ("alias", np.uint32)])
    for v in range(300):
        lo, hi = int(offsets[v]), int(offsets[v + 1])
        if lo == hi:
            continue
        prob, alias = O.alias_build(weights[lo:hi])
        np.testing.assert_array_equal(tables["prob"][lo:hi].view(np.uint32), prob.view(np.uint32), err_msg="vertex %d" % v)
        np.testing.assert_array_equal(tables["alias"][lo:hi], alias.astype(np.uint32), err_msg="vertex %d" % v)


@pytest.mark.parametrize("L,aug,shuffle_base", [(40, 5, 5), (7, 3, 1), (6, 6, 4), (4, 9, 3), (5, 2, 2), (2, 1, 7)])
@pytest.mark.parametrize("per_walk_kernel", [0, 1])
def test_single_block_fill_over_long_slices(L, aug, shuffle_base, per_walk_kernel):
    """P = 1: the tiled, coalesced fill (a CTA owns 32 walks = one contiguous range of slice offsets, written class by
    class of the pseudo shuffle) and the thread-per-walk fill (`fill_per_walk` = 1) against the sequential append, over
    several launches that continue each other and a slice that ends inside the last one."""
    import torch
    import gpu_util
    from graphvite_b200 import _lib
    from gpu_util import stream_pointer
    lib = _lib.lib
    rng = np.random.RandomState(L * 10 + aug)
    per_walk = sum(min(aug, L - j) for j in range(L))
    walks = [70, 1, 333, 64, 5]
    total = sum(walks) * per_walk
    slice_length = total - per_walk * 3 - 1           # ends inside the last launch, not on a walk boundary
    pool_size = (slice_length + 100 + shuffle_base) // shuffle_base * shuffle_base
    start = 37 if 37 + slice_length <= pool_size else 0
    end = start + slice_length
    expected = [np.full((pool_size, 2), 0xFFFFFFFF, dtype=np.uint32)]
    fill = np.zeros(1, dtype=np.int64)
    d_pool = torch.full((pool_size, 2), -1, dtype=torch.int32, device=gpu_util.DEVICE)
    pointers = torch.tensor([d_pool.data_ptr()], dtype=torch.int64, device=gpu_util.DEVICE)
    d_fill = torch.zeros(1, dtype=torch.int64, device=gpu_util.DEVICE)
    d_last = torch.zeros(1, dtype=torch.int64, device=gpu_util.DEVICE)
    params = _lib.FillParams(1, L, aug, shuffle_base, pool_size, start, end)
    assert lib.gv_cuda_set_tunable(b"fill_per_walk", per_walk_kernel) == 0
    try:
        first_walk, last_expected = 0, 0
        for num_walk in walks:
            chains = np.zeros((L + 1, num_walk, 2), dtype=np.uint32)
            chains[:, :, 1] = rng.randint(0, 100000, (L + 1, num_walk))
            last = sequential_fill(chains, 1, L, aug, shuffle_base, pool_size, start, end, fill, expected)
            if last >= 0:
                last_expected = max(last_expected, first_walk + last)
            d_chains = gpu_util.to_device(torch.from_numpy(chains.view(np.int32)))
            _lib.check(lib.gv_cuda_fill_pool(ctypes.byref(params), d_chains.data_ptr(), num_walk, first_walk,
                                             pointers.data_ptr(), d_fill.data_ptr(), d_last.data_ptr(), None,
                                             stream_pointer()))
            gpu_util.synchronize()
            first_walk += num_walk
            np.testing.assert_array_equal(d_fill.cpu().numpy(), fill)
    finally:
        assert lib.gv_cuda_set_tunable(b"fill_per_walk", 0) == 0
    np.testing.assert_array_equal(d_pool.cpu().numpy().view(np.uint32), expected[0])
    assert int(d_last.cpu().numpy()[0]) == last_expected
