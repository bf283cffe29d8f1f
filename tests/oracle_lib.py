"""ctypes access to oracle/liboracle.so -- the CPU restatement of the reference.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this.  The library is built on demand with `make -C oracle liboracle.so` (g++ only).
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
c = ctypes

OPTIMIZERS = {
    # name: (type, lr, weight_decay, a, b, epsilon) -- same table as oracle/make_golden.py
    "SGD": (0, 0.025, 0.005, 0.0, 0.0, 0.0),
    "Momentum": (1, 0.01, 0.001, 0.9, 0.0, 0.0),
    "AdaGrad": (2, 0.05, 0.001, 0.0, 0.0, 1e-10),
    "RMSprop": (3, 0.001, 0.001, 0.99, 0.0, 1e-8),
    "Adam": (4, 0.001, 0.001, 0.9, 0.999, 1e-8),
}


def _build():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    sources = [os.path.join(ORACLE_DIR, name) for name in ("gv_oracle.cpp", "gv_oracle_kg.cpp", "gv_oracle_common.h")]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(source) for source in sources):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


def ptr(array):
    return None if array is None else array.ctypes.data


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = c.CDLL(_build())
    V, I, F, U64 = c.c_void_p, c.c_int, c.c_float, c.c_uint64
    L.og_last_error.restype = c.c_char_p
    L.og_alias_build.argtypes = [V, U64, V, V]
    L.og_alias_sample.argtypes = [V, V, U64, V, U64, I, V]
    L.og_curand_uniform_double.argtypes = [U64, V, I, V]
    L.og_graph_load.restype = V
    L.og_graph_load.argtypes = [c.c_char_p, I, I, c.c_char_p, c.c_char_p]
    L.og_graph_free.argtypes = [V]
    for name in ("og_graph_num_vertex", "og_graph_num_edge", "og_graph_num_directed_edge"):
        getattr(L, name).restype = U64
        getattr(L, name).argtypes = [V]
    L.og_graph_id2name.restype = c.c_char_p
    L.og_graph_id2name.argtypes = [V, U64]
    L.og_graph_vertex_weights.argtypes = [V, V]
    L.og_graph_flat.argtypes = [V, V, V, V, V]
    L.og_partition.argtypes = [V, U64, I, V, V]
    L.og_schedule.argtypes = [I, I, V, I]
    L.og_lr.restype = F
    L.og_lr.argtypes = [I, F, I, I]
    L.og_train_batch.argtypes = [I, V, V, V, V, V, V, V, V, U64, I, I, F, F, F, F, F, F, V]
    L.og_predict_batch.argtypes = [I, V, V, V, U64, V]
    L.og_solver_create.restype = V
    L.og_solver_create.argtypes = [I, I, I]
    L.og_solver_free.argtypes = [V]
    L.og_solver_seeds.argtypes = [V, V, V]
    L.og_solver_build.argtypes = [V, V, I, I, F, F, F, F, F, I, I, I, I]
    L.og_solver_train_begin.argtypes = [V, c.c_char_p, I, I, I, I, I, I, F, F, I, F, F, I]
    L.og_solver_train_episode.argtypes = [V]
    L.og_solver_fill_pool.argtypes = [V]
    L.og_solver_info.argtypes = [V, V]
    L.og_solver_pool.restype = c.POINTER(c.c_uint32)
    L.og_solver_pool.argtypes = [V, I, I, I]
    L.og_solver_locations.argtypes = [V, V, V]
    L.og_solver_embeddings.restype = c.POINTER(c.c_float)
    L.og_solver_embeddings.argtypes = [V, I]
    L.og_solver_moments.restype = c.POINTER(c.c_float)
    L.og_solver_moments.argtypes = [V, I, I]
    L.og_solver_negative_table.restype = c.c_int64
    L.og_solver_negative_table.argtypes = [V, I, V, V]
    L.og_solver_last_negatives.restype = c.POINTER(c.c_uint32)
    L.og_solver_last_negatives.argtypes = [V]
    L.og_solver_last_loss.restype = c.POINTER(c.c_float)
    L.og_solver_last_loss.argtypes = [V]
    L.og_solver_logged_loss.argtypes = [V, V, I]
    L.og_solver_predict.argtypes = [V, V, U64, V]
    L.og_solver_edge_table.argtypes = [V, V, V]
    L.og_solver_vertex_edge_tables.argtypes = [V, V, V]
    _lib = L
    return L


def check(status):
    if status is None or (isinstance(status, int) and status < 0):
        raise RuntimeError(lib().og_last_error().decode())
    return status


def alias_build(weights):
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    prob = np.zeros(len(weights), dtype=np.float32)
    alias = np.zeros(len(weights), dtype=np.uint64)
    check(lib().og_alias_build(ptr(weights), len(weights), ptr(prob), ptr(alias)))
    return prob, alias


def alias_sample(prob, alias, random, gpu_path):
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    alias = np.ascontiguousarray(alias, dtype=np.uint64)
    random = np.ascontiguousarray(random, dtype=np.float64)
    out = np.zeros(len(random) // 2, dtype=np.uint64)
    check(lib().og_alias_sample(ptr(prob), ptr(alias), len(prob), ptr(random), len(out), int(gpu_path), ptr(out)))
    return out


def curand_uniform_double(seed, chunks):
    chunks = np.ascontiguousarray(chunks, dtype=np.uint64)
    out = np.zeros(int(chunks.sum()), dtype=np.float64)
    check(lib().og_curand_uniform_double(int(seed), ptr(chunks), len(chunks), ptr(out)))
    return out


class OracleGraph(object):
    def __init__(self, file_name, as_undirected=True, normalization=False, delimiters=" \t\r\n", comment="#"):
        self.handle = lib().og_graph_load(file_name.encode(), int(as_undirected), int(normalization),
                                          delimiters.encode(), comment.encode())
        if not self.handle:
            raise RuntimeError(lib().og_last_error().decode())
        self.num_vertex = lib().og_graph_num_vertex(self.handle)
        self.num_edge = lib().og_graph_num_edge(self.handle)
        self.num_directed_edge = lib().og_graph_num_directed_edge(self.handle)

    def flat(self):
        m, n = self.num_directed_edge, self.num_vertex
        u, v = np.zeros(m, dtype=np.uint32), np.zeros(m, dtype=np.uint32)
        w, offsets = np.zeros(m, dtype=np.float32), np.zeros(n, dtype=np.uint64)
        lib().og_graph_flat(self.handle, ptr(u), ptr(v), ptr(w), ptr(offsets))
        return u, v, w, offsets

    def vertex_weights(self):
        out = np.zeros(self.num_vertex, dtype=np.float32)
        lib().og_graph_vertex_weights(self.handle, ptr(out))
        return out

    def id2name(self):
        return [lib().og_graph_id2name(self.handle, i).decode() for i in range(self.num_vertex)]

    def __del__(self):
        if getattr(self, "handle", None):
            lib().og_graph_free(self.handle)
            self.handle = None


def partition(weights, num_partition):
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    part_of = np.zeros(len(weights), dtype=np.int32)
    local_of = np.zeros(len(weights), dtype=np.uint32)
    check(lib().og_partition(ptr(weights), len(weights), num_partition, ptr(part_of), ptr(local_of)))
    return part_of, local_of


def schedule(num_partition, num_worker):
    out = np.zeros(4096, dtype=np.int32)
    steps = check(lib().og_schedule(num_partition, num_worker, ptr(out), len(out)))
    width = 1 if num_partition == 1 else num_worker
    return out[:steps * width * 2].reshape(steps, width, 2)


def train_batch(dim, vertex, context, moments, batch, negatives, optimizer, negative_weight, lr=None):
    """Sequential restatement of the train kernels on numpy matrices (updated in place).
    optimizer = (type, lr, weight_decay, a, b, epsilon); moments = [vm1, cm1, vm2, cm2] or Nones."""
    otype, olr, wd, a, b, eps = optimizer
    if lr is None:
        lr = olr
    batch = np.ascontiguousarray(batch, dtype=np.uint32)
    negatives = np.ascontiguousarray(negatives, dtype=np.uint32)
    n = batch.shape[0]
    k = negatives.size // n if n else 0
    loss = np.zeros(n, dtype=np.float32)
    vm1, cm1, vm2, cm2 = moments if moments is not None else (None, None, None, None)
    check(lib().og_train_batch(dim, ptr(vertex), ptr(context), ptr(vm1), ptr(cm1), ptr(vm2), ptr(cm2), ptr(batch),
                               ptr(negatives), n, k, otype, lr, wd, a, b, eps, negative_weight, ptr(loss)))
    return loss


def predict_batch(dim, vertex, context, batch):
    batch = np.ascontiguousarray(batch, dtype=np.uint32)
    logits = np.zeros(batch.shape[0], dtype=np.float32)
    lib().og_predict_batch(dim, ptr(vertex), ptr(context), ptr(batch), batch.shape[0], ptr(logits))
    return logits


class OracleSolver(object):
    """Sequential restatement of GraphSolver (single process, any number of emulated workers)."""

    def __init__(self, graph, dim, num_worker=1, num_sampler_per_worker=1, reset_engine=True):
        L = lib()
        if reset_engine:
            L.og_reset_global_engine()
        self.graph, self.dim = graph, dim
        self.handle = check(L.og_solver_create(dim, num_worker, num_sampler_per_worker))
        self.num_worker = num_worker
        self.num_sampler = num_sampler_per_worker * num_worker

    def __del__(self):
        if getattr(self, "handle", None):
            lib().og_solver_free(self.handle)
            self.handle = None

    def seeds(self):
        sampler = np.zeros(self.num_sampler, dtype=np.uint64)
        worker = np.zeros(self.num_worker, dtype=np.uint64)
        lib().og_solver_seeds(self.handle, ptr(sampler), ptr(worker))
        return sampler, worker

    def build(self, optimizer="SGD", num_partition=0, num_negative=1, batch_size=100000, episode_size=0,
              schedule=1):
        otype, lr, wd, a, b, eps = OPTIMIZERS[optimizer] if isinstance(optimizer, str) else optimizer
        check(lib().og_solver_build(self.handle, self.graph.handle, otype, schedule, lr, wd, a, b, eps,
                                    num_partition, num_negative, batch_size, episode_size))

    def train_begin(self, model="LINE", num_epoch=2000, resume=False, augmentation_step=0, random_walk_length=40,
                    random_walk_batch_size=100, shuffle_base=0, p=1, q=1, positive_reuse=1,
                    negative_sample_exponent=0.75, negative_weight=5, log_frequency=1000):
        check(lib().og_solver_train_begin(self.handle, model.encode(), num_epoch, int(resume), augmentation_step,
                                          random_walk_length, random_walk_batch_size, shuffle_base, p, q,
                                          positive_reuse, negative_sample_exponent, negative_weight, log_frequency))

    def train_episode(self):
        return check(lib().og_solver_train_episode(self.handle)) == 1

    def train(self, **kwargs):
        self.train_begin(**kwargs)
        while self.train_episode():
            pass

    def info(self):
        out = np.zeros(10, dtype=np.int32)
        lib().og_solver_info(self.handle, ptr(out))
        keys = ["num_partition", "episode_size", "batch_size", "augmentation_step", "shuffle_base", "num_batch",
                "batch_id", "pool_id", "num_sampler", "partition_size"]
        return dict(zip(keys, out.tolist()))

    def pool(self, side, head, tail):
        info = self.info()
        n = info["episode_size"] * info["batch_size"]
        pointer = lib().og_solver_pool(self.handle, side, head, tail)
        return np.ctypeslib.as_array(pointer, shape=(n, 2)).copy()

    def locations(self):
        part_of = np.zeros(self.graph.num_vertex, dtype=np.int32)
        local_of = np.zeros(self.graph.num_vertex, dtype=np.uint32)
        lib().og_solver_locations(self.handle, ptr(part_of), ptr(local_of))
        return part_of, local_of

    def embeddings(self, which):
        pointer = lib().og_solver_embeddings(self.handle, which)
        return np.ctypeslib.as_array(pointer, shape=(self.graph.num_vertex, self.dim))

    def negative_table(self, tail_partition):
        n = check(lib().og_solver_negative_table(self.handle, tail_partition, None, None))
        prob, alias = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.uint32)
        lib().og_solver_negative_table(self.handle, tail_partition, ptr(prob), ptr(alias))
        return prob, alias

    def last_negatives(self, batch_size, num_negative):
        pointer = lib().og_solver_last_negatives(self.handle)
        return np.ctypeslib.as_array(pointer, shape=(batch_size * num_negative,)).copy()

    def last_loss(self, batch_size):
        pointer = lib().og_solver_last_loss(self.handle)
        return np.ctypeslib.as_array(pointer, shape=(batch_size,)).copy()

    def logged_loss(self):
        count = lib().og_solver_logged_loss(self.handle, None, 0)
        out = np.zeros(count, dtype=np.float32)
        lib().og_solver_logged_loss(self.handle, ptr(out), count)
        return out

    def edge_table(self):
        m = self.graph.num_directed_edge
        prob, alias = np.zeros(m, dtype=np.float32), np.zeros(m, dtype=np.uint64)
        lib().og_solver_edge_table(self.handle, ptr(prob), ptr(alias))
        return prob, alias

    def vertex_edge_tables(self):
        m = self.graph.num_directed_edge
        prob, alias = np.zeros(m, dtype=np.float32), np.zeros(m, dtype=np.uint32)
        lib().og_solver_vertex_edge_tables(self.handle, ptr(prob), ptr(alias))
        return prob, alias

    def predict(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
        out = np.zeros(pairs.shape[0], dtype=np.float32)
        check(lib().og_solver_predict(self.handle, ptr(pairs), pairs.shape[0], ptr(out)))
        return out
