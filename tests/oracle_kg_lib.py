"""ctypes access to the knowledge-graph half of oracle/liboracle.so (oracle/gv_oracle_kg.cpp).

TEST INFRASTRUCTURE, same rules as oracle_lib.py.  The solver / kernel restatement is parity-unpinned
until the kg_* golden fixtures exist (see the header of gv_oracle_kg.cpp)."""
import ctypes as c

import numpy as np

from oracle_lib import OPTIMIZERS, check, lib as _base_lib, ptr

_ready = False


def lib():
    global _ready
    L = _base_lib()
    if _ready:
        return L
    V, I, F, U64, S = c.c_void_p, c.c_int, c.c_float, c.c_uint64, c.c_char_p
    L.og_kg_graph_load.restype = V
    L.og_kg_graph_load.argtypes = [S, I]
    L.og_kg_graph_from_triplets.restype = V
    L.og_kg_graph_from_triplets.argtypes = [c.POINTER(S), c.POINTER(S), c.POINTER(S), V, U64, I]
    L.og_kg_graph_free.argtypes = [V]
    L.og_kg_graph_sizes.argtypes = [V, V]
    L.og_kg_graph_flat.argtypes = [V, V, V, V, V, V]
    L.og_kg_graph_entity.restype = S
    L.og_kg_graph_entity.argtypes = [V, U64]
    L.og_kg_graph_relation.restype = S
    L.og_kg_graph_relation.argtypes = [V, U64]
    L.og_kg_forward.restype = F
    L.og_kg_forward.argtypes = [S, I, V, V, V, F]
    L.og_kg_train_batch.argtypes = [S, I, c.c_uint32, V, V, V, V, V, V, V, V, V, V, V, U64, I, I, F, F, F, F, F, F, F,
                                    F, V]
    L.og_kg_solver_create.restype = V
    L.og_kg_solver_create.argtypes = [I, I, I]
    L.og_kg_solver_free.argtypes = [V]
    L.og_kg_solver_build.argtypes = [V, V, I, I, F, F, F, F, F, I, I, I, I]
    L.og_kg_solver_set_emulation.argtypes = [V, I, I]
    L.og_kg_solver_set_reference_cache.argtypes = [V, I]
    L.og_kg_solver_train_begin.argtypes = [V, S, I, I, F, F, F, I, I, F, I]
    L.og_kg_solver_train_episode.argtypes = [V]
    L.og_kg_solver_info.argtypes = [V, V]
    L.og_kg_solver_pool.restype = c.POINTER(c.c_uint32)
    L.og_kg_solver_pool.argtypes = [V, I, I, I]
    L.og_kg_solver_locations.argtypes = [V, V, V]
    L.og_kg_solver_matrix.restype = c.POINTER(c.c_float)
    L.og_kg_solver_matrix.argtypes = [V, I, I]
    L.og_kg_solver_last_negatives.restype = c.c_int64
    L.og_kg_solver_last_negatives.argtypes = [V, V]
    L.og_kg_solver_last_loss.argtypes = [V, V]
    L.og_kg_solver_logged_loss.argtypes = [V, V, I]
    L.og_kg_solver_schedule.argtypes = [V, V, I]
    L.og_kg_solver_predict.argtypes = [V, V, U64, V]
    _ready = True
    return L


MODELS = ("TransE", "DistMult", "ComplEx", "SimplE", "RotatE", "QuatE")


class OracleKnowledgeGraph(object):
    def __init__(self, source, normalization=False):
        L = lib()
        if isinstance(source, str):
            self.handle = L.og_kg_graph_load(source.encode(), int(normalization))
        else:
            triplets = list(source)
            n = len(triplets)
            columns = [(c.c_char_p * n)(*[str(t[i]).encode() for t in triplets]) for i in range(3)]
            weights = None
            if n and len(triplets[0]) == 4:
                weights = np.array([t[3] for t in triplets], dtype=np.float32)
            self.handle = L.og_kg_graph_from_triplets(columns[0], columns[1], columns[2], ptr(weights), n,
                                                      int(normalization))
        if not self.handle:
            raise RuntimeError(L.og_last_error().decode())
        sizes = np.zeros(3, dtype=np.uint64)
        L.og_kg_graph_sizes(self.handle, ptr(sizes))
        self.num_vertex, self.num_edge, self.num_relation = (int(x) for x in sizes)

    def flat(self):
        m = self.num_edge
        h, t, r = (np.zeros(m, dtype=np.uint32) for _ in range(3))
        w, vw = np.zeros(m, dtype=np.float32), np.zeros(self.num_vertex, dtype=np.float32)
        lib().og_kg_graph_flat(self.handle, ptr(h), ptr(t), ptr(r), ptr(w), ptr(vw))
        return h, t, r, w, vw

    def id2entity(self):
        return [lib().og_kg_graph_entity(self.handle, i).decode() for i in range(self.num_vertex)]

    def id2relation(self):
        return [lib().og_kg_graph_relation(self.handle, i).decode() for i in range(self.num_relation)]

    def __del__(self):
        if getattr(self, "handle", None):
            lib().og_kg_graph_free(self.handle)
            self.handle = None


def forward(model, head, tail, relation, margin_or_l3):
    head, tail, relation = (np.ascontiguousarray(x, dtype=np.float32) for x in (head, tail, relation))
    return float(lib().og_kg_forward(model.encode(), len(head), ptr(head), ptr(tail), ptr(relation), margin_or_l3))


def train_batch(model, dim, entity, relation, moments, batch, negatives, optimizer, relation_lr_multiplier=1.0,
                margin_or_l3=12.0, adversarial_temperature=2.0, lr=None, tail=None, num_head=None):
    """Sequential restatement of the KG train kernels on numpy matrices (updated in place).
    entity: the head matrix (and the tail matrix too unless `tail` is given); moments =
    [entity_m1, relation_m1, entity_m2, relation_m2] (or None), shared between head and tail like the
    matrices; batch [n][3] = {relation, tail, head}; negatives [n][k] ids into head rows, then tail rows."""
    otype, olr, wd, a, b, eps = optimizer
    batch = np.ascontiguousarray(batch, dtype=np.uint32)
    negatives = np.ascontiguousarray(negatives, dtype=np.uint32)
    n = batch.shape[0]
    k = negatives.size // n if n else 0
    loss = np.zeros(n, dtype=np.float32)
    em1, rm1, em2, rm2 = moments if moments is not None else (None, None, None, None)
    tail_matrix = entity if tail is None else tail
    if tail is not None and moments is not None:
        raise ValueError("separate tail matrices are only supported without moments")
    check(lib().og_kg_train_batch(model.encode(), dim, entity.shape[0] if num_head is None else num_head, ptr(entity),
                                  ptr(tail_matrix), ptr(relation), ptr(em1), ptr(em1), ptr(rm1), ptr(em2), ptr(em2),
                                  ptr(rm2), ptr(batch), ptr(negatives), n, k, otype, olr if lr is None else lr, wd, a,
                                  b, eps, relation_lr_multiplier, margin_or_l3, adversarial_temperature, ptr(loss)))
    return loss


class OracleKGSolver(object):
    def __init__(self, graph, dim, num_worker=1, num_sampler_per_worker=1, reset_engine=True):
        L = lib()
        if reset_engine:
            L.og_reset_global_engine()
        self.graph, self.dim = graph, dim
        self.handle = L.og_kg_solver_create(dim, num_worker, num_sampler_per_worker)

    def set_emulation(self, shuffle_override=-1, synchronous_relation=False):
        """multi-worker emulation switches (see KGSolver in gv_oracle_kg.cpp); call before build()"""
        lib().og_kg_solver_set_emulation(self.handle, shuffle_override, int(synchronous_relation))

    def set_reference_cache(self, on=True):
        """restate the reference's per-worker partition cache, its incoherent "tail hit" included (call before build())"""
        lib().og_kg_solver_set_reference_cache(self.handle, int(on))

    def build(self, optimizer="Adam", num_partition=0, num_negative=64, batch_size=100000, episode_size=0, schedule=1):
        otype, lr, wd, a, b, eps = OPTIMIZERS[optimizer] if isinstance(optimizer, str) else optimizer
        check(lib().og_kg_solver_build(self.handle, self.graph.handle, otype, schedule, lr, wd, a, b, eps,
                                       num_partition, num_negative, batch_size, episode_size))

    def train_begin(self, model="RotatE", num_epoch=2000, resume=False, relation_lr_multiplier=1, margin=12,
                    l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1, adversarial_temperature=2,
                    log_frequency=100):
        check(lib().og_kg_solver_train_begin(self.handle, model.encode(), num_epoch, int(resume),
                                             relation_lr_multiplier, margin, l3_regularization, sample_batch_size,
                                             positive_reuse, adversarial_temperature, log_frequency))

    def train_episode(self):
        return check(lib().og_kg_solver_train_episode(self.handle)) == 1

    def train(self, **kwargs):
        self.train_begin(**kwargs)
        while self.train_episode():
            pass

    def info(self):
        out = np.zeros(10, dtype=np.int32)
        lib().og_kg_solver_info(self.handle, ptr(out))
        keys = ["num_partition", "episode_size", "batch_size", "num_batch", "batch_id", "pool_id", "num_sampler",
                "assignment_offset", "last_negative_count", "shuffle_partition"]
        return dict(zip(keys, out.tolist()))

    def pool(self, side, head, tail):
        info = self.info()
        n = info["episode_size"] * info["batch_size"]
        pointer = lib().og_kg_solver_pool(self.handle, side, head, tail)
        return np.ctypeslib.as_array(pointer, shape=(n, 3)).copy()

    def locations(self):
        part_of = np.zeros(self.graph.num_vertex, dtype=np.int32)
        local_of = np.zeros(self.graph.num_vertex, dtype=np.uint32)
        lib().og_kg_solver_locations(self.handle, ptr(part_of), ptr(local_of))
        return part_of, local_of

    def matrix(self, which, order=0):
        rows = self.graph.num_vertex if which == 0 else self.graph.num_relation
        pointer = lib().og_kg_solver_matrix(self.handle, which, order)
        return np.ctypeslib.as_array(pointer, shape=(rows, self.dim))

    entity_embeddings = property(lambda self: self.matrix(0))
    relation_embeddings = property(lambda self: self.matrix(1))

    def last_negatives(self):
        n = lib().og_kg_solver_last_negatives(self.handle, None)
        out = np.zeros(n, dtype=np.uint32)
        lib().og_kg_solver_last_negatives(self.handle, ptr(out))
        return out

    def last_loss(self):
        n = lib().og_kg_solver_last_loss(self.handle, None)
        out = np.zeros(n, dtype=np.float32)
        lib().og_kg_solver_last_loss(self.handle, ptr(out))
        return out

    def logged_loss(self):
        n = lib().og_kg_solver_logged_loss(self.handle, None, 0)
        out = np.zeros(n, dtype=np.float32)
        lib().og_kg_solver_logged_loss(self.handle, ptr(out), n)
        return out

    def schedule(self, num_worker):
        out = np.zeros(8192, dtype=np.int32)
        steps = check(lib().og_kg_solver_schedule(self.handle, ptr(out), len(out)))
        width = 1 if self.info()["num_partition"] == 1 else num_worker
        return out[:steps * width * 2].reshape(steps, width, 2)

    def predict(self, triplets):
        triplets = np.ascontiguousarray(triplets, dtype=np.uint32)
        out = np.zeros(len(triplets), dtype=np.float32)
        check(lib().og_kg_solver_predict(self.handle, ptr(triplets), len(triplets), ptr(out)))
        return out

    def __del__(self):
        if getattr(self, "handle", None):
            lib().og_kg_solver_free(self.handle)
            self.handle = None
