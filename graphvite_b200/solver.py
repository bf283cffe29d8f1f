"""Solvers -- mirrors graphvite.solver.GraphSolver (reference include/bind.h:383-513 over
include/instance/graph.cuh:587-813 and include/core/solver.h).  One process drives one GPU;
several processes (torchrun) form the reference's multi-GPU solver, see `distributed.py`."""
import ctypes
import os

import numpy as np

from . import _lib
from .base import auto, cfg, dtype
from .graph import Graph, KnowledgeGraph
from .optimizer import as_optimizer

lib = _lib.lib
_DIMS = (32, 64, 96, 128, 256, 512)  # src/graphvite.cu:52-59


class GraphSolver(object):
    """GraphSolver(dim, float_type=dtype.float32, index_type=dtype.uint32, device_ids=[],
    num_sampler_per_worker=auto, gpu_memory_limit=auto)

    Extra keyword arguments (not in the reference): rank, world_size for one-process-per-GPU runs.
    Several `device_ids` in one process (the reference's multi-GPU signature, core/solver.h:184-213) return a front
    end that starts one worker process per listed GPU (`graphvite_b200/multi.py`).
    """

    def __new__(cls, dim, float_type=None, index_type=None, device_ids=(), *args, **kwargs):
        if cls is GraphSolver and len(list(device_ids)) > 1 and kwargs.get("world_size", 1) == 1 and \
                int(os.environ.get("WORLD_SIZE", "1")) == 1:
            from .multi import SpawnedGraphSolver
            return SpawnedGraphSolver(dim, float_type, index_type, device_ids, *args, **kwargs)
        return super(GraphSolver, cls).__new__(cls)

    def __init__(self, dim, float_type=None, index_type=None, device_ids=(), num_sampler_per_worker=auto,
                 gpu_memory_limit=auto, rank=0, world_size=1):
        float_type = cfg.float_type if float_type is None else float_type
        index_type = cfg.index_type if index_type is None else index_type
        if dim not in _DIMS or float_type != dtype.float32 or index_type != dtype.uint32:
            # python/graphvite/helper.py:95-105
            raise ValueError("Can't find an instantiation of GraphSolver with dim = %s, float_type = %s, "
                             "index_type = %s" % (dim, float_type, index_type))
        device_ids = list(device_ids)
        ids = (ctypes.c_int * max(1, len(device_ids)))(*device_ids)
        self._handle = lib.gv_solver_create(int(dim), ids, len(device_ids), int(num_sampler_per_worker),
                                            int(gpu_memory_limit), int(rank), int(world_size))
        if not self._handle:
            raise _lib.GVError(_lib.last_error())
        self.dim = dim
        self._world_size = int(world_size)
        self._graph = None       # the solver borrows the graph (core/solver.h:289): keep it alive
        self._optimizer = None
        self._descriptor = None  # keeps the ctypes schedule callback alive
        self._exchange = None
        if world_size > 1:
            import torch
            from . import distributed
            device = torch.device("cuda", device_ids[0] if device_ids else torch.cuda.current_device())
            distributed.attach(self, device)

    def close(self):
        """Free the solver now.  With world_size > 1 this is a COLLECTIVE call (every rank, same order):
        the ranks first unmap each other's sample pools, synchronise, and only then free their own."""
        handle, self._handle = getattr(self, "_handle", None), None
        if not handle:
            return
        if self._world_size > 1:
            import torch.distributed as dist
            lib.gv_solver_release_peers(handle)
            if dist.is_available() and dist.is_initialized():
                dist.barrier()
        lib.gv_solver_destroy(handle)

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            lib.gv_solver_destroy(handle)

    # -- bind.h:449-453 -------------------------------------------------------------------
    def build(self, graph, optimizer=auto, num_partition=auto, num_negative=1, batch_size=100000,
              episode_size=auto):
        """build(graph, optimizer=auto, num_partition=auto, num_negative=1, batch_size=100000, episode_size=auto)"""
        if not isinstance(graph, Graph):
            raise TypeError("build(): incompatible function arguments (graph must be a Graph)")
        optimizer = as_optimizer(optimizer)
        descriptor = optimizer._descriptor()
        if self._world_size > 1 and self._graph is not None:
            # rebuilding frees the sample pools other ranks have mapped: unmap everywhere first
            import torch.distributed as dist
            _lib.check(lib.gv_solver_release_peers(self._handle))
            dist.barrier()
        _lib.check(lib.gv_solver_build(self._handle, graph._handle, ctypes.byref(descriptor), int(num_partition),
                                       int(num_negative), int(batch_size), int(episode_size)))
        self._graph, self._optimizer, self._descriptor = graph, optimizer, descriptor

    # -- bind.h:466-471 -------------------------------------------------------------------
    def train(self, model="LINE", num_epoch=2000, resume=False, augmentation_step=auto, random_walk_length=40,
              random_walk_batch_size=100, shuffle_base=auto, p=1, q=1, positive_reuse=1,
              negative_sample_exponent=0.75, negative_weight=5, log_frequency=1000):
        """train(model='LINE', num_epoch=2000, resume=False, augmentation_step=auto, random_walk_length=40,
        random_walk_batch_size=100, shuffle_base=auto, p=1, q=1, positive_reuse=1,
        negative_sample_exponent=0.75, negative_weight=5, log_frequency=1000)"""
        _lib.check(lib.gv_solver_train(self._handle, model.encode(), int(num_epoch), int(bool(resume)),
                                       int(augmentation_step), int(random_walk_length), int(random_walk_batch_size),
                                       int(shuffle_base), float(p), float(q), int(positive_reuse),
                                       float(negative_sample_exponent), float(negative_weight), int(log_frequency)))

    # -- bind.h:495 -----------------------------------------------------------------------
    def predict(self, samples):
        """predict(samples): logits for an (?, 2) array of (v, c) vertex ids."""
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        if samples.ndim != 2 or samples.shape[1] != 2:
            raise _lib.GVError("Expect an array with shape (?, 2), but shape (%s) is found" %
                               ", ".join(str(x) for x in samples.shape))
        logits = np.empty(samples.shape[0], dtype=np.float32)
        _lib.check(lib.gv_solver_predict(self._handle, samples.ctypes.data, samples.shape[0], logits.ctypes.data))
        return logits

    def clear(self):
        """clear(): free CPU and GPU memory, except the embeddings on CPU."""
        _lib.check(lib.gv_solver_clear(self._handle))

    # -- numpy views, bind.h:90-106,439-442 -------------------------------------------------
    def _view(self, which):
        rows, dim = ctypes.c_uint64(), ctypes.c_int()
        pointer = lib.gv_solver_embeddings(self._handle, which, ctypes.byref(rows), ctypes.byref(dim))
        if rows.value == 0:
            return np.zeros((0, dim.value), dtype=np.float32)
        return np.ctypeslib.as_array(pointer, shape=(rows.value, dim.value))

    @property
    def vertex_embeddings(self):
        """Vertex node embeddings (2D numpy view of solver-owned memory, mutable)."""
        return self._view(0)

    @property
    def context_embeddings(self):
        """Context node embeddings (2D numpy view of solver-owned memory, mutable)."""
        return self._view(1)

    # -- read-only attributes, bind.h:415-436 ---------------------------------------------------
    def _attributes(self):
        buffer = ctypes.create_string_buffer(4096)
        lib.gv_solver_attributes(self._handle, buffer, len(buffer))
        return dict(line.split("=", 1) for line in buffer.value.decode().splitlines() if "=" in line)

    optimizer = property(lambda self: self._optimizer)
    model = property(lambda self: self._attributes()["model"])
    resume = property(lambda self: bool(int(self._attributes()["resume"])))

    def __getattr__(self, name):
        if name in _INT_ATTRIBUTES:
            return int(self._attributes()[name])
        if name in _FLOAT_ATTRIBUTES:
            return float(self._attributes()[name])
        raise AttributeError("'GraphSolver' object has no attribute '%s'" % name)

    def __repr__(self):
        if not getattr(self, "_handle", None):
            return "<GraphSolver (closed)>"
        buffer = ctypes.create_string_buffer(8192)
        lib.gv_solver_info(self._handle, buffer, len(buffer))
        return buffer.value.decode()

    # -- extras ---------------------------------------------------------------------------------
    @property
    def logged_loss(self):
        """Mean batch losses in the order the reference would LOG them (core/solver.h:1541-1549)."""
        count = lib.gv_solver_logged_loss(self._handle, None, 0)
        out = np.zeros(count, dtype=np.float32)
        lib.gv_solver_logged_loss(self._handle, out.ctypes.data, count)
        return out

    @property
    def stats(self):
        """Counters of the last train(): positives, seconds in train kernels / train loop / samplers, launches."""
        out = np.zeros(5, dtype=np.float64)
        lib.gv_solver_stats(self._handle, out.ctypes.data, 5)
        return dict(zip(["positives", "kernel_seconds", "train_seconds", "sample_seconds", "launches"], out))


class KnowledgeGraphSolver(object):
    """KnowledgeGraphSolver(dim, float_type=dtype.float32, index_type=dtype.uint32, device_ids=[],
    num_sampler_per_worker=auto, gpu_memory_limit=auto)

    Knowledge graph embedding solver (reference include/bind.h:516-639 over
    include/instance/knowledge_graph.cuh:531-677).  Models: TransE, DistMult, ComplEx, SimplE, RotatE, QuatE.
    Extra keyword arguments (not in the reference): rank, world_size for one-process-per-GPU runs.
    Several `device_ids` in one process (the reference's signature, core/solver.h:184-213) return a front end that
    starts one worker process per listed GPU (`graphvite_b200/multi.py`).
    """

    def __new__(cls, dim, float_type=None, index_type=None, device_ids=(), *args, **kwargs):
        if cls is KnowledgeGraphSolver and len(list(device_ids)) > 1 and kwargs.get("world_size", 1) == 1 and \
                int(os.environ.get("WORLD_SIZE", "1")) == 1:
            if dim not in _KG_DIMS:
                raise ValueError("Can't find an instantiation of KnowledgeGraphSolver with dim = %s" % (dim,))
            from .multi import SpawnedKnowledgeGraphSolver
            return SpawnedKnowledgeGraphSolver(dim, float_type, index_type, device_ids, *args, **kwargs)
        return super(KnowledgeGraphSolver, cls).__new__(cls)

    def __init__(self, dim, float_type=None, index_type=None, device_ids=(), num_sampler_per_worker=auto,
                 gpu_memory_limit=auto, rank=0, world_size=1):
        float_type = cfg.float_type if float_type is None else float_type
        index_type = cfg.index_type if index_type is None else index_type
        if dim not in _KG_DIMS or float_type != dtype.float32 or index_type != dtype.uint32:
            raise ValueError("Can't find an instantiation of KnowledgeGraphSolver with dim = %s, float_type = %s, "
                             "index_type = %s" % (dim, float_type, index_type))
        device_ids = list(device_ids)
        ids = (ctypes.c_int * max(1, len(device_ids)))(*device_ids)
        self._handle = lib.gv_kg_solver_create(int(dim), ids, len(device_ids), int(num_sampler_per_worker),
                                               int(gpu_memory_limit), int(rank), int(world_size))
        if not self._handle:
            raise _lib.GVError(_lib.last_error())
        self.dim = dim
        self._world_size = int(world_size)
        self._graph = None
        self._optimizer = None
        self._descriptor = None
        self._exchange = None
        if world_size > 1:
            import torch
            from . import distributed
            device = torch.device("cuda", device_ids[0] if device_ids else torch.cuda.current_device())
            distributed.attach_knowledge_graph(self, device)

    def close(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            lib.gv_kg_solver_destroy(handle)

    __del__ = close

    # -- bind.h:579-594 -------------------------------------------------------------------
    def build(self, graph, optimizer=auto, num_partition=auto, num_negative=64, batch_size=100000,
              episode_size=auto):
        """build(graph, optimizer=auto, num_partition=auto, num_negative=64, batch_size=100000, episode_size=auto)"""
        if not isinstance(graph, KnowledgeGraph):
            raise TypeError("build(): incompatible function arguments (graph must be a KnowledgeGraph)")
        optimizer = as_optimizer(optimizer)
        descriptor = optimizer._descriptor()
        _lib.check(lib.gv_kg_solver_build(self._handle, graph._handle, ctypes.byref(descriptor), int(num_partition),
                                          int(num_negative), int(batch_size), int(episode_size)))
        self._graph, self._optimizer, self._descriptor = graph, optimizer, descriptor

    # -- bind.h:596-619 -------------------------------------------------------------------
    def train(self, model="RotatE", num_epoch=2000, resume=False, relation_lr_multiplier=1, margin=12,
              l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1, adversarial_temperature=2,
              log_frequency=100):
        """train(model='RotatE', num_epoch=2000, resume=False, relation_lr_multiplier=1, margin=12,
        l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1, adversarial_temperature=2,
        log_frequency=100)"""
        _lib.check(lib.gv_kg_solver_train(self._handle, model.encode(), int(num_epoch), int(bool(resume)),
                                          float(relation_lr_multiplier), float(margin), float(l3_regularization),
                                          int(sample_batch_size), int(positive_reuse), float(adversarial_temperature),
                                          int(log_frequency)))

    # -- bind.h:621-629 -------------------------------------------------------------------
    def predict(self, samples):
        """predict(samples): logits for an (?, 3) array of triplets ordered (h, t, r)."""
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        if samples.ndim != 2 or samples.shape[1] != 3:
            raise _lib.GVError("Expect an array with shape (?, 3), but shape (%s) is found" %
                               ", ".join(str(x) for x in samples.shape))
        logits = np.empty(samples.shape[0], dtype=np.float32)
        _lib.check(lib.gv_kg_solver_predict(self._handle, samples.ctypes.data, samples.shape[0], logits.ctypes.data))
        return logits

    def clear(self):
        """clear(): free CPU and GPU memory, except the embeddings on CPU."""
        _lib.check(lib.gv_kg_solver_clear(self._handle))

    # -- numpy views, bind.h:569-572 ----------------------------------------------------------
    def _view(self, which):
        rows, dim = ctypes.c_uint64(), ctypes.c_int()
        pointer = lib.gv_kg_solver_embeddings(self._handle, which, ctypes.byref(rows), ctypes.byref(dim))
        if rows.value == 0:
            return np.zeros((0, dim.value), dtype=np.float32)
        return np.ctypeslib.as_array(pointer, shape=(rows.value, dim.value))

    @property
    def entity_embeddings(self):
        """Entity embeddings (2D numpy view of solver-owned memory, mutable)."""
        return self._view(0)

    @property
    def relation_embeddings(self):
        """Relation embeddings (2D numpy view; RotatE uses the first dim / 2 entries of a row as phases)."""
        return self._view(1)

    def _attributes(self):
        buffer = ctypes.create_string_buffer(4096)
        lib.gv_kg_solver_attributes(self._handle, buffer, len(buffer))
        return dict(line.split("=", 1) for line in buffer.value.decode().splitlines() if "=" in line)

    optimizer = property(lambda self: self._optimizer)
    model = property(lambda self: self._attributes()["model"])
    resume = property(lambda self: bool(int(self._attributes()["resume"])))

    def __getattr__(self, name):
        if name in _KG_INT_ATTRIBUTES:
            return int(self._attributes()[name])
        if name in _KG_FLOAT_ATTRIBUTES:
            return float(self._attributes()[name])
        raise AttributeError("'KnowledgeGraphSolver' object has no attribute '%s'" % name)

    def __repr__(self):
        if not getattr(self, "_handle", None):
            return "<KnowledgeGraphSolver (closed)>"
        buffer = ctypes.create_string_buffer(8192)
        lib.gv_kg_solver_info(self._handle, buffer, len(buffer))
        return buffer.value.decode()

    @property
    def logged_loss(self):
        """Mean batch losses in the order the reference would LOG them (core/solver.h:1541-1549)."""
        count = lib.gv_kg_solver_logged_loss(self._handle, None, 0)
        out = np.zeros(count, dtype=np.float32)
        lib.gv_kg_solver_logged_loss(self._handle, out.ctypes.data, count)
        return out

    @property
    def stats(self):
        out = np.zeros(5, dtype=np.float64)
        lib.gv_kg_solver_stats(self._handle, out.ctypes.data, 5)
        return dict(zip(["positives", "kernel_seconds", "train_seconds", "sample_seconds", "launches"], out))


_KG_DIMS = (32, 64, 96, 128, 256, 512, 1024, 2048)  # src/graphvite.cu:61-70
_KG_INT_ATTRIBUTES = {"num_partition", "num_negative", "sample_batch_size", "num_epoch", "episode_size", "batch_size",
                      "positive_reuse", "log_frequency", "num_worker", "num_sampler", "gpu_memory_limit",
                      "gpu_memory_cost", "num_batch", "batch_id", "pool_id", "partition_size", "rank",
                      "assignment_offset", "shuffle_partition"}
_KG_FLOAT_ATTRIBUTES = {"negative_sample_exponent", "relation_lr_multiplier", "margin", "l3_regularization",
                        "adversarial_temperature"}

_INT_ATTRIBUTES = {"num_partition", "num_negative", "num_epoch", "episode_size", "batch_size", "augmentation_step",
                   "random_walk_length", "random_walk_batch_size", "shuffle_base", "positive_reuse", "log_frequency",
                   "num_worker", "num_sampler", "gpu_memory_limit", "gpu_memory_cost", "num_batch", "batch_id",
                   "pool_id", "partition_size", "rank", "chunk_batches"}
_FLOAT_ATTRIBUTES = {"negative_sample_exponent", "negative_weight", "p", "q"}

__all__ = ["GraphSolver", "KnowledgeGraphSolver"]
