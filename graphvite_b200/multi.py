"""GraphSolver / KnowledgeGraphSolver(dim, device_ids=[0, 1, ...]) in ONE Python process -- the reference's multi-GPU
signature (include/core/solver.h:184-213, include/bind.h:445-447,575-577: a list of device ids, one worker per id).

The reference drives its GPUs from threads of one process.  Here one process drives one GPU, so
this front end starts one worker process per listed GPU (multiprocessing "spawn"), each of which creates the real
`GraphSolver(rank=r, world_size=W)` (or `KnowledgeGraphSolver`) on its GPU and joins a torch.distributed group on 127.0.0.1 (NCCL; gloo under the
tests' CUDA emulation); the front end forwards build / train / predict / clear and keeps the embeddings in shared
memory, so that `solver.vertex_embeddings` (`entity_embeddings`, `relation_embeddings`) is the same kind of mutable
numpy view the single-GPU solver returns.
What it cannot forward: a graph that was not loaded through `Graph.load` / `Graph.load_arrays` / `KnowledgeGraph.load`
of this process (the workers re-load it from the recorded recipe) and a custom LR schedule that cannot be pickled.
"""
import multiprocessing
import os
import pickle
import socket
import sys
import traceback
from multiprocessing import shared_memory

import numpy as np

_TIMEOUT = float(os.environ.get("GV_MULTI_TIMEOUT", 1800))  # seconds the front end waits for one command


def _optimizer_spec(optimizer):
    from .optimizer import as_optimizer
    optimizer = as_optimizer(optimizer)
    schedule = optimizer.schedule.type
    if schedule == "custom":
        schedule = optimizer.schedule.schedule_function
        try:
            pickle.dumps(schedule)
        except Exception:
            raise ValueError("a custom LR schedule must be picklable (a module-level function) to reach the worker "
                             "processes of a multi-GPU solver")
    fields = {name: getattr(optimizer, name) for name in ("momentum", "alpha", "beta1", "beta2", "epsilon")}
    return type(optimizer).__name__, float(optimizer.lr), float(optimizer.weight_decay), fields, schedule


def _make_optimizer(spec):
    from . import optimizer as O
    name, lr, weight_decay, fields, schedule = spec
    if name == "_Default":
        return O._Default(lr)
    cls = getattr(O, name)
    keywords = {"SGD": (), "Momentum": ("momentum",), "AdaGrad": ("epsilon",), "RMSprop": ("alpha", "epsilon"),
                "Adam": ("beta1", "beta2", "epsilon")}[name]
    return cls(lr, weight_decay, schedule=schedule, **{k: fields[k] for k in keywords})


def _load_graph(recipe):
    from .graph import Graph, KnowledgeGraph
    kind = recipe[0]
    if kind in ("kg_file", "kg_triplets"):
        graph = KnowledgeGraph()
        graph.load(recipe[1], **recipe[2])
        return graph
    graph = Graph()
    if kind == "file":
        graph.load(recipe[1], **recipe[2])
    elif kind == "edges":
        graph.load(recipe[1], **recipe[2])
    elif kind == "arrays":
        graph.load_arrays(recipe[1], recipe[2], recipe[3], **recipe[4])
    else:
        raise ValueError("unknown graph recipe `%s`" % kind)
    return graph


def _embedding_views(solver, knowledge_graph):
    """The solver's own numpy views in the order of the shared blocks."""
    if knowledge_graph:
        return [solver.entity_embeddings, solver.relation_embeddings]
    return [solver.vertex_embeddings, solver.context_embeddings]


def _worker_main(knowledge_graph, rank, world, device_id, port, dim, num_sampler_per_worker, gpu_memory_limit,
                 connection):
    """One worker process = one rank = one GPU.  Serves commands until "close"."""
    emulated = os.environ.get("GV_EMULATE") == "1"
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch
        import torch.distributed as dist
        from . import distributed
        from .solver import GraphSolver, KnowledgeGraphSolver
        if emulated:
            dist.init_process_group("gloo")
        else:
            torch.cuda.set_device(device_id)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device_id))
        solver_class = KnowledgeGraphSolver if knowledge_graph else GraphSolver
        solver = solver_class(dim, device_ids=[0 if emulated else device_id],
                              num_sampler_per_worker=num_sampler_per_worker, gpu_memory_limit=gpu_memory_limit,
                              rank=rank, world_size=world)
        if emulated:  # host buffers over gloo
            (distributed.attach_knowledge_graph if knowledge_graph else distributed.attach)(solver, None)
        graph, views, blocks = None, None, None
        connection.send(("ok", None))
    except BaseException:
        connection.send(("error", traceback.format_exc()))
        return

    def shared_views(names, shapes):
        nonlocal blocks
        blocks = [shared_memory.SharedMemory(name=name) for name in names]
        return [np.ndarray(tuple(shape), dtype=np.float32, buffer=block.buf) for shape, block in zip(shapes, blocks)]

    def push_views():  # the user may have edited the shared views since the last call
        for own, shared in zip(_embedding_views(solver, knowledge_graph), views):
            own[:] = shared

    while True:
        try:
            command = connection.recv()
        except EOFError:
            break
        name, payload = command[0], command[1:]
        try:
            result = None
            if name == "build":
                recipe, optimizer_spec, kwargs, names, shapes = payload
                graph = _load_graph(recipe)
                solver.build(graph, _make_optimizer(optimizer_spec), **kwargs)
                views = shared_views(names, shapes)
                result = solver._attributes()
            elif name == "train":
                kwargs, = payload
                if kwargs.get("resume"):
                    push_views()
                solver.train(**kwargs)
                if rank == 0:  # every rank ends train() with complete matrices; one copy is enough
                    for own, shared in zip(_embedding_views(solver, knowledge_graph), views):
                        shared[:] = own
                result = solver._attributes()
            elif name == "predict":
                samples, = payload
                if rank == 0:
                    push_views()
                    result = solver.predict(samples)
            elif name == "attributes":
                result = solver._attributes()
            elif name == "info":
                result = repr(solver)
            elif name == "clear":
                solver.clear()
            elif name == "close":
                solver.close()  # collective
                if dist.is_initialized():
                    dist.destroy_process_group()
                for block in blocks or []:
                    block.close()
                connection.send(("ok", None))
                return
            else:
                raise ValueError("unknown command `%s`" % name)
            connection.send(("ok", result))
        except BaseException:
            connection.send(("error", traceback.format_exc()))


class _SpawnedSolver(object):
    """Plumbing shared by the two front ends: worker processes, command pipes, shared-memory embedding blocks."""
    _knowledge_graph = False
    _name = "GraphSolver"

    def __init__(self, dim, float_type=None, index_type=None, device_ids=(), num_sampler_per_worker=0,
                 gpu_memory_limit=0, **kwargs):
        from .base import cfg, dtype
        float_type = cfg.float_type if float_type is None else float_type
        index_type = cfg.index_type if index_type is None else index_type
        if float_type != dtype.float32 or index_type != dtype.uint32:
            raise ValueError("Can't find an instantiation of %s with dim = %s, float_type = %s, "
                             "index_type = %s" % (self._name, dim, float_type, index_type))
        if kwargs:
            raise TypeError("unexpected arguments for a multi-GPU solver: %s" % sorted(kwargs))
        self.dim = dim
        self.device_ids = list(device_ids)
        self._world = len(self.device_ids)
        self._graph = self._optimizer = None
        self._blocks, self._views = [], None
        self._attribute_cache = {}
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        context = multiprocessing.get_context("spawn")
        self._workers, self._pipes = [], []
        for rank, device_id in enumerate(self.device_ids):
            parent, child = context.Pipe()
            process = context.Process(target=_worker_main, daemon=True,
                                      args=(self._knowledge_graph, rank, self._world, device_id, port, dim,
                                            int(num_sampler_per_worker), int(gpu_memory_limit), child))
            process.start()
            child.close()
            self._workers.append(process)
            self._pipes.append(parent)
        self._collect("start")

    # ---- plumbing ---------------------------------------------------------------------------
    def _collect(self, what):
        results = []
        for rank, pipe in enumerate(self._pipes):
            if not pipe.poll(_TIMEOUT):
                self._abort()
                raise RuntimeError("worker %d did not answer `%s` within %d s" % (rank, what, _TIMEOUT))
            try:
                status, payload = pipe.recv()
            except EOFError:
                self._abort()
                raise RuntimeError("worker %d died during `%s`" % (rank, what))
            if status != "ok":
                self._abort()
                raise RuntimeError("worker %d failed in `%s`:\n%s" % (rank, what, payload))
            results.append(payload)
        return results

    def _call(self, name, *payload):
        if not self._pipes:
            raise RuntimeError("the solver was closed")
        for pipe in self._pipes:
            pipe.send((name,) + payload)
        return self._collect(name)

    def _abort(self):
        for process in self._workers:
            if process.is_alive():
                process.terminate()
        self._pipes, self._workers = [], []
        self._release_shared()

    def _release_shared(self):
        self._views = None
        for block in self._blocks:
            try:
                block.close()
                block.unlink()
            except (FileNotFoundError, BufferError):
                pass
        self._blocks = []

    def _build(self, graph, optimizer, kwargs, shapes):
        recipe = getattr(graph, "_recipe", None)
        if recipe is None:
            raise ValueError("a multi-GPU solver re-loads the graph in its worker processes: load it with the load() "
                             "(or Graph.load_arrays()) of this process first")
        self._release_shared()
        self._blocks = [shared_memory.SharedMemory(create=True, size=max(1, int(np.prod(shape)) * 4))
                        for shape in shapes]
        self._views = [np.ndarray(tuple(shape), dtype=np.float32, buffer=block.buf)
                       for shape, block in zip(shapes, self._blocks)]
        for view in self._views:
            view[:] = 0
        answers = self._call("build", recipe, _optimizer_spec(optimizer), kwargs, [b.name for b in self._blocks],
                             [tuple(shape) for shape in shapes])
        self._attribute_cache = answers[0]
        self._graph = graph
        from .optimizer import as_optimizer
        self._optimizer = as_optimizer(optimizer)

    def clear(self):
        self._call("clear")

    def close(self):
        if self._pipes:
            try:
                self._call("close")
            finally:
                for process in self._workers:
                    process.join(timeout=30)
                self._pipes, self._workers = [], []
        self._release_shared()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, index, columns):
        return self._views[index] if self._views else np.zeros((0, columns), dtype=np.float32)

    optimizer = property(lambda self: self._optimizer)
    model = property(lambda self: self._attributes()["model"])
    resume = property(lambda self: bool(int(self._attributes()["resume"])))

    def _attributes(self):
        return self._attribute_cache or self._call("attributes")[0]

    def __repr__(self):
        return self._call("info")[0] if self._pipes else "<%s (closed)>" % self._name


class SpawnedGraphSolver(_SpawnedSolver):
    """What `GraphSolver(dim, device_ids=[...several...])` returns: the same methods and attributes, executed by one
    worker process per listed GPU."""

    # ---- the reference's surface (bind.h:383-513) ---------------------------------------------
    def build(self, graph, optimizer=0, num_partition=0, num_negative=1, batch_size=100000, episode_size=0):
        kwargs = dict(num_partition=int(num_partition), num_negative=int(num_negative), batch_size=int(batch_size),
                      episode_size=int(episode_size))
        self._build(graph, optimizer, kwargs, [(graph.num_vertex, self.dim)] * 2)

    def train(self, model="LINE", num_epoch=2000, resume=False, augmentation_step=0, random_walk_length=40,
              random_walk_batch_size=100, shuffle_base=0, p=1, q=1, positive_reuse=1, negative_sample_exponent=0.75,
              negative_weight=5, log_frequency=1000):
        kwargs = dict(model=model, num_epoch=int(num_epoch), resume=bool(resume),
                      augmentation_step=int(augmentation_step), random_walk_length=int(random_walk_length),
                      random_walk_batch_size=int(random_walk_batch_size), shuffle_base=int(shuffle_base), p=float(p),
                      q=float(q), positive_reuse=int(positive_reuse),
                      negative_sample_exponent=float(negative_sample_exponent), negative_weight=float(negative_weight),
                      log_frequency=int(log_frequency))
        self._attribute_cache = self._call("train", kwargs)[0]

    def predict(self, samples):
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        if samples.ndim != 2 or samples.shape[1] != 2:
            raise ValueError("Expect an array with shape (?, 2), but shape (%s) is found" %
                             ", ".join(str(x) for x in samples.shape))
        return self._call("predict", samples)[0]

    vertex_embeddings = property(lambda self: self._view(0, self.dim))
    context_embeddings = property(lambda self: self._view(1, self.dim))

    def __getattr__(self, name):
        from .solver import _FLOAT_ATTRIBUTES, _INT_ATTRIBUTES
        if name in _INT_ATTRIBUTES:
            return int(self._attributes()[name])
        if name in _FLOAT_ATTRIBUTES:
            return float(self._attributes()[name])
        raise AttributeError("'GraphSolver' object has no attribute '%s'" % name)


class SpawnedKnowledgeGraphSolver(_SpawnedSolver):
    """What `KnowledgeGraphSolver(dim, device_ids=[...several...])` returns (bind.h:516-639): one worker process per
    listed GPU; entity blocks move between them by NCCL send / recv, relation deltas by one all-reduce per step."""
    _knowledge_graph = True
    _name = "KnowledgeGraphSolver"

    def build(self, graph, optimizer=0, num_partition=0, num_negative=64, batch_size=100000, episode_size=0):
        kwargs = dict(num_partition=int(num_partition), num_negative=int(num_negative), batch_size=int(batch_size),
                      episode_size=int(episode_size))
        # relation rows have the entity dimension (RotatE uses the first dim / 2 entries as phases)
        self._build(graph, optimizer, kwargs, [(graph.num_vertex, self.dim), (graph.num_relation, self.dim)])

    def train(self, model="RotatE", num_epoch=2000, resume=False, relation_lr_multiplier=1, margin=12,
              l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1, adversarial_temperature=2,
              log_frequency=100):
        kwargs = dict(model=model, num_epoch=int(num_epoch), resume=bool(resume),
                      relation_lr_multiplier=float(relation_lr_multiplier), margin=float(margin),
                      l3_regularization=float(l3_regularization), sample_batch_size=int(sample_batch_size),
                      positive_reuse=int(positive_reuse), adversarial_temperature=float(adversarial_temperature),
                      log_frequency=int(log_frequency))
        self._attribute_cache = self._call("train", kwargs)[0]

    def predict(self, samples):
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        if samples.ndim != 2 or samples.shape[1] != 3:
            raise ValueError("Expect an array with shape (?, 3), but shape (%s) is found" %
                             ", ".join(str(x) for x in samples.shape))
        return self._call("predict", samples)[0]

    entity_embeddings = property(lambda self: self._view(0, self.dim))
    relation_embeddings = property(lambda self: self._view(1, self.dim))

    def __getattr__(self, name):
        from .solver import _KG_FLOAT_ATTRIBUTES, _KG_INT_ATTRIBUTES
        if name in _KG_INT_ATTRIBUTES:
            return int(self._attributes()[name])
        if name in _KG_FLOAT_ATTRIBUTES:
            return float(self._attributes()[name])
        raise AttributeError("'KnowledgeGraphSolver' object has no attribute '%s'" % name)
