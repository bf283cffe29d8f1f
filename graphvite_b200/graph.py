"""Graphs -- mirrors graphvite.graph.Graph (reference include/bind.h:109-187 over
include/instance/graph.cuh:62-277) and graphvite.graph.KnowledgeGraph (bind.h:237-314 over
include/instance/knowledge_graph.cuh:67-284)."""
import ctypes
import os

from . import _lib
from .base import cfg, dtype

lib = _lib.lib


class _NameMap(object):
    """Read-only name -> id mapping backed by the native graph (bind.h:135,265-266)."""

    def __init__(self, graph, lookup=None, names="id2name"):
        self._graph = graph
        self._lookup = lookup or lib.gv_graph_name2id
        self._names = names

    def __getitem__(self, name):
        index = self._lookup(self._graph._handle, str(name).encode())
        if index < 0:
            raise KeyError(name)
        return index

    def __contains__(self, name):
        return self._lookup(self._graph._handle, str(name).encode()) >= 0

    def get(self, name, default=None):
        index = self._lookup(self._graph._handle, str(name).encode())
        return default if index < 0 else index

    def __len__(self):
        return len(getattr(self._graph, self._names))

    def __iter__(self):
        return iter(getattr(self._graph, self._names))

    def keys(self):
        return getattr(self._graph, self._names)

    def items(self):
        return [(name, i) for i, name in enumerate(getattr(self._graph, self._names))]


class Graph(object):
    """Graph(index_type=dtype.uint32): normal graphs without attributes."""

    def __init__(self, index_type=None):
        index_type = cfg.index_type if index_type is None else index_type
        if index_type != dtype.uint32:
            raise ValueError("Can't find an instantiation of Graph with index_type = %s" % (index_type,))
        self._handle = lib.gv_graph_create()
        self._id2name = None
        self._recipe = None  # how this graph was loaded: lets the workers of a multi-GPU solver load it again

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            lib.gv_graph_destroy(handle)

    # -- load overloads, bind.h:133-159 -------------------------------------------------
    def load(self, *args, **kwargs):
        """load(file_name, as_undirected=True, normalization=False, delimiters=' \\t\\r\\n', comment='#')
        load(edge_list, as_undirected=True, normalization=False)
        load(weighted_edge_list, as_undirected=True, normalization=False)"""
        self._id2name = None
        names = ["file_name", "as_undirected", "normalization", "delimiters", "comment"]
        for alias in ("edge_list", "weighted_edge_list"):
            if alias in kwargs:
                kwargs["file_name"] = kwargs.pop(alias)
        params = dict(zip(names, args))
        for key, value in kwargs.items():
            if key not in names or key in params:
                raise TypeError("load(): incompatible function arguments")
            params[key] = value
        if "file_name" not in params:
            raise TypeError("load(): incompatible function arguments")
        source = params["file_name"]
        as_undirected = bool(params.get("as_undirected", True))
        normalization = bool(params.get("normalization", False))
        if isinstance(source, (str, bytes)):
            delimiters = params.get("delimiters", " \t\r\n")
            comment = params.get("comment", "#")
            path = source if isinstance(source, bytes) else source.encode()
            _lib.check(lib.gv_graph_load_file(self._handle, path, int(as_undirected), int(normalization),
                                              delimiters.encode(), comment.encode()))
            self._recipe = ("file", os.path.abspath(source if isinstance(source, str) else source.decode()),
                            dict(as_undirected=as_undirected, normalization=normalization, delimiters=delimiters,
                                 comment=comment))
            return
        if "delimiters" in params or "comment" in params:
            raise TypeError("load(): incompatible function arguments")
        edges = list(source)
        count = len(edges)
        u_names = (ctypes.c_char_p * count)(*[str(e[0]).encode() for e in edges])
        v_names = (ctypes.c_char_p * count)(*[str(e[1]).encode() for e in edges])
        weights = None
        if count and len(edges[0]) == 3:
            weights = (ctypes.c_float * count)(*[float(e[2]) for e in edges])
        _lib.check(lib.gv_graph_load_edges(self._handle, u_names, v_names, weights, count, int(as_undirected),
                                           int(normalization)))
        self._recipe = ("edges", [tuple(e) for e in edges], dict(as_undirected=as_undirected, normalization=normalization))

    def load_arrays(self, u, v, weights=None, as_undirected=True, normalization=False):
        """load_arrays(u, v, weights=None, as_undirected=True, normalization=False): binary edge list.
        Not in the reference: the graph `load(edge_list=[(str(a), str(b)) ...])` would build (same ids, same edge
        order, same counts) from integer numpy arrays, without materialising the names -- for inputs of Friendster's
        size (1.8e9 edge lines) a Python list of tuples is not an option."""
        import numpy as np
        self._id2name = None
        u = np.ascontiguousarray(u, dtype=np.uint32)
        v = np.ascontiguousarray(v, dtype=np.uint32)
        if u.shape != v.shape or u.ndim != 1:
            raise ValueError("load_arrays(): u and v must be 1-D arrays of the same length")
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        if w is not None and w.shape != u.shape:
            raise ValueError("load_arrays(): weights must match the edge arrays")
        _lib.check(lib.gv_graph_load_id_edges(self._handle, u.ctypes.data, v.ctypes.data,
                                              w.ctypes.data if w is not None else None, len(u), int(as_undirected),
                                              int(normalization)))
        self._recipe = ("arrays", u, v, w, dict(as_undirected=as_undirected, normalization=normalization))

    def save(self, file_name, weighted=True, anonymous=False):
        """save(file_name, weighted=True, anonymous=False): save the graph in edge-list format."""
        _lib.check(lib.gv_graph_save(self._handle, file_name.encode(), int(weighted), int(anonymous)))

    # -- read-only attributes, bind.h:128-135 --------------------------------------------
    @property
    def num_vertex(self):
        return int(lib.gv_graph_num_vertex(self._handle))

    @property
    def num_edge(self):
        return int(lib.gv_graph_num_edge(self._handle))

    @property
    def as_undirected(self):
        return bool(lib.gv_graph_as_undirected(self._handle))

    @property
    def normalization(self):
        return bool(lib.gv_graph_normalization(self._handle))

    @property
    def id2name(self):
        if self._id2name is None or len(self._id2name) != self.num_vertex:
            self._id2name = [lib.gv_graph_id2name(self._handle, i).decode() for i in range(self.num_vertex)]
        return self._id2name

    @property
    def name2id(self):
        return _NameMap(self)

    def __repr__(self):
        buffer = ctypes.create_string_buffer(4096)
        lib.gv_graph_info(self._handle, buffer, len(buffer))
        return buffer.value.decode()


class WordGraph(Graph):
    """WordGraph(index_type=dtype.uint32): normal graphs of word co-occurrences (bind.h:190-234 over
    include/instance/word_graph.cuh:42-166).  A Graph in every other respect: solvers take it like any Graph."""

    def load(self, file_name, window=5, min_count=5, normalization=False, delimiters=" \t\r\n", comment="#"):
        """load(file_name, window=5, min_count=5, normalization=False, delimiters=' \\t\\r\\n', comment='#')"""
        self._id2name = None
        path = file_name if isinstance(file_name, bytes) else str(file_name).encode()
        _lib.check(lib.gv_graph_load_corpus(self._handle, path, int(window), int(min_count), int(bool(normalization)),
                                            delimiters.encode(), comment.encode()))

    def __repr__(self):
        return Graph.__repr__(self).replace("Graph<", "WordGraph<", 1)


class KnowledgeGraph(object):
    """KnowledgeGraph(index_type=dtype.uint32): knowledge graphs (triplets `head relation tail [weight]`)."""

    def __init__(self, index_type=None):
        index_type = cfg.index_type if index_type is None else index_type
        if index_type != dtype.uint32:
            raise ValueError("Can't find an instantiation of KnowledgeGraph with index_type = %s" % (index_type,))
        self._handle = lib.gv_kgraph_create()
        self._names = None
        self._recipe = None  # how this graph was loaded: lets the workers of a multi-GPU solver load it again

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            lib.gv_kgraph_destroy(handle)

    # -- load overloads, bind.h:273-299 ---------------------------------------------------
    def load(self, *args, **kwargs):
        """load(file_name, normalization=False, delimiters=' \\t\\r\\n', comment='#')
        load(triplet_list, normalization=False)
        load(weighted_triplet_list, normalization=False)"""
        self._names = None
        self._recipe = None
        names = ["file_name", "normalization", "delimiters", "comment"]
        for alias in ("triplet_list", "weighted_triplet_list"):
            if alias in kwargs:
                kwargs["file_name"] = kwargs.pop(alias)
        params = dict(zip(names, args))
        for key, value in kwargs.items():
            if key not in names or key in params:
                raise TypeError("load(): incompatible function arguments")
            params[key] = value
        if "file_name" not in params or len(args) > len(names):
            raise TypeError("load(): incompatible function arguments")
        source = params["file_name"]
        normalization = bool(params.get("normalization", False))
        if isinstance(source, (str, bytes)):
            delimiters = params.get("delimiters", " \t\r\n")
            comment = params.get("comment", "#")
            path = source if isinstance(source, bytes) else source.encode()
            _lib.check(lib.gv_kgraph_load_file(self._handle, path, int(normalization), delimiters.encode(),
                                               comment.encode()))
            self._recipe = ("kg_file", os.path.abspath(source if isinstance(source, str) else source.decode()),
                            dict(normalization=normalization, delimiters=delimiters, comment=comment))
            return
        if "delimiters" in params or "comment" in params:
            raise TypeError("load(): incompatible function arguments")
        triplets = list(source)
        count = len(triplets)
        columns = [(ctypes.c_char_p * count)(*[str(t[i]).encode() for t in triplets]) for i in range(3)]
        weights = None
        if count and len(triplets[0]) == 4:
            weights = (ctypes.c_float * count)(*[float(t[3]) for t in triplets])
        _lib.check(lib.gv_kgraph_load_triplets(self._handle, columns[0], columns[1], columns[2], weights, count,
                                               int(normalization)))
        self._recipe = ("kg_triplets", [tuple(t) for t in triplets], dict(normalization=normalization))

    def save(self, file_name, anonymous=False):
        """save(file_name, anonymous=False): save the graph in triplet-list format (head, tail, relation)."""
        _lib.check(lib.gv_kgraph_save(self._handle, file_name.encode(), int(anonymous)))

    # -- read-only attributes, bind.h:261-268 ------------------------------------------------
    @property
    def num_vertex(self):
        return int(lib.gv_kgraph_num_vertex(self._handle))

    @property
    def num_edge(self):
        return int(lib.gv_kgraph_num_edge(self._handle))

    @property
    def num_relation(self):
        return int(lib.gv_kgraph_num_relation(self._handle))

    @property
    def normalization(self):
        return bool(lib.gv_kgraph_normalization(self._handle))

    def _name_lists(self):
        if self._names is None or len(self._names[0]) != self.num_vertex or len(self._names[1]) != self.num_relation:
            self._names = ([lib.gv_kgraph_id2entity(self._handle, i).decode() for i in range(self.num_vertex)],
                           [lib.gv_kgraph_id2relation(self._handle, i).decode() for i in range(self.num_relation)])
        return self._names

    @property
    def id2entity(self):
        return self._name_lists()[0]

    @property
    def id2relation(self):
        return self._name_lists()[1]

    @property
    def entity2id(self):
        return _NameMap(self, lib.gv_kgraph_entity2id, "id2entity")

    @property
    def relation2id(self):
        return _NameMap(self, lib.gv_kgraph_relation2id, "id2relation")

    def __repr__(self):
        buffer = ctypes.create_string_buffer(4096)
        lib.gv_kgraph_info(self._handle, buffer, len(buffer))
        return buffer.value.decode()


__all__ = ["Graph", "WordGraph", "KnowledgeGraph"]
