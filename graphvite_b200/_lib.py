"""ctypes binding of libgv_b200.so (the C ABI declared in include/gv_b200.h).

The library is loaded eagerly and loudly: there is no CPU or PyTorch fallback for any
entry point of this package.  If the shared object is missing, build it with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C graphvite_b200/csrc``).
"""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_here, "libgv_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "graphvite_b200: %s not found. The CUDA extension is mandatory (no CPU fallback); "
        "build it with `make -C graphvite_b200/csrc` or `__graft_entry__.build()`." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

c_void_p, c_int, c_float, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p
c_uint32, c_uint64, c_int64, c_size_t, c_double = (ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64,
                                                   ctypes.c_size_t, ctypes.c_double)
P = ctypes.POINTER

SCHEDULE_FN = ctypes.CFUNCTYPE(c_float, c_int, c_int, c_void_p)
EXCHANGE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p, c_int, c_uint64, c_void_p, c_void_p)
HOST_ALLGATHER_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_uint64, c_void_p)
ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_uint64, c_void_p, c_void_p)


class OptimizerDesc(ctypes.Structure):
    """gv_optimizer_t"""
    _fields_ = [("type", c_int), ("lr", c_float), ("weight_decay", c_float), ("a", c_float), ("b", c_float),
                ("epsilon", c_float), ("schedule", c_int), ("schedule_fn", SCHEDULE_FN), ("schedule_ctx", c_void_p)]


class DeviceOptimizer(ctypes.Structure):
    """gv_device_optimizer_t"""
    _fields_ = [("type", c_int), ("weight_decay", c_float), ("a", c_float), ("b", c_float), ("epsilon", c_float)]


class Matrices(ctypes.Structure):
    """gv_matrices_t (device pointers)"""
    _fields_ = [("dim", c_int), ("vertex", c_void_p), ("context", c_void_p), ("vertex_m1", c_void_p),
                ("context_m1", c_void_p), ("vertex_m2", c_void_p), ("context_m2", c_void_p)]


class DeviceGraph(ctypes.Structure):
    """gv_device_graph_t (device pointers)"""
    _fields_ = [("num_vertex", c_uint32), ("num_edge", c_uint64), ("offsets", c_void_p), ("edge_u", c_void_p),
                ("edge_v", c_void_p), ("edge_prob", c_void_p), ("edge_alias", c_void_p),
                ("vertex_tables", c_void_p), ("locations", c_void_p)]


class KgMatrices(ctypes.Structure):
    """gv_kg_matrices_t"""
    _fields_ = [("dim", c_int), ("num_head", c_uint32), ("head", c_void_p), ("tail", c_void_p),
                ("relation", c_void_p), ("head_m1", c_void_p), ("tail_m1", c_void_p), ("relation_m1", c_void_p),
                ("head_m2", c_void_p), ("tail_m2", c_void_p), ("relation_m2", c_void_p)]


class DeviceKGraph(ctypes.Structure):
    """gv_device_kgraph_t (device pointers)"""
    _fields_ = [("num_edge", c_uint64), ("edge_h", c_void_p), ("edge_t", c_void_p), ("edge_r", c_void_p),
                ("edge_prob", c_void_p), ("edge_alias", c_void_p), ("locations", c_void_p)]


class TableShards(ctypes.Structure):
    """gv_table_shards_t"""
    _fields_ = [("num_shard", c_int), ("shard", c_void_p * 16), ("first_entry", ctypes.c_ulonglong * 17)]


class FillParams(ctypes.Structure):
    """gv_fill_params_t"""
    _fields_ = [("num_partition", c_int), ("walk_length", c_int), ("augmentation_step", c_int),
                ("shuffle_base", c_int), ("pool_size", c_uint64), ("start", c_uint64), ("end", c_uint64),
                ("attributes", c_void_p)]


# every symbol include/gv_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gv_last_error": (c_char_p, []),
    "gv_version": (c_char_p, []),
    # device layer
    "gv_cuda_train_block": (c_int, [P(Matrices), c_void_p, c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_uint32,
                                    c_void_p, P(DeviceOptimizer), c_void_p, c_uint32, c_float, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    "gv_rng_create": (c_void_p, [ctypes.c_ulonglong, c_void_p]),
    "gv_rng_destroy": (None, [c_void_p]),
    "gv_rng_generate": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "gv_rng_position": (c_uint64, [c_void_p]),
    "gv_rng_state_bytes": (c_size_t, []),
    "gv_rng_save": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gv_rng_restore": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gv_cuda_set_tunable": (c_int, [c_char_p, ctypes.c_long]),
    "gv_cuda_get_tunable": (ctypes.c_long, [c_char_p]),
    "gv_cuda_sample_negatives": (c_int, [c_void_p, c_uint32, c_void_p, c_uint64, c_void_p, c_void_p]),
    "gv_cuda_predict": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    "gv_cuda_random_walk": (c_int, [P(DeviceGraph), c_void_p, c_uint32, c_int, c_uint64, c_uint32, c_uint64,
                                    c_void_p, c_void_p]),
    "gv_cuda_vertex_tables_build": (c_int, [P(DeviceGraph), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_node2vec_build": (c_int, [P(DeviceGraph), c_void_p, c_void_p, c_void_p, c_uint64, c_uint32, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_biased_walk": (c_int, [P(DeviceGraph), c_void_p, c_void_p, c_void_p, c_uint32, c_int, c_uint64, c_uint32,
                                    c_uint64, c_void_p, c_void_p]),
    "gv_cuda_biased_walk_sharded": (c_int, [P(DeviceGraph), P(TableShards), c_void_p, c_void_p, c_uint32, c_int, c_uint64,
                                            c_uint32, c_uint64, c_void_p, c_void_p]),
    "gv_cuda_fill_scratch_bytes": (c_size_t, [c_uint32, c_int]),
    "gv_cuda_fill_pool": (c_int, [P(FillParams), c_void_p, c_uint32, c_uint64, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "gv_cuda_kg_train_block": (c_int, [P(KgMatrices), c_int, c_void_p, c_uint64, c_int, c_void_p, c_void_p, c_uint32,
                                       c_void_p, P(DeviceOptimizer), c_void_p, c_uint32, c_float, c_float, c_float,
                                       c_void_p, c_void_p, c_int, c_void_p]),
    "gv_cuda_kg_predict": (c_int, [P(KgMatrices), c_int, c_void_p, c_uint64, c_float, c_void_p, c_void_p]),
    "gv_cuda_kg_draw": (c_int, [P(DeviceKGraph), c_void_p, c_uint32, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_kg_relation_delta": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "gv_cuda_kg_relation_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "gv_cuda_move_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p]),
    "gv_cuda_fill_float": (c_int, [c_void_p, c_uint64, ctypes.c_float, c_void_p]),
    "gv_cuda_fill_identity": (c_int, [c_void_p, c_uint64, c_void_p]),
    "gv_cuda_expand_sources": (c_int, [c_void_p, ctypes.c_uint32, c_void_p, c_void_p]),
    "gv_cuda_fill_count": (c_int, [P(FillParams), c_void_p, c_uint32, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_fill_scatter": (c_int, [P(FillParams), c_void_p, c_uint32, c_uint64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "gv_cuda_fill_advance": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_fill_staging_bytes": (c_size_t, [c_uint32, c_int, c_int]),
    "gv_cuda_fill_scatter_staged": (c_int, [P(FillParams), c_void_p, c_uint32, c_uint64, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_cuda_peer_control_bytes": (c_size_t, [c_int, c_int]),
    "gv_cuda_peer_exchange": (c_int, [c_int, c_int, c_int, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    # graph
    "gv_graph_create": (c_void_p, []),
    "gv_graph_destroy": (None, [c_void_p]),
    "gv_graph_load_file": (c_int, [c_void_p, c_char_p, c_int, c_int, c_char_p, c_char_p]),
    "gv_graph_load_corpus": (c_int, [c_void_p, c_char_p, c_int, c_int, c_int, c_char_p, c_char_p]),
    "gv_graph_load_edges": (c_int, [c_void_p, P(c_char_p), P(c_char_p), P(c_float), c_uint64, c_int, c_int]),
    "gv_graph_save": (c_int, [c_void_p, c_char_p, c_int, c_int]),
    "gv_graph_num_vertex": (c_uint64, [c_void_p]),
    "gv_graph_num_edge": (c_uint64, [c_void_p]),
    "gv_graph_as_undirected": (c_int, [c_void_p]),
    "gv_graph_normalization": (c_int, [c_void_p]),
    "gv_graph_id2name": (c_char_p, [c_void_p, c_uint64]),
    "gv_graph_load_id_edges": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_int, c_int]),
    "gv_graph_name2id": (c_int64, [c_void_p, c_char_p]),
    "gv_graph_flatten": (c_uint64, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_graph_info": (c_int, [c_void_p, c_char_p, c_size_t]),
    "gv_alias_build": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p]),
    # knowledge graph
    "gv_kgraph_create": (c_void_p, []),
    "gv_kgraph_destroy": (None, [c_void_p]),
    "gv_kgraph_load_file": (c_int, [c_void_p, c_char_p, c_int, c_char_p, c_char_p]),
    "gv_kgraph_load_triplets": (c_int, [c_void_p, P(c_char_p), P(c_char_p), P(c_char_p), P(c_float), c_uint64,
                                        c_int]),
    "gv_kgraph_save": (c_int, [c_void_p, c_char_p, c_int]),
    "gv_kgraph_num_vertex": (c_uint64, [c_void_p]),
    "gv_kgraph_num_edge": (c_uint64, [c_void_p]),
    "gv_kgraph_num_relation": (c_uint64, [c_void_p]),
    "gv_kgraph_normalization": (c_int, [c_void_p]),
    "gv_kgraph_id2entity": (c_char_p, [c_void_p, c_uint64]),
    "gv_kgraph_id2relation": (c_char_p, [c_void_p, c_uint64]),
    "gv_kgraph_entity2id": (c_int64, [c_void_p, c_char_p]),
    "gv_kgraph_relation2id": (c_int64, [c_void_p, c_char_p]),
    "gv_kgraph_flatten": (c_uint64, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gv_kgraph_info": (c_int, [c_void_p, c_char_p, c_size_t]),
    # solver
    "gv_solver_create": (c_void_p, [c_int, P(c_int), c_int, c_int, c_uint64, c_int, c_int]),
    "gv_solver_destroy": (None, [c_void_p]),
    "gv_solver_set_exchange": (c_int, [c_void_p, EXCHANGE_FN, c_void_p]),
    "gv_solver_set_host_allgather": (c_int, [c_void_p, HOST_ALLGATHER_FN, c_void_p]),
    "gv_solver_release_peers": (c_int, [c_void_p]),
    "gv_solver_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "gv_solver_build": (c_int, [c_void_p, c_void_p, P(OptimizerDesc), c_int, c_int, c_int, c_int]),
    "gv_solver_train": (c_int, [c_void_p, c_char_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                c_int, c_float, c_float, c_int]),
    "gv_solver_predict": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "gv_solver_clear": (c_int, [c_void_p]),
    "gv_solver_embeddings": (P(c_float), [c_void_p, c_int, P(c_uint64), P(c_int)]),
    "gv_solver_info": (c_int, [c_void_p, c_char_p, c_size_t]),
    "gv_solver_attributes": (c_int, [c_void_p, c_char_p, c_size_t]),
    "gv_solver_logged_loss": (c_int, [c_void_p, c_void_p, c_int]),
    "gv_solver_stats": (c_int, [c_void_p, c_void_p, c_int]),
    # knowledge-graph solver
    "gv_kg_solver_create": (c_void_p, [c_int, P(c_int), c_int, c_int, c_uint64, c_int, c_int]),
    "gv_kg_solver_destroy": (None, [c_void_p]),
    "gv_kg_solver_set_exchange": (c_int, [c_void_p, EXCHANGE_FN, c_void_p]),
    "gv_kg_solver_set_allreduce": (c_int, [c_void_p, ALLREDUCE_FN, c_void_p]),
    "gv_kg_solver_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "gv_kg_solver_build": (c_int, [c_void_p, c_void_p, P(OptimizerDesc), c_int, c_int, c_int, c_int]),
    "gv_kg_solver_train": (c_int, [c_void_p, c_char_p, c_int, c_int, c_float, c_float, c_float, c_int, c_int, c_float,
                                   c_int]),
    "gv_kg_solver_train_begin": (c_int, [c_void_p, c_char_p, c_int, c_int, c_float, c_float, c_float, c_int, c_int,
                                         c_float, c_int]),
    "gv_kg_solver_train_episode": (c_int, [c_void_p]),
    "gv_kg_solver_train_end": (c_int, [c_void_p]),
    "gv_kg_solver_predict": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "gv_kg_solver_clear": (c_int, [c_void_p]),
    "gv_kg_solver_embeddings": (P(c_float), [c_void_p, c_int, P(c_uint64), P(c_int)]),
    "gv_kg_solver_info": (c_int, [c_void_p, c_char_p, c_size_t]),
    "gv_kg_solver_attributes": (c_int, [c_void_p, c_char_p, c_size_t]),
    "gv_kg_solver_logged_loss": (c_int, [c_void_p, c_void_p, c_int]),
    "gv_kg_solver_stats": (c_int, [c_void_p, c_void_p, c_int]),
    "gv_kg_solver_locations": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gv_kg_solver_pool": (c_int64, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "gv_kg_solver_last_negatives": (c_int, [c_void_p, c_void_p]),
    "gv_kg_schedule": (c_int, [c_int, c_int, c_void_p, c_int]),
    # test hooks
    "gv_schedule_plan": (c_int, [c_int, c_int, c_int, c_void_p, c_int]),
    "gv_reset_global_engine": (None, [c_uint32]),
    "gv_engine_self_check": (c_int, [c_uint32, c_uint64]),
    "gv_solver_locations": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gv_solver_pool": (c_int64, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "gv_solver_train_begin": (c_int, [c_void_p, c_char_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_float, c_int, c_float, c_float, c_int]),
    "gv_solver_train_episode": (c_int, [c_void_p]),
    "gv_solver_train_end": (c_int, [c_void_p]),
    "gv_solver_train_step": (c_int, [c_void_p]),
    "gv_solver_device_timer": (c_double, [c_void_p, c_int]),
    "gv_solver_last_negatives": (c_int, [c_void_p, c_void_p]),
}

for _name, (_restype, _argtypes) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library is stale: rebuild it
    _fn.restype = _restype
    _fn.argtypes = _argtypes


class GVError(RuntimeError):
    """An error reported by libgv_b200 (the reference aborts the process instead)."""


def check(status):
    if status != 0:
        raise GVError(lib.gv_last_error().decode("utf-8", "replace"))
    return status


def last_error():
    return lib.gv_last_error().decode("utf-8", "replace")
