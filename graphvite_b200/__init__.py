"""graphvite_b200 -- a B200-native drop-in for the node-embedding path of GraphVite.

Mirrors the reference's Python surface (python/graphvite/__init__.py:38-49):
``graphvite_b200.graph.Graph``, ``graphvite_b200.solver.GraphSolver``,
``graphvite_b200.optimizer.*``, ``auto``, ``dtype`` -- over a C-ABI shared library
(include/gv_b200.h) whose hot loop is hand-written CUDA for sm_100a.
"""
from ._lib import lib as _clib, GVError, LIB_PATH  # noqa: F401  (fails loudly without the extension)
from .base import auto, dtype, KiB, MiB, GiB, cfg

__version__ = _clib.gv_version().decode()

uint32, uint64, float32, float64 = dtype.uint32, dtype.uint64, dtype.float32, dtype.float64

from . import graph, optimizer, solver  # noqa: E402
from . import application  # noqa: E402

__all__ = ["graph", "optimizer", "solver", "application", "auto", "dtype", "cfg", "GVError",
           "uint32", "uint64", "float32", "float64", "KiB", "MiB", "GiB"]
