"""Constants of the reference's Python package (python/graphvite/base.py:29-58, src/graphvite.cu:76-101)."""
import enum
from types import SimpleNamespace

auto = 0  # graphvite::kAuto (include/util/common.h:29), exported as module attribute `auto`


class dtype(enum.IntEnum):
    """bind.h:52-57"""
    uint32 = 0
    uint64 = 1
    float32 = 2
    float64 = 3


def KiB(size):
    return int(size) << 10


def MiB(size):
    return int(size) << 20


def GiB(size):
    return int(size) << 30


# global config defaults, python/graphvite/base.py:33-38 (no ~/.graphvite side effects here)
cfg = SimpleNamespace(backend="graphvite", float_type=dtype.float32, index_type=dtype.uint32)
