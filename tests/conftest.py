import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

# GV_EMULATE=1 (child runs of tests/test_emulated_kernels.py only): import the package from tests/emu/_pkg,
# i.e. the unmodified Python files next to a libgv_b200.so built from the product's sources for the host
# on top of tests/emu's CUDA emulation.  Test infrastructure like the oracle; never a product path.
EMULATED = os.environ.get("GV_EMULATE") == "1"
if EMULATED:
    import subprocess
    subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu", "_pkg"))


def pytest_report_header(config):
    if EMULATED:
        from graphvite_b200 import _lib
        return "graphvite_b200 under test: %s -- %s" % (_lib.LIB_PATH, _lib.lib.gv_version().decode())


def _ensure_built():
    """The shared objects are git-ignored build products: compile them once if a fresh checkout has
    none (same recipe as __graft_entry__.build(); building is not a fallback -- the product still
    refuses to run without its CUDA extension)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "graphvite_b200", "libgv_b200.so")):
        subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "graphvite_b200", "csrc")],
                              stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def toy_graph_file(golden_dir):
    return os.path.join(golden_dir, "toy_graph.txt")
