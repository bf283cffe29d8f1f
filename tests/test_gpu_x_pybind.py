"""Training through the compiled pybind11 module `libgraphvite` (graphvite_b200/pybind/libgraphvite.so): the same
C ABI as the ctypes mirror, so with the same engine seed the two surfaces draw the same pools and initial embeddings;
the Hogwild training itself is racy, so trained norms agree statistically.  Also: numpy views alias solver memory
(bind.h:90-106), predict() scores with the written-back matrices, a custom Python LR schedule is called back without
the GIL being held by train().  The bodies run in a child interpreter (tests/pybind_gpu_worker.py): the reference's
own module is called `libgraphvite` too and other tests of the session import it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "graphvite_b200", "pybind", "libgraphvite.so")


@pytest.mark.parametrize("case", ["graph", "schedule_kg"])
def test_training_through_the_pybind_module(case):
    if os.environ.get("GV_EMULATE") == "1":
        pytest.skip("the pybind module links the nvcc build of libgv_b200")
    if not os.path.exists(PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "graphvite_b200", "csrc"), "pybind"])
    done = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pybind_gpu_worker.py"), case],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert done.returncode == 0 and done.stdout.strip().endswith("ok"), done.stdout[-4000:]
