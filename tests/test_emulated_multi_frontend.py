"""`GraphSolver(dim, device_ids=[a, b, ...])` in one process -- the reference's multi-GPU signature
(core/solver.h:184-213) -- starts one worker process per GPU (graphvite_b200/multi.py).  Executed here on the CUDA
emulation with two workers over gloo: build / train / resume / predict / numpy views in shared memory / graph recipes /
close."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_front_end_under_emulation():
    subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    package = os.path.join(ROOT, "tests", "emu", "_pkg")
    env = dict(os.environ, GV_EMULATE="1", GV_EMU_IPC="1", OMP_NUM_THREADS="1", GV_MULTI_TIMEOUT="600",
               PYTHONPATH=package + os.pathsep + ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    done = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multi_frontend_worker.py")], env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert done.returncode == 0 and "front end ok" in done.stdout, done.stdout[-5000:]
