"""The GPU parity tests, EXECUTED on the CPU: the product's kernel sources (graphvite_b200/csrc/*.cu) are
compiled for the host on top of the CUDA emulation in tests/emu (one fiber per CUDA thread; warp
collectives, CTA barriers and a malloc-backed runtime emulated) and the `-m gpu` test files are run
against that build in a child pytest process with GV_EMULATE=1.

What this proves: the kernels' and the host runtime's LOGIC -- indexing, aliasing cases, reductions,
barrier placement (a divergent barrier aborts the emulator), the sampler / fill / solver state machines --
agrees with the oracle, on machines without a GPU.  What it cannot prove: anything about speed, about
hardware-only faults, or about races between CTAs (they run one after another here).  The same files run
on a real B200 with `pytest -m gpu`; the emulated build is test infrastructure like oracle/ and is never
loaded by the package (it lives under tests/emu/_pkg, reached only through GV_EMULATE=1 in conftest.py).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (files, what they cover)
GROUPS = [
    (["tests/test_gpu_kernels.py", "tests/test_gpu_yy_later_kernels.py"],
     "train / sample / predict / rng / move_rows kernels"),
    (["tests/test_gpu_y_fill.py"], "pool-fill kernels (count / scan / scatter, direct)"),
    (["tests/test_gpu_solver.py", "tests/test_gpu_x_solver_more.py"], "GraphSolver end to end vs the oracle"),
    (["tests/test_gpu_zz_kg_kernels.py", "tests/test_gpu_zz_kg_solver.py", "tests/test_gpu_zzzz_kg_full_size.py"],
     "knowledge-graph kernels and solver (the full-size file at its reduced shape)"),
]


# kernel-level files once more with the warps of a CTA and the lanes of a warp visited in reverse order: which warp
# runs ahead of a barrier decides which write-after-read hazards between warps can show up, which lane runs first
# decides whether a shared-memory exchange inside a warp that lacks its __syncwarp() goes unnoticed
GROUPS.append((["tests/test_gpu_zz_kg_kernels.py", "tests/test_gpu_y_fill.py"], "kernels, warps scheduled in reverse"))


@pytest.mark.parametrize("files,what", GROUPS,
                         ids=[g[0][0].split("test_gpu_")[1][:-3] + ("-reverse" if "reverse" in g[1] else "") for g in GROUPS])
def test_gpu_suite_under_cuda_emulation(files, what):
    files = [f for f in files if os.path.exists(os.path.join(ROOT, f))]
    # GV_UPLOAD_CHUNK: the toy graphs' arrays go through the chunked, double-buffered staging path of the big ones
    env = dict(os.environ, GV_EMULATE="1", GV_EMU_BACKTRACE="1", GV_UPLOAD_CHUNK="1024")
    if "reverse" in what:
        env["GV_EMU_WARP_ORDER"] = "reverse"
        env["GV_EMU_LANE_ORDER"] = "reverse"
    result = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files,
                            cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            timeout=1500)
    tail = result.stdout[-3000:]
    assert result.returncode == 0, "%s failed under emulation:\n%s" % (what, tail)
    assert " passed" in tail, tail
