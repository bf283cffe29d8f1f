"""bench.py's orchestration -- K timed steps, collective teardown of the steady-state solver, a second solver for the
end-to-end leg, one JSON line from rank 0 -- executed on the CPU at world_size 1 and 2 (tests/bench_emulated_worker.py:
CUDA emulation + gloo + emulated CUDA IPC).  The numbers are meaningless here; the keys of the contract and the
bookkeeping (edges trained, launches counted, partitions) are checked."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"]


@pytest.mark.parametrize("world", [1, 2])
def test_bench_flow_under_emulation(world):
    subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    worker = os.path.join(ROOT, "tests", "bench_emulated_worker.py")
    env = dict(os.environ, GV_EMULATE="1", GV_EMU_IPC="1", GV_EMU_BACKTRACE="1", OMP_NUM_THREADS="1")
    if world == 1:
        command = [sys.executable, worker]
    else:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                   "--master-addr", "127.0.0.1", "--master-port", str(port), worker]
    result = subprocess.run(command, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert result.returncode == 0, result.stderr[-4000:]
    lines = [line for line in result.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1, result.stdout[-2000:]  # rank 0 alone prints
    line = json.loads(lines[0])
    for key in KEYS:
        assert key in line, key
    assert line["n_gpus"] == world and line["steps"] == 4 and line["warmup"] == 2
    assert line["config"]["num_partition"] == world and line["scaling"] == "weak"
    assert line["value"] > 0 and line["gpu_launches"] > 0
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["achieved"] > 0
    assert line["e2e"]["value"] > 0 and line["e2e"]["edges"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert "end-to-end train() done" in result.stderr
    assert line["model_quality"]["auc"] > 0 and line["model_quality"]["vertex_norm"] > 0
    if world > 1:  # the multi-rank parity self-check against the oracle's N-worker emulation ran and passed
        assert line["parity_ok"] is True, line.get("parity")
        assert set(line["parity"]) >= {"line_P2", "line_P4", "node2vec_P2", "rotate_adam_P4"} and not line["parity_not_run"]
