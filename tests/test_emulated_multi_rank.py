"""world_size > 1 on a machine without GPUs: tests/multi_rank_worker.py under the CUDA emulation of tests/emu
(GV_EMULATE=1), one process per emulated device over gloo.  The ranks run the product's own multi-GPU code
paths -- block directory and NCCL-shaped exchange of the node-embedding solver; entity-block movement,
tied-weight schedule and relation all-reduce of the knowledge-graph solver -- and compare pools (bit-exact) and
embeddings (rtol 1e-3) with the oracle's N-worker emulation.  On real GPUs the same worker runs over NCCL
(tests/test_gpu_w_multi.py)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# optimizer column of the node-embedding cases: "ipc" = partitioned sampling through emulated CUDA IPC (device
# memory in POSIX shared memory, the peer-exchange kernels of different processes really wait for each other),
# "replicated" = every rank samples all blocks itself (the fallback when IPC is unavailable)
# "node2vec" = "ipc" with the biased walk: the per-edge alias tables are sharded over the ranks and read through IPC
# "direct" = "ipc" with GV_DIRECT_PEER_SCATTER=1 (8-byte peer stores instead of staging + coalesced forwarding)
CASES = [("graph", 2, 2, "ipc"), ("graph", 2, 4, "ipc"), ("graph", 2, 2, "replicated"), ("graph", 4, 4, "node2vec"),
         ("graph", 2, 4, "direct"),
         ("kg", 2, 4, "SGD"), ("kg", 2, 4, "Adam"), ("kg", 4, 8, "Adam")]


@pytest.mark.parametrize("solver,world,partitions,optimizer", CASES,
                         ids=["%s-w%d-p%d-%s" % case for case in CASES])
def test_multi_rank_under_emulation(solver, world, partitions, optimizer):
    subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], stdout=subprocess.DEVNULL)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, GV_EMULATE="1", GV_EMU_BACKTRACE="1", GV_TEST_SOLVER=solver,
               GV_TEST_PARTITIONS=str(partitions), GV_TEST_OPTIMIZER=optimizer, OMP_NUM_THREADS="1",
               GV_EMU_IPC="1" if optimizer in ("ipc", "node2vec", "direct") else "0", GV_LOG="1",
               GV_TEST_MODEL="node2vec" if optimizer == "node2vec" else "LINE")
    if optimizer == "direct":
        env["GV_DIRECT_PEER_SCATTER"] = "1"
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(ROOT, "tests", "multi_rank_worker.py")]
    result = subprocess.run(command, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert result.returncode == 0, result.stdout[-6000:]
    for rank in range(world):
        assert "rank %d ok" % rank in result.stdout
    if solver == "graph":  # the solver reports when it could not map the peers' pools
        assert ("falling back to replicated sampling" in result.stdout) == (optimizer == "replicated")
    if optimizer == "node2vec":
        assert "sharded over %d ranks" % world in result.stdout
