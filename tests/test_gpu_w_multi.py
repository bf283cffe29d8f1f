"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): world_size ranks, one per GPU, NCCL block movement
(and, for the knowledge-graph solver, the NCCL all-reduce of the relation deltas); pools bit-exact and
embeddings equal to the oracle's N-worker run.  The same worker runs on CPU under the CUDA emulation over gloo
(tests/test_emulated_multi_rank.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def launch(world, env):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(ROOT, "tests", "multi_rank_worker.py")]
    result = subprocess.run(command, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert result.returncode == 0, result.stdout[-6000:]
    for rank in range(world):
        assert "rank %d ok" % rank in result.stdout


@pytest.mark.parametrize("world,partitions,mode", [(2, 2, "staged"), (2, 4, "staged"), (2, 2, "replicated"),
                                                   (2, 4, "direct"), (4, 4, "staged"), (8, 8, "staged")])
def test_multi_gpu_matches_oracle(world, partitions, mode):
    """staged: partitioned sampling, pairs of peer-owned blocks staged locally and forwarded with coalesced NVLink
    stores (the default); direct: 8-byte peer stores; replicated: every rank samples all blocks itself (no CUDA IPC)."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, GV_TEST_SOLVER="graph", GV_TEST_PARTITIONS=str(partitions))
    env.pop("GV_EMULATE", None)
    if mode == "replicated":
        env["GV_REPLICATED_SAMPLING"] = "1"
    if mode == "direct":
        env["GV_DIRECT_PEER_SCATTER"] = "1"
    launch(world, env)


@pytest.mark.parametrize("world,partitions", [(2, 2), (4, 4)])
def test_multi_gpu_node2vec_tables_sharded_over_the_ranks(world, partitions):
    """node2vec's per-edge alias tables: every rank builds 1/W of them, the walk kernel reads the peers' shards
    through CUDA IPC; pools bit-exact with the oracle."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, GV_TEST_SOLVER="graph", GV_TEST_PARTITIONS=str(partitions), GV_TEST_MODEL="node2vec")
    env.pop("GV_EMULATE", None)
    launch(world, env)


@pytest.mark.parametrize("world,partitions,optimizer", [(2, 4, "SGD"), (2, 4, "Adam"), (4, 8, "Adam")])
def test_multi_gpu_knowledge_graph_matches_oracle(world, partitions, optimizer):
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, GV_TEST_SOLVER="kg", GV_TEST_PARTITIONS=str(partitions), GV_TEST_OPTIMIZER=optimizer)
    env.pop("GV_EMULATE", None)
    launch(world, env)
