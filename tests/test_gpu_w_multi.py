"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): world_size ranks, one per GPU, NCCL block
rotation; pools of the owned blocks bit-exact and embeddings equal to the oracle's N-worker run."""
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


@pytest.mark.parametrize("world,partitions,replicated", [(2, 2, 0), (2, 4, 0), (2, 2, 1), (4, 4, 0), (8, 8, 0)])
def test_multi_gpu_matches_oracle(world, partitions, replicated):
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, GV_TEST_PARTITIONS=str(partitions))
    if replicated:  # every rank samples all blocks itself (no CUDA IPC): the fallback path
        env["GV_REPLICATED_SAMPLING"] = "1"
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    result = subprocess.run(command, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert result.returncode == 0, result.stdout[-6000:]
    for rank in range(world):
        assert "rank %d ok" % rank in result.stdout
