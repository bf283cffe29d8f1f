"""Multi-rank host logic on CPU: the block directory / rotation plan for 2..8 workers, and the
exchange callback moving real buffers between two gloo processes (world_size = 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("num_worker,num_partition", [(1, 1), (1, 3), (2, 2), (2, 4), (4, 4), (4, 8), (8, 8), (8, 16)])
def test_rotation_plan_is_consistent(num_worker, num_partition):
    from graphvite_b200 import distributed
    episodes = 3
    plan = distributed.schedule_plan(num_partition, num_worker, episodes)
    reference = O.schedule(num_partition, num_worker)  # oracle restatement of get_schedule
    steps = reference.shape[0]
    assert plan.shape == (steps * episodes, reference.shape[1], 6)
    groups = max(1, num_partition // num_worker)
    owner = {h: h % num_worker for h in range(num_partition)}
    for index, step in enumerate(plan):
        np.testing.assert_array_equal(step[:, :2], reference[index % steps])
        sends = {}
        for rank, (head, tail, source, give, destination, held) in enumerate(step):
            assert tail % num_worker == rank or num_partition == 1  # context blocks never move
            assert source == owner[head]                            # the block comes from where it is
            if give >= 0:
                assert owner[give] == rank and destination != rank
                sends[(rank, destination)] = give
        for rank, (head, tail, source, give, destination, held) in enumerate(step):
            if source != rank:  # every receive has the matching send
                assert sends.pop((source, rank)) == head
        assert not sends
        for rank, (head, *_rest) in enumerate(step):
            owner[head] = rank
        for rank in range(step.shape[0]):
            held = sum(1 for h in owner if owner[h] == rank)
            assert held == step[rank, 5] == groups  # one block per group per rank, always


def test_plan_rejects_bad_partition_counts():
    from graphvite_b200 import distributed, GVError
    with pytest.raises(GVError):
        distributed.schedule_plan(3, 2)
    with pytest.raises(GVError):
        distributed.schedule_plan(1, 2)


WORKER = r"""
import ctypes, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["GV_ROOT"])
from graphvite_b200 import distributed

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
exchange = distributed.make_exchange(device=None)
peer = 1 - rank
nbytes = 1 << 20
send = np.full(nbytes, rank + 1, dtype=np.uint8)
recv = np.zeros(nbytes, dtype=np.uint8)
# ring shift: both directions at once, like a sub-episode boundary
assert exchange(send.ctypes.data, peer, recv.ctypes.data, peer, nbytes, None, None) == 0
assert (recv == peer + 1).all()
# one-directional legs
recv[:] = 0
if rank == 0:
    assert exchange(send.ctypes.data, 1, None, -1, nbytes, None, None) == 0
else:
    assert exchange(None, -1, recv.ctypes.data, 0, nbytes, None, None) == 0
    assert (recv == 1).all()
# follow the rotation plan of a 2-worker, 4-partition solver with tagged blocks
plan = distributed.schedule_plan(4, 2, 2)
blocks = {h: np.full(4096, h, dtype=np.uint8) for h in range(4) if h % 2 == rank}
spare = np.zeros(4096, dtype=np.uint8)
for step in plan:
    head, tail, source, give, destination, held = step[rank]
    incoming = spare if source != rank else None
    status = exchange(blocks[give].ctypes.data if give >= 0 else None, destination,
                      incoming.ctypes.data if incoming is not None else None, source if incoming is not None else -1,
                      4096, None, None)
    assert status == 0
    if give >= 0:
        spare = blocks.pop(give)
    if incoming is not None:
        blocks[head] = incoming
    assert (blocks[head] == head).all(), (rank, head)
    assert len(blocks) == held
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
"""


def test_exchange_callback_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    processes = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   GV_ROOT=ROOT)
        processes.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
    for rank, process in enumerate(processes):
        output, _ = process.communicate(timeout=180)
        assert process.returncode == 0, output
        assert "rank %d ok" % rank in output
