"""One process per GPU: the block exchange between sub-episodes over torch.distributed.

The reference moves partition blocks GPU -> pageable host -> CPU scatter / gather -> GPU
(include/core/solver.h:1349-1428).  Here a rank hands the vertex block it just trained straight
to the rank that trains it next: NCCL send/recv over NVLink / NVSwitch, issued on the solver's own
CUDA stream.  PyTorch is only the plumbing (process group + NCCL binding); with the gloo backend
the same code moves host buffers, which is how the CPU tests exercise it.
"""
import ctypes
import sys
import traceback

import numpy as np

from . import _lib


class _DevicePointer(object):
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, pointer, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (pointer, False),
                                         "version": 3, "strides": None}


def _as_tensor(pointer, nbytes, device):
    import torch
    if device is None:  # host memory (gloo)
        buffer = (ctypes.c_uint8 * nbytes).from_address(pointer)
        return torch.from_numpy(np.ctypeslib.as_array(buffer))
    return torch.as_tensor(_DevicePointer(pointer, nbytes), device=device)


def make_exchange(device=None, group=None):
    """Build the gv_exchange_fn callback.  `device` = torch device of this rank, or None to move
    host buffers (gloo).  Returns the ctypes callback (keep a reference to it)."""
    import torch
    import torch.distributed as dist

    def exchange(send, dst, recv, src, nbytes, stream, ctx):
        try:
            def run():
                ops = []
                if dst >= 0 and send:
                    ops.append(dist.P2POp(dist.isend, _as_tensor(send, nbytes, device), dst, group))
                if src >= 0 and recv:
                    ops.append(dist.P2POp(dist.irecv, _as_tensor(recv, nbytes, device), src, group))
                if ops:
                    for request in dist.batch_isend_irecv(ops):
                        request.wait()
            if device is None:
                run()
            else:
                # make NCCL order itself after / before the solver's stream instead of the default one
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
                    run()
            return 0
        except Exception:  # never let an exception cross the C boundary
            traceback.print_exc(file=sys.stderr)
            return -1

    return _lib.EXCHANGE_FN(exchange)


_host_group = None


def make_host_allgather():
    """gv_host_allgather_fn over a gloo group of its own (host buffers): trades the CUDA IPC handles of the sample-pool
    arenas once per build(), and is the barrier of the ranks' SAMPLER threads twice per episode -- those run next to
    the main threads' block exchange, so they must not share a process group with it (two threads issuing collectives
    on one group would interleave differently on different ranks).  Must be created by all ranks collectively."""
    import torch
    import torch.distributed as dist
    global _host_group
    if _host_group is None:
        _host_group = dist.new_group(backend="gloo")
    group = _host_group
    world = dist.get_world_size()

    def allgather(send, recv, nbytes, ctx):
        try:
            source = _as_tensor(send, nbytes, None).clone()
            chunks = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(chunks, source, group=group)
            _as_tensor(recv, nbytes * world, None).copy_(torch.cat(chunks))
            return 0
        except Exception:
            traceback.print_exc(file=sys.stderr)
            return -1

    return _lib.HOST_ALLGATHER_FN(allgather)


def attach(solver, device=None, group=None):
    """Wire a GraphSolver created with world_size > 1 to the default process group: the NCCL block
    exchange and the host all-gather that enables partitioned sampling over NVLink peer memory."""
    callback = make_exchange(device, group)
    _lib.check(_lib.lib.gv_solver_set_exchange(solver._handle, callback, None))
    gather = make_host_allgather()
    _lib.check(_lib.lib.gv_solver_set_host_allgather(solver._handle, gather, None))
    solver._exchange = (callback, gather)
    return solver


def make_allreduce(device=None, group=None):
    """Build the gv_allreduce_fn callback: sum `count` floats in place over all ranks (the relation deltas of
    the knowledge-graph solver).  `device` = torch device of this rank, or None for host buffers (gloo)."""
    import torch
    import torch.distributed as dist

    def allreduce(pointer, count, stream, ctx):
        try:
            def run():
                tensor = _as_tensor(pointer, count * 4, device).view(torch.float32)
                dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
            if device is None:
                run()
            else:
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
                    run()
            return 0
        except Exception:
            traceback.print_exc(file=sys.stderr)
            return -1

    return _lib.ALLREDUCE_FN(allreduce)


def attach_knowledge_graph(solver, device=None, group=None):
    """Wire a KnowledgeGraphSolver created with world_size > 1 to the default process group: entity blocks
    move by NCCL send / recv, relation deltas are summed by an NCCL all-reduce, both on the solver's stream."""
    exchange = make_exchange(device, group)
    reduce = make_allreduce(device, group)
    _lib.check(_lib.lib.gv_kg_solver_set_exchange(solver._handle, exchange, None))
    _lib.check(_lib.lib.gv_kg_solver_set_allreduce(solver._handle, reduce, None))
    solver._exchange = (exchange, reduce)
    return solver


def kg_schedule(num_partition, num_worker):
    """SolverMixin::get_schedule for tied weights as an int array [steps, worker, 2] = (head, tail) partition."""
    capacity = (num_partition * num_partition * 4 + 16) * num_worker * 2
    out = np.zeros(capacity, dtype=np.int32)
    steps = _lib.lib.gv_kg_schedule(num_partition, num_worker, out.ctypes.data, capacity)
    if steps < 0:
        raise _lib.GVError(_lib.last_error())
    width = 1 if num_partition == 1 else num_worker
    return out[:steps * width * 2].reshape(steps, width, 2)


def schedule_plan(num_partition, num_worker, num_episode=1):
    """The (head, tail, source, give, destination, held) table of gv_schedule_plan as an int array
    [episode * steps, worker, 6]."""
    capacity = num_partition * num_partition * num_worker * 6 * num_episode + 64
    out = np.zeros(capacity, dtype=np.int32)
    steps = _lib.lib.gv_schedule_plan(num_partition, num_worker, num_episode, out.ctypes.data, capacity)
    if steps < 0:
        raise _lib.GVError(_lib.last_error())
    width = 1 if num_partition == 1 else num_worker
    return out[:steps * num_episode * width * 6].reshape(steps * num_episode, width, 6)
