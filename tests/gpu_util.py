"""Helpers that drive the C ABI of libgv_b200 with torch-owned device memory (tests, smoke, bench).
PyTorch is plumbing here: it owns the device buffers and the stream; the kernels are ours."""
import ctypes
import os

import numpy as np
import torch

from graphvite_b200 import _lib

lib = _lib.lib

# GV_EMULATE=1 (set by tests/test_emulated_kernels.py for its child pytest runs): the package on sys.path
# is tests/emu/_pkg, whose libgv_b200.so is the product's sources compiled for the host on top of the CUDA
# emulation in tests/emu -- "device" memory is then host memory and torch only lends CPU tensors.
EMULATED = os.environ.get("GV_EMULATE") == "1"
DEVICE = "cpu" if EMULATED else "cuda"


def to_device(tensor):
    """torch tensor -> the device the kernels run on (a private copy in both modes)."""
    return tensor.clone() if EMULATED else tensor.cuda()


def zeros(*shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device=DEVICE)


def full(shape, value, dtype):
    return torch.full(shape, value, dtype=dtype, device=DEVICE)


def synchronize():
    if not EMULATED:
        torch.cuda.synchronize()


def dev(array, dtype=None):
    """numpy -> device tensor (keeps unsigned data by viewing it as the signed type of equal width)."""
    array = np.ascontiguousarray(array if dtype is None else array.astype(dtype))
    views = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}
    if array.dtype in views:
        array = array.view(views[array.dtype])
    return to_device(torch.from_numpy(array))


def host(tensor, dtype):
    return tensor.cpu().numpy().view(dtype)


def stream_pointer():
    if EMULATED:
        return None
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def alias_entries(prob, alias):
    """interleave prob / alias into gv_alias_entry_t[]"""
    table = np.zeros(len(prob), dtype=[("prob", np.float32), ("alias", np.uint32)])
    table["prob"] = prob
    table["alias"] = alias
    return table


def run_train_block(dim, vertex, context, moments, batch, negatives, optimizer, negative_weight, lr=None,
                    batch_size=None, num_warps=0, random=None, negative_table=None, per_sample_loss=True):
    """Run gv_cuda_train_block on copies of the numpy matrices; returns the updated copies.
    optimizer = (type, lr, weight_decay, a, b, epsilon).  Either `negatives` [n][k] or
    (`random` [n*k*2] doubles, `negative_table` (prob, alias)) selects the negative source."""
    otype, olr, wd, a, b, eps = optimizer
    n = batch.shape[0]
    batch_size = batch_size or max(1, n)
    num_batch = (n + batch_size - 1) // batch_size
    if lr is None:
        lr = np.full(num_batch, olr, dtype=np.float32)
    d_vertex, d_context = dev(vertex), dev(context)
    names = ["vm1", "cm1", "vm2", "cm2"]
    d_moments = [dev(m) if m is not None else None for m in (moments or [None] * 4)]
    matrices = _lib.Matrices()
    matrices.dim = dim
    matrices.vertex, matrices.context = d_vertex.data_ptr(), d_context.data_ptr()
    for field, tensor in zip(["vertex_m1", "context_m1", "vertex_m2", "context_m2"], d_moments):
        setattr(matrices, field, tensor.data_ptr() if tensor is not None else None)
    d_batch = dev(batch, np.uint32)
    d_lr = dev(np.asarray(lr, dtype=np.float32))
    d_loss = torch.zeros(max(1, n), dtype=torch.float32, device=DEVICE) if per_sample_loss else None
    d_batch_loss = torch.zeros(num_batch, dtype=torch.float32, device=DEVICE)
    device_optimizer = _lib.DeviceOptimizer(otype, wd, a, b, eps)
    if negatives is not None:
        k = negatives.size // n if n else 0
        d_negatives, d_random, d_table, count, d_out = dev(negatives, np.uint32), None, None, 0, None
    else:
        prob, alias = negative_table
        k = len(random) // (2 * n)
        d_negatives = None
        d_random = dev(np.asarray(random, dtype=np.float64))
        d_table = to_device(torch.from_numpy(alias_entries(prob, alias).view(np.int64)))
        count = len(prob)
        d_out = torch.zeros(n * k, dtype=torch.int32, device=DEVICE)
    _lib.check(lib.gv_cuda_train_block(
        ctypes.byref(matrices), d_batch.data_ptr(), n, k,
        d_negatives.data_ptr() if d_negatives is not None else None,
        d_random.data_ptr() if d_random is not None else None,
        d_table.data_ptr() if d_table is not None else None, count,
        d_out.data_ptr() if d_out is not None else None,
        ctypes.byref(device_optimizer), d_lr.data_ptr(), batch_size, float(negative_weight),
        d_loss.data_ptr() if d_loss is not None else None, d_batch_loss.data_ptr(), num_warps, stream_pointer()))
    synchronize()
    result = {"vertex": d_vertex.cpu().numpy(), "context": d_context.cpu().numpy(),
              "loss": d_loss.cpu().numpy()[:n] if d_loss is not None else None,
              "batch_loss": d_batch_loss.cpu().numpy()}
    for name, tensor in zip(names, d_moments):
        result[name] = tensor.cpu().numpy() if tensor is not None else None
    if negatives is None:
        result["negatives"] = host(d_out, np.uint32)
    return result


def sample_negatives(prob, alias, random):
    n = len(random) // 2
    d_table = to_device(torch.from_numpy(alias_entries(prob, alias).view(np.int64)))
    d_random = dev(np.asarray(random, dtype=np.float64))
    d_out = torch.zeros(n, dtype=torch.int32, device=DEVICE)
    _lib.check(lib.gv_cuda_sample_negatives(d_table.data_ptr(), len(prob), d_random.data_ptr(), n, d_out.data_ptr(),
                                            stream_pointer()))
    synchronize()
    return host(d_out, np.uint32)


def predict(dim, vertex, context, batch):
    d_vertex, d_context, d_batch = dev(vertex), dev(context), dev(batch, np.uint32)
    n = batch.shape[0]
    d_logits = torch.zeros(n, dtype=torch.float32, device=DEVICE)
    _lib.check(lib.gv_cuda_predict(dim, d_vertex.data_ptr(), d_context.data_ptr(), d_batch.data_ptr(), n,
                                   d_logits.data_ptr(), stream_pointer()))
    synchronize()
    return d_logits.cpu().numpy()
