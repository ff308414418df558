# coding=utf-8
"""Data-parallel plumbing: one process per GPU, utterances sharded
over ranks, ONE all-reduce(SUM) of the flat gradient buffer per step.  The reference
differentiates the SUM of the per-item losses and accumulates mini-batch gradients by
addition (models/AcousticModel.py:386-401), so summing over ranks makes N ranks x 1
mini-batch identical to the reference run with mini_batch_size = N.

Transport: device buffers go through the library's own communicator (rs_comm_init /
rs_allreduce_sum: NCCL bound inside librnnspeech_b200.so, enqueued on the caller's CUDA stream);
torch.distributed is the rendezvous (it carries the 128-byte NCCL id) and the transport of the CPU
(gloo) tests.  RS_COMM=torch forces torch.distributed for device buffers too.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

_COMM = {"handle": None, "key": None}


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items, rank=None, world_size=None):
    """Rank r takes items r, r + N, r + 2N, ... of the first floor(len / N) * N items: every rank gets the SAME
    number of items, hence the same number of mini-batches and optimizer steps (the collectives stay in lockstep)."""
    if rank is None:
        rank, world_size = world()
    usable = (len(items) // world_size) * world_size
    return items[:usable][rank::world_size]


def _library_comm(device):
    """The in-library NCCL communicator of this process (created on first use, collectively)."""
    rank, n = world()
    key = (rank, n, int(device.index if device.index is not None else torch.cuda.current_device()))
    if _COMM["handle"] is not None and _COMM["key"] == key:
        return _COMM["handle"]
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.call("rs_comm_unique_id", buf, 128)
    box = [bytes(buf.raw) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    handle = ctypes.c_void_p()
    with torch.cuda.device(device):
        _lib.call("rs_comm_init", ctypes.byref(handle), box[0], 128, rank, n)
    _COMM["handle"], _COMM["key"] = handle, key
    return handle


def close():
    if _COMM["handle"] is not None:
        _lib.raw("rs_comm_destroy")(_COMM["handle"])
        _COMM["handle"] = _COMM["key"] = None


def allreduce_sum_(flat):
    """In-place SUM over ranks of a flat float32 tensor (gradients or the step accumulators)."""
    _, n = world()
    if n == 1:
        return flat
    if flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and os.environ.get("RS_COMM", "") != "torch":
        comm = _library_comm(flat.device)
        _lib.call("rs_allreduce_sum", comm, flat.data_ptr(), flat.numel(), torch.cuda.current_stream(flat.device).cuda_stream)
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat
