# coding=utf-8
"""Data-parallel plumbing: one process per GPU, utterances sharded
over ranks, ONE all-reduce(SUM) of the flat gradient buffer per step.  The reference
differentiates the SUM of the per-item losses and accumulates mini-batch gradients by
addition (models/AcousticModel.py:386-401), so summing over ranks makes N ranks x 1
mini-batch identical to the reference run with mini_batch_size = N.  torch.distributed
(NCCL on GPUs, gloo in the CPU tests) is the transport; there is no other exchange.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items, rank=None, world_size=None):
    """Rank r takes items r, r + N, r + 2N, ... (each rank keeps whole mini-batches)."""
    if rank is None:
        rank, world_size = world()
    return items[rank::world_size]


def allreduce_sum_(flat):
    """In-place SUM over ranks of a flat tensor (gradients or the 3 step accumulators)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat
