"""Data-parallel plumbing of the learners (one process per GPU, torch.distributed).

The reference has no data parallelism; the B200 design shards envs and replay per
rank and exchanges exactly one thing per update: the flat gradient arena, summed over ranks
before the global-norm clip and AdamW (which then apply grad_scale = 1/world, so every rank takes
the step of the mean gradient of the concatenated batch and parameters stay identical).
"""
import torch
import torch.distributed as dist


def world_size(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def allreduce_sum_(flat_grad, group=None):
    """In-place sum of the flat gradient arena over the ranks (NCCL over NVLink on GPUs, gloo in
    the CPU tests).  The mean is taken inside the optimiser kernel (grad_scale)."""
    if world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_(flat_params, src=0, group=None):
    """Make rank ``src``'s parameters everyone's (initialisation)."""
    if world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def params_in_sync(flat_params, group=None, atol=0.0):
    """True when every rank holds the same parameters (cheap periodic assertion)."""
    if world_size(group) <= 1:
        return True
    lo, hi = flat_params.clone(), flat_params.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool((hi - lo).abs().max().item() <= atol)
