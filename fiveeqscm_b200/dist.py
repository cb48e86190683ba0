"""Multi-GPU plumbing: shard the member axis, reduce the summary statistics.

Members are independent, so ranks never exchange data during the integration.  The only
collective is at the end: integer histogram counts are summed (bitwise independent of the GPU
count), moment sums are summed, and min/max are reduced with MAX over [max, -min].
Works over any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_bounds(n_member: int, world_size: int, rank: int, align: int = 128):
    """Contiguous member block of `rank`: [lo, hi).  Block edges are multiples of `align`
    (a CTA's worth of members) except the last one, so no rank runs a ragged CTA but the last."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    blocks = (n_member + align - 1) // align
    lo_b = (blocks * rank) // world_size
    hi_b = (blocks * (rank + 1)) // world_size
    return min(lo_b * align, n_member), min(hi_b * align, n_member)


def allreduce_stats(hist, moments, group=None):
    """In-place all-reduce of (hist int64 [n_t][bins], moments float64 [n_t][4]) across ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hist, moments
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    sums = moments[:, 0:2].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    ext = torch.stack([moments[:, 3], -moments[:, 2]], dim=1).contiguous()
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    moments[:, 0:2] = sums
    moments[:, 3] = ext[:, 0]
    moments[:, 2] = -ext[:, 1]
    return hist, moments
