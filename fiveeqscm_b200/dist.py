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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def numa_cpus_of_gpu(pci_domain: int, pci_bus: int, pci_device: int, sysfs: str = "/sys"):
    """(numa node, CPUs of that node) for the GPU at PCI domain:bus:device, read from sysfs;
    (-1, empty set) when the platform does not say (single-socket hosts report node -1)."""
    import os
    bdf = "%04x:%02x:%02x.0" % (pci_domain, pci_bus, pci_device)
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read().strip())
        if node < 0:
            return -1, set()
        return node, _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
    except (OSError, ValueError):
        return -1, set()


def bind_to_gpu_numa_node(device_index: int):
    """One process per GPU on a multi-socket host: run this rank on the CPUs of the NUMA node its
    GPU hangs off, BEFORE allocating page-locked host buffers (first touch then places them on that
    node), so the host pipeline's H2D / D2H copies do not cross the socket interconnect.
    Returns dict(node, cpus) describing what was done (cpus == 0: left unchanged)."""
    import os

    import torch
    p = torch.cuda.get_device_properties(device_index)
    node, cpus = numa_cpus_of_gpu(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return dict(node=node, cpus=len(allowed))
    except (AttributeError, OSError):
        pass
    return dict(node=node, cpus=0)
