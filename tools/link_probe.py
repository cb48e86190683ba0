#!/usr/bin/env python
"""Host <-> device link ceiling of this box, per rank and in aggregate, with N ranks copying at once.

    python tools/link_probe.py                                   one GPU
    torchrun --nproc-per-node N tools/link_probe.py              N GPUs, every rank copying at once

Prints one JSON object (rank 0): for contiguous and pitched copies, each direction alone and both
directions together, the GB/s of the slowest and fastest rank and the sum over ranks.  Also records
what the platform says about NUMA placement (sysfs), which is what the host pipeline's binding uses.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def probe_all(L, local, world, dist, torch, nbytes=256 << 20, reps=8):
    out = {}
    g = (C.c_double * 2)()
    for name, rows in (("contiguous", 1), ("pitched_2208x", 2208)):
        nb = nbytes // max(rows, 1) // 16 * 16 * max(rows, 1)
        for mode, key in ((0, "h2d"), (1, "d2h"), (2, "both")):
            if world > 1:
                dist.barrier()
            from fiveeqscm_b200 import _abi
            _abi.check(L.ufair_link_probe(local, nb if mode != 1 else 0, nb if mode != 0 else 0, rows, reps, g, None))
            v = torch.tensor([g[0], g[1]], dtype=torch.float64, device="cuda")
            if world > 1:
                all_v = [torch.zeros_like(v) for _ in range(world)]
                dist.all_gather(all_v, v)
                v = torch.stack(all_v)
            else:
                v = v[None]
            v = v.cpu().numpy()
            rec = {}
            for col, d in ((0, "h2d"), (1, "d2h")):
                if mode == 2 or d == key:
                    rec[d] = {"min": float(v[:, col].min()), "max": float(v[:, col].max()), "sum": float(v[:, col].sum())}
            out["%s_%s" % (name, key)] = rec
    L.ufair_link_probe(local, 0, 0, 0, 1, g, None)
    return out


def main():
    import torch
    import torch.distributed as dist
    from fiveeqscm_b200 import _abi, dist as D
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    binding = D.bind_rank_to_cpus(local, world) if hasattr(D, "bind_rank_to_cpus") else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _abi.lib()
    res = probe_all(L, local, world, dist, torch)
    p = torch.cuda.get_device_properties(local)
    info = {"world": world, "binding": binding, "cpus_allowed": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(),
            "gpu_pci": "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id),
            "gpu_numa": D.numa_cpus_of_gpu(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)[0], "probe_gbs": res}
    if world > 1:
        infos = [None] * world
        dist.all_gather_object(infos, {k: info[k] for k in ("gpu_pci", "gpu_numa", "cpus_allowed", "binding")})
        info["ranks"] = infos
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(info))


if __name__ == "__main__":
    main()
