"""Multi-GPU story of the path: shard the independent units (event windows, images) across
ranks, no collective on the data path.  One process per GPU; the only
``torch.distributed`` traffic is the optional gather of per-rank timings / checksums that
``bench.py`` and the tests use."""
from __future__ import annotations

import numpy as np

__all__ = ["shard_round_robin", "shard_lpt", "gather_objects", "bind_to_gpu_numa"]


def shard_round_robin(n_units: int, world_size: int, rank: int) -> list:
    """Unit ``u`` -> rank ``u mod world_size`` (what DistributedSampler does for the whole
    dataset in the reference, builder.py:135-141)."""
    assert 0 <= rank < world_size
    return list(range(rank, n_units, world_size))


def shard_lpt(costs, world_size: int, rank: int) -> list:
    """Longest-processing-time greedy for ragged windows: units sorted by decreasing cost,
    each given to the currently least-loaded rank.  Deterministic (ties break on index)."""
    assert 0 <= rank < world_size
    costs = np.asarray(costs, dtype=np.int64)
    order = sorted(range(len(costs)), key=lambda u: (-int(costs[u]), u))
    load = [0] * world_size
    mine = []
    for u in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += int(costs[u])
        if r == rank:
            mine.append(u)
    return sorted(mine)


def gather_objects(obj, group=None) -> list:
    """all_gather of a small Python object (timings, checksums); identity without a
    process group.  Works on gloo (CPU tests) and nccl."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def bind_to_gpu_numa(device_index: int) -> dict:
    """Restrict this process (one per GPU) to the CPUs NVML reports as local to its GPU, so that the pinned host
    buffers it allocates afterwards land on that GPU's NUMA node (Linux places pages on the node of the allocating
    thread) and the H2D / D2H DMA of the host-buffer path does not cross the inter-socket link.  Never widens the
    set the process was given and never leaves it empty; returns what it did (``{"bound": bool, ...}``)."""
    import os
    info = {"bound": False}
    if os.environ.get("CMDA_NO_NUMA_BIND"):
        return dict(info, skipped="CMDA_NO_NUMA_BIND")
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(device_index)
            bus_id = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        info.update(gpu_local_cpus=len(local), allowed_cpus=len(allowed), target_cpus=len(target))
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
    except Exception as e:            # no NVML, no permission, not Linux: leave the process where it is
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info
