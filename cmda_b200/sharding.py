"""Multi-GPU story of the path: shard the independent units (event windows, images) across
ranks, no collective on the data path.  One process per GPU; the only
``torch.distributed`` traffic is the optional gather of per-rank timings / checksums that
``bench.py`` and the tests use."""
from __future__ import annotations

import numpy as np

__all__ = ["shard_round_robin", "shard_lpt", "gather_objects"]


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
