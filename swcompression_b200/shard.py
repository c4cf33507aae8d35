"""Multi-GPU sharding of independent units: contiguous ranges of the unit list balanced by
compressed+decompressed bytes; no data-path collective.  torch.distributed is used only for the barrier, the
max-over-ranks time and (optionally) gathering per-rank result tables."""
import numpy as np


def partition(weights, world):
    """Split range(len(weights)) into `world` contiguous ranges with near-equal total weight.
    Returns a list of (begin, end)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world
    c = np.concatenate([[0.0], np.cumsum(w)])
    total = c[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(c, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def my_range(weights, rank, world):
    return partition(weights, world)[rank]


def reduce_max_time(seconds, device=None):
    """max over ranks of a local duration (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local, device=None):
    """all-gather one integer per rank (e.g. units decoded, error count)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(local)]
    t = torch.tensor([int(local)], dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
