"""Multi-GPU sharding of independent units.

Units (Deflate blocks / gzip members, LZ4 blocks, .bz2 streams, .xz streams) are independent, so decoding needs no
exchange: rank 0 owns the unit list, cuts it into contiguous ranges balanced by compressed+decompressed bytes
(`partition`), broadcasts the tables and sends every rank its shard of compressed bytes (`scatter_units`); each rank
decodes its shard with the single-GPU batched call; decoded buffers stay sharded, or travel to rank 0 (`gather_to_root`)
or to everybody (`allgather`).  The collectives are torch.distributed calls: NCCL over NVLink on GPUs, gloo on CPU for the
host-logic tests."""
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


def scatter_units(buf, lens, caps, device, src=0, extra=None):
    """Rank `src` holds `buf` (uint8 tensor on `device`: all compressed units back to back, unit i at sum(lens[:i])... or
    any layout described by `offs`) — here: units are contiguous in unit order, so a shard is one byte range.

    buf/lens/caps (and `extra`, one more int64 per unit, e.g. the unpadded length) are only read on rank `src` (others pass
    None).  Every rank gets back
        (local_buf, local_lens, local_caps, (begin, end))        [+ local_extra when `extra` was given on `src`]
    where local_buf is a uint8 tensor on `device` with this rank's units back to back and begin/end its unit range.
    Collectives: one broadcast of the table sizes, one of the (lens, caps) tables, one send/recv per non-root rank."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    hdr = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0] = len(lens); hdr[1] = 0 if extra is None else 1
    dist.broadcast(hdr, src=src)
    n, has_extra = int(hdr[0].item()), bool(hdr[1].item())
    cols = 3 if has_extra else 2
    table = torch.zeros(cols * n, dtype=torch.int64, device=device)
    if rank == src:
        table[:n] = torch.as_tensor(np.asarray(lens, dtype=np.int64), device=device)
        table[n:2 * n] = torch.as_tensor(np.asarray(caps, dtype=np.int64), device=device)
        if has_extra:
            table[2 * n:] = torch.as_tensor(np.asarray(extra, dtype=np.int64), device=device)
    dist.broadcast(table, src=src)                                   # "ncclBroadcast of the offset tables"
    t = table.cpu().numpy()
    all_lens, all_caps = t[:n], t[n:2 * n]
    parts = partition(all_lens + all_caps, world)
    starts = np.concatenate([[0], np.cumsum(all_lens)])
    b, e = parts[rank]
    nbytes = int(starts[e] - starts[b])
    if rank == src:
        ops = []
        for r in range(world):
            rb, re = parts[r]
            if r != src and re > rb:
                ops.append(dist.P2POp(dist.isend, buf[int(starts[rb]):int(starts[re])], r))
        local = buf[int(starts[b]):int(starts[e])].clone()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    else:
        local = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if nbytes:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, local, src)]):
                w.wait()
    if has_extra:
        return local, all_lens[b:e].copy(), all_caps[b:e].copy(), (b, e), t[2 * n:][b:e].copy()
    return local, all_lens[b:e].copy(), all_caps[b:e].copy(), (b, e)


def gather_to_root(local, dst=0):
    """Decoded shard -> rank `dst`.  Returns the list of per-rank tensors on `dst` (None elsewhere).  Sizes may differ per rank:
    the sizes travel first (all_gather of one integer), the payload with grouped send/recv."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = gather_counts(local.numel(), device=local.device)
    if rank == dst:
        outs = [local if r == dst else torch.empty(sizes[r], dtype=local.dtype, device=local.device) for r in range(world)]
        ops = [dist.P2POp(dist.irecv, outs[r], r) for r in range(world) if r != dst and sizes[r]]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return outs
    if local.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst)]):
            w.wait()
    return None


def allgather(local):
    """Decoded shard -> every rank.  Uniform sizes use one all_gather_into_tensor (ncclAllGather); ragged sizes a list all_gather
    padded to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = gather_counts(local.numel(), device=local.device)
    if len(set(sizes)) == 1:
        out = torch.empty(world * sizes[0], dtype=local.dtype, device=local.device)
        if sizes[0]:
            dist.all_gather_into_tensor(out, local)
        return [out[r * sizes[0]:(r + 1) * sizes[0]] for r in range(world)]
    m = max(sizes)
    padded = torch.zeros(m, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    outs = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(outs, padded)
    return [outs[r][:sizes[r]] for r in range(world)]
