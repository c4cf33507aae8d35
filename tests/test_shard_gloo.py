"""N>1 host logic on CPU: contiguous byte-balanced sharding + the max-over-ranks reduction, world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from swcompression_b200 import shard


def test_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    w = rng.integers(1000, 70000, size=10007)
    for world in (1, 2, 3, 8):
        parts = shard.partition(w, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(w)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sums = [w[a:b].sum() for a, b in parts]
        assert max(sums) - min(sums) <= 2 * w.max()
    assert shard.partition([], 4) == [(0, 0)] * 4
    assert shard.partition([5], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = np.arange(1, 1001)
    a, b = shard.my_range(w, rank, world)
    counts = shard.gather_counts(b - a)
    tmax = shard.reduce_max_time(0.5 + rank)
    q.put((rank, a, b, counts, tmax))
    dist.destroy_process_group()


def _scatter_worker(rank, world, port, q):
    """rank 0 owns ragged zlib units; scatter -> every rank inflates its shard (zlib stands in for the GPU call: this test is
    about the host logic) -> gather to root and allgather reproduce the unit order."""
    import zlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    raws = [bytes(rng.integers(0, 7, size=int(rng.integers(0, 5000)), dtype=np.uint8)) for _ in range(37)]
    buf = lens = caps = None
    if rank == 0:
        units = [zlib.compress(r) for r in raws]
        buf = torch.from_numpy(np.frombuffer(b"".join(units), dtype=np.uint8).copy())
        lens = [len(u) for u in units]
        caps = [len(r) for r in raws]
    local, llens, lcaps, (b, e) = shard.scatter_units(buf, lens, caps, torch.device("cpu"))
    off = np.concatenate([[0], np.cumsum(llens)])
    blob = local.numpy().tobytes()
    outs = [zlib.decompress(blob[off[i]:off[i + 1]]) for i in range(e - b)]
    ok = outs == raws[b:e] and [len(o) for o in outs] == list(lcaps)
    mine = torch.from_numpy(np.frombuffer(b"".join(outs), dtype=np.uint8).copy()) if outs and sum(map(len, outs)) else torch.zeros(0, dtype=torch.uint8)
    g = shard.gather_to_root(mine)
    if rank == 0:
        ok = ok and b"".join(t.numpy().tobytes() for t in g) == b"".join(raws)
    a = shard.allgather(mine)
    ok = ok and b"".join(t.numpy().tobytes() for t in a) == b"".join(raws)
    u = shard.allgather(torch.full((5,), rank, dtype=torch.uint8))          # uniform sizes: all_gather_into_tensor path
    ok = ok and [int(t[0]) for t in u] == list(range(world))
    q.put((rank, b, e, ok))
    dist.destroy_process_group()


def test_scatter_decode_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_scatter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 37
    assert res[0][3] and res[1][3]


def test_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    (r0, a0, b0, c0, t0), (r1, a1, b1, c1, t1) = res
    assert a0 == 0 and b0 == a1 and b1 == 1000
    assert c0 == c1 == [b0 - a0, b1 - a1] and sum(c0) == 1000
    assert t0 == t1 == 1.5
