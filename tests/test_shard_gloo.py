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
