"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): disjoint, complete sharding of the stream and the
max-over-ranks reduction that bench.py uses for multi-GPU timings."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200", "seqpurge_b200"))
import sharding  # noqa: E402  (imported as a plain module: no CUDA library needed)


def test_shards_are_disjoint_and_complete():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            s = sharding.shard_batches(5, 1000, r, world)
            assert len(s.batches) == 5 and s.first_pair[0] == s.batches[0] * 1000
            seen += list(s.batches)
        assert sorted(seen) == list(range(5 * world))
    with pytest.raises(ValueError):
        sharding.shard_batches(1, 1, 2, 2)
    assert [sharding.round_robin_device(s, 3) for s in range(7)] == [0, 1, 2, 0, 1, 2, 0]
    assert sharding.aggregate_throughput(1000, 4, 2.0) == 2000.0
    assert sharding.reduce_max(3.5) == 3.5


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = sharding.shard_batches(3, 8, rank, world)
    # every rank "processes" its shard: checksum of the pair indices it owns, gathered to prove the cover is exact
    mine = torch.tensor([sum(range(f, f + 8)) for f in s.first_pair], dtype=torch.int64).sum()
    dist.all_reduce(mine)
    elapsed = 1.0 + rank  # the slowest rank decides
    mx = sharding.reduce_max(elapsed, dist)
    out.put((rank, int(mine.item()), mx))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total_pairs = world * 3 * 8
    for rank, checksum, mx in res:
        assert checksum == sum(range(total_pairs))
        assert mx == 2.0
