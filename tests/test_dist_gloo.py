"""Host-side multi-GPU logic on CPU: shard arithmetic, and the statistics all-reduce over a
world_size-2 gloo group (the GPU job uses the same code over NCCL)."""
import os
import socket

import pytest

from gym_reinmav_b200 import dist as rdist


@pytest.mark.parametrize("n,world", [(1 << 20, 1), (1 << 20, 8), (1 << 23, 8), (1000, 3), (129, 2), (128, 4), (5, 2)])
def test_shards_cover_without_overlap(n, world):
    ranges = [rdist.shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (lo, hi), (lo2, hi2) in zip(ranges, ranges[1:]):
        assert hi == lo2 and lo <= hi
    sizes = [hi - lo for lo, hi in ranges]
    assert all(lo % 128 == 0 for lo, hi in ranges if hi > lo)
    assert max(sizes) - min(sizes) < 256          # one tile, plus the ragged last tile
    with pytest.raises(ValueError):
        rdist.shard_range(n, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert rdist.world() == (rank, world, rank)
        lo, hi = rdist.shard_range(1000, rank, world)
        mine = {"episodes": 10.0 * (rank + 1), "return_sum": -100.0 * (rank + 1), "length_sum": 700.0 + rank,
                "env_steps": float(hi - lo) * 50}
        total = rdist.all_reduce_stats(mine)
        out.put((rank, lo, hi, total))
    finally:
        dist.destroy_process_group()


def test_stats_all_reduce_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0), (r1, lo1, hi1, t1) = got
    assert (lo0, hi1) == (0, 1000) and hi0 == lo1
    assert t0 == t1
    assert t0["episodes"] == 30.0 and t0["return_sum"] == -300.0 and t0["length_sum"] == 1401.0
    assert t0["env_steps"] == 1000 * 50 and t0["mean_return"] == -10.0


def test_single_process_stats_pass_through():
    s = rdist.all_reduce_stats({"episodes": 4, "return_sum": -8.0, "length_sum": 40, "env_steps": 100})
    assert s["mean_return"] == -2.0 and s["mean_length"] == 10.0
