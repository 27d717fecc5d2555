"""Multi-GPU plumbing: envs shard embarrassingly (independent batches, no exchange on the step path).

One process per GPU (``torchrun``).  Rank r owns the global env ids ``[lo, hi)`` returned by
``shard_range``; Philox reset draws are keyed by the GLOBAL id, so a job's results do not depend on the
number of GPUs.  The only collective is optional: summing four episode-statistics scalars for logging
(NCCL when the tensors are on GPUs, gloo in the CPU tests).
"""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> (0, 1, 0))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_total, rank, world_size, tile=128):
    """Contiguous slice of ``n_total`` global env ids for ``rank``.  Slices are multiples of the
    128-env state tile (except possibly the last), differ by at most one tile, and cover [0, n_total)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    tiles = (n_total + tile - 1) // tile
    base, extra = divmod(tiles, world_size)
    lo_t = rank * base + min(rank, extra)
    hi_t = lo_t + base + (1 if rank < extra else 0)
    return min(lo_t * tile, n_total), min(hi_t * tile, n_total)


def make_sharded_vec_env(env_id, n_total, seed=0, **kwargs):
    """The QuadVecEnv of this rank's shard of a job with ``n_total`` envs (device = LOCAL_RANK)."""
    from .vec_env import QuadVecEnv
    rank, size, local = world()
    lo, hi = shard_range(n_total, rank, size)
    if hi <= lo:
        raise ValueError("rank %d got an empty shard (n_total=%d, world=%d)" % (rank, n_total, size))
    return QuadVecEnv(env_id, hi - lo, device=local, seed=seed, env_id_offset=lo, **kwargs)


def all_reduce_stats(stats, device=None):
    """Sum a ``{'episodes','return_sum','length_sum','env_steps'}`` dict over all ranks
    (``torch.distributed`` must be initialised; a single process returns its own numbers) and add the means."""
    import torch
    import torch.distributed as dist
    keys = ("episodes", "return_sum", "length_sum", "env_steps")
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = dict(zip(keys, t.tolist()))
    n = max(out["episodes"], 1.0)
    out["mean_return"], out["mean_length"] = out["return_sum"] / n, out["length_sum"] / n
    return out


def all_reduce_episode_stats(venv):
    """Job-wide episode statistics of sharded QuadVecEnvs (needs ``monitor=True``)."""
    return all_reduce_stats(venv.episode_stats(), device="cuda:%d" % venv.device)
