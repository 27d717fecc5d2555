"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11; Random123) in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Independent restatement of the
published algorithm; pinned by the Random123 known-answer vectors listed in
SURVEY.md Appendix D (tests/test_philox.py).

It replaces the reference's MT19937 stream (``gym.utils.seeding.np_random`` at
/root/reference/gym_reinmav/envs/native/quadrotor3d.py:77-79, consumed by
``reset()`` at :182-185): a per-env sequential generator cannot be reproduced
lane-parallel at N = 2^20, so both the CUDA path and this oracle draw
``U[-1, 1)`` from the same counter-based stream instead (documented deviation,
DESIGN.md "RNG").

Counter / key convention shared with the CUDA path (csrc/rmv_math.cuh):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (env_id & 0xffffffff, env_id >> 32, index, (stream << 8) | block)

``env_id`` is the GLOBAL env index (shard offset + local index), ``index`` is
the episode number for reset draws or the step number for random-policy action
draws, ``stream`` separates the consumers, ``block`` selects the group of four
32-bit words.  Draw ``j`` of a consumer is word ``j % 4`` of block ``j // 4``.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)
SH32 = np.uint64(32)

# stream ids (must match csrc/rmv_math.cuh)
STREAM_RESET_BASE = 0      # + env kind (0..3)
STREAM_ACTION = 16


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32 on arrays (or scalars) of uint32 values.

    All arguments broadcast together; returns four uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = (np.asarray(v).astype(np.uint64) & MASK
                              for v in (c0, c1, c2, c3, k0, k1))
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    for _ in range(10):
        p0 = M0 * c0            # 32x32 -> 64 bit product, fits in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> SH32, p0 & MASK
        hi1, lo1 = p1 >> SH32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def u32_to_unit_f32(x):
    """Top 24 bits -> float32 in [0, 1) (exact: a multiple of 2^-24)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def uniform_f32(seed, env_ids, index, stream, n_draws, lo, hi):
    """``n_draws`` float32 uniforms per env: shape ``[n_draws, len(env_ids)]``.

    value = lo + (hi - lo) * u, every operation rounded to float32 in that
    order — the order the CUDA path uses.  For (lo, hi) = (-1, 1) every value
    is exact (a multiple of 2^-23), so the float64 oracle sees the same numbers.
    """
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    index = np.asarray(index, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    lo32, hi32 = np.float32(lo), np.float32(hi)
    width = np.float32(hi32 - lo32)
    out = np.empty((n_draws, env_ids.shape[0]), dtype=np.float32)
    for block in range((n_draws + 3) // 4):
        words = philox4x32_10(env_ids & MASK, env_ids >> SH32, index,
                              np.uint64((stream << 8) | block), k0, k1)
        for lane in range(4):
            j = 4 * block + lane
            if j < n_draws:
                out[j] = lo32 + width * u32_to_unit_f32(words[lane])
    return out
