"""Philox4x32-10: oracle pinned to the Random123 known-answer vectors,
host build of the kernel's generator pinned to the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import philox

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_oracle_random123_kat(ctr, key, want):
    got = philox.philox4x32_10(*ctr, *key)
    assert tuple(int(v) for v in got) == want


def test_uniform_is_exact_and_in_range():
    u = philox.uniform_f32(7, np.arange(1000), np.ones(1000), 2, 10, -1.0, 1.0)
    assert u.dtype == np.float32 and u.shape == (10, 1000)
    assert (u >= -1).all() and (u < 1).all()
    # multiples of 2^-23: identical in float32 and float64
    assert np.array_equal(u.astype(np.float64) * 2 ** 23, np.round(u.astype(np.float64) * 2 ** 23))
    assert abs(u.mean()) < 0.02 and abs(u.std() - 1 / np.sqrt(3)) < 0.01


def test_streams_and_indices_differ():
    a = philox.uniform_f32(1, [5], [1], 2, 4, -1, 1)
    assert not np.array_equal(a, philox.uniform_f32(1, [5], [2], 2, 4, -1, 1))
    assert not np.array_equal(a, philox.uniform_f32(1, [6], [1], 2, 4, -1, 1))
    assert not np.array_equal(a, philox.uniform_f32(2, [5], [1], 2, 4, -1, 1))
    assert not np.array_equal(a, philox.uniform_f32(1, [5], [1], 3, 4, -1, 1))


def test_kernel_generator_host_build_matches(host_twin):
    for ctr, key, want in KATS:
        c = (ctypes.c_uint * 4)(*ctr)
        k = (ctypes.c_uint * 2)(*key)
        o = (ctypes.c_uint * 4)()
        host_twin.twin_philox(c, k, o)
        assert tuple(o) == want
