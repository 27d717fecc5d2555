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


def test_key_schedule_generator_of_the_redraw_pass_matches(host_twin):
    """The step kernels' re-draw pass runs Philox with a host-precomputed key schedule (rmv_math.cuh
    philox4x32_10_ks): same function as the plain generator, on the KATs and on random counters / seeds."""
    host_twin.twin_philox_ks.argtypes = [ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
    for ctr, key, want in KATS:
        o = (ctypes.c_uint * 4)()
        host_twin.twin_philox_ks(key[0] | (key[1] << 32), (ctypes.c_uint * 4)(*ctr), o)
        assert tuple(o) == want
    rng = np.random.RandomState(5)
    for _ in range(2000):
        ctr = [int(x) for x in rng.randint(0, 2 ** 32, size=4, dtype=np.uint64)]
        seed = int(rng.randint(0, 2 ** 63, dtype=np.uint64)) * 2 + int(rng.randint(0, 2))
        o = (ctypes.c_uint * 4)()
        host_twin.twin_philox_ks(seed, (ctypes.c_uint * 4)(*ctr), o)
        want = philox.philox4x32_10(*ctr, seed & 0xFFFFFFFF, seed >> 32)
        assert tuple(o) == tuple(int(v) for v in want)


def test_one_fma_uniform_equals_the_two_step_form_for_every_input(host_twin):
    """uniform_pm1(bits) = fma(bits >> 8, 2^-23, -1) against uniform_from_bits(bits, -1, 2) for all 2^24 values of
    the 24 bits that matter (and a few patterns in the ignored low byte): identical bits, all multiples of 2^-23."""
    n = 1 << 24
    bits = (np.arange(n, dtype=np.uint64) << 8).astype(np.uint32)
    bits[::7] |= 0xA5                                      # the low byte is discarded
    fast = np.empty(n, dtype=np.float32)
    plain = np.empty(n, dtype=np.float32)
    host_twin.twin_uniform_pm1.argtypes = [ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    host_twin.twin_uniform_pm1(n, bits.ctypes.data, fast.ctypes.data, plain.ctypes.data)
    assert np.array_equal(fast.view(np.uint32), plain.view(np.uint32))
    assert fast[0] == -1.0 and fast[-1] == np.float32(1.0 - 2.0 ** -23)
    want = (-1.0 + np.arange(n, dtype=np.float64) * 2.0 ** -23).astype(np.float32)
    assert np.array_equal(fast, want)
