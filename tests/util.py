"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle  # noqa: F401

ENV_IDS = list(KIND_OF_ID)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(env_id, repair="A"):
    name = env_id + (".repairB" if repair == "B" else "") + ".npz"
    return np.load(os.path.join(GOLDEN, name))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def same_bits(a, b):
    """Bitwise equality of two float arrays (NaN payloads included)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(bits(a), bits(b))


def random_actions(kind, shape_prefix, rng, dtype=np.float32):
    lo, hi = ACTION_BOUNDS[kind]
    return rng.uniform(lo, hi, size=tuple(shape_prefix) + (ACTION_DIM[kind],)).astype(dtype)


def wide_states(kind, n, rng):
    """Reset-like and wide states, plus (slung kinds) loads placed around the tether length."""
    S = STATE_DIM[kind]
    s = rng.uniform(-1, 1, size=(n, S))
    s[n // 2:] *= 3.0
    if S in (9, 16):
        d = 2 if S == 9 else 3
        lp = slice(5, 7) if S == 9 else slice(10, 13)
        L = 0.5 if S == 9 else 1.5
        k = n // 4
        dirs = rng.normal(size=(k, d))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        s[:k, lp] = s[:k, 0:d] + dirs * L * rng.uniform(0.9, 1.2, size=(k, 1))
    return s.astype(np.float32)
