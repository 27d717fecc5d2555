"""The arithmetic the kernels inline (csrc/rmv_math.cuh), compiled for the host, must equal the
NumPy float32 model (oracle T1) bit for bit.  CPU only — the same comparison runs against the real
kernels in tests/test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest

from oracle import philox
from oracle.batch import ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle, default_params
from oracle.sincos32 import sincos32
from tests.util import random_actions, same_bits, wide_states


class CParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("mass", "load_mass", "dt", "gravity", "tether_len", "pos_done",
                                              "vel_done", "thrust_scale")] + \
               [("thrust_clamp", ctypes.c_int32), ("first_done_reward", ctypes.c_float),
                ("later_done_reward", ctypes.c_float)]


def _cparams(d):
    return CParams(d["mass"], d["load_mass"], d["dt"], d["gravity"], d["tether_len"], d["pos_done"], d["vel_done"],
                   d["thrust_scale"], d["thrust_clamp"], d["first_done_reward"], d["later_done_reward"])


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_sincos_bitwise_and_accurate(host_twin):
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e5, 1e5, 100000), rng.uniform(-1e9, 1e9, 2000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1e38, -3e38, 1e5, 100000.01, np.pi / 2, -np.pi]]
                       ).astype(np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    host_twin.twin_sincos(ctypes.c_long(len(x)), _p(x), _p(s), _p(c))
    s2, c2 = sincos32(x)
    assert same_bits(s, s2) and same_bits(c, c2)
    small = np.abs(x) <= 100
    xd = x[small].astype(np.float64)
    assert np.abs(s[small] - np.sin(xd)).max() < 2.5e-7
    assert np.abs(c[small] - np.cos(xd)).max() < 2.5e-7
    assert np.isnan(s[np.isinf(x)]).all() and np.isnan(c[np.isnan(x)]).all()


@pytest.mark.parametrize("env_id", list(KIND_OF_ID))
def test_reset_draws_bitwise(host_twin, env_id):
    kind = KIND_OF_ID[env_id]
    S = STATE_DIM[kind]
    out = np.empty(S, np.float32)
    for seed, eid, ep in [(0, 0, 1), (1, 5, 2), (2 ** 63 + 5, 2 ** 33 + 7, 12345), (20190312, 999999, 77)]:
        assert host_twin.twin_reset_draws(kind, ctypes.c_ulonglong(seed), ctypes.c_ulonglong(eid),
                                          ctypes.c_uint(ep), _p(out)) == 0
        want = philox.uniform_f32(seed, [eid], [ep], philox.STREAM_RESET_BASE + kind, S, -1.0, 1.0)[:, 0]
        assert same_bits(out, want)


@pytest.mark.parametrize("env_id", list(KIND_OF_ID))
def test_step_arithmetic_bitwise(host_twin, env_id):
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    rng = np.random.RandomState(100 + kind)
    n = 100000
    s0 = wide_states(kind, n, rng)
    if S in (9, 16):          # states exactly as a taut step leaves them: on the switching surface
        o = BatchOracle(kind, n // 4, dtype=np.float32)
        o.state = s0[:n // 4].T.copy()
        s0[:n // 4] = o.step(random_actions(kind, (n // 4,), rng), auto_reset=False)[0]
    a = random_actions(kind, (n,), rng)
    pd = default_params(kind)
    cp = _cparams(pd)
    so, norms = np.empty_like(s0), np.empty((n, 3), np.float32)
    assert host_twin.twin_step(kind, ctypes.byref(cp), ctypes.c_long(n), _p(s0), _p(a), _p(so), _p(norms)) == 0
    o = BatchOracle(kind, n, dtype=np.float32)
    o.state = s0.T.copy()
    obs, rew, done = o.step(a, auto_reset=False)
    nan_both = np.isnan(so) & np.isnan(obs)
    assert ((so.view(np.uint32) == obs.view(np.uint32)) | nan_both).all()
    derived = np.empty(4, np.float32)
    host_twin.twin_derived(ctypes.byref(cp), _p(derived))
    # the kernel thresholds SQUARED norms against the largest y with sqrt(y) <= threshold
    done2 = (norms[:, 0] > derived[0]) | (norms[:, 1] > derived[1])
    assert np.array_equal(done, done2)
    assert 0.02 < done.mean() < 0.6
    rew2 = np.where(done2, np.float32(1.0), -norms[:, 2]).astype(np.float32)
    assert same_bits(rew, rew2)
    if o.tether_margin is not None:
        taut = (o.tether_margin >= 0).mean()
        assert 0.1 < taut < 0.95


def test_squared_threshold_is_equivalent_to_the_norm_comparison(host_twin):
    """sq_threshold(t) = max{y : sqrt_rn(y) <= t}, so  sqrt(y) > t  <=>  y > sq_threshold(t)  for every y."""
    host_twin.twin_sq_threshold.restype = ctypes.c_float
    host_twin.twin_sq_threshold.argtypes = [ctypes.c_float]
    rng = np.random.RandomState(5)
    for t in [3.0, 10.0, 2.0, 0.0, 1e-20, 1e19, 0.1, 2.5, 7.77] + list(rng.uniform(0, 100, 40)):
        t = np.float32(t)
        y = np.float32(host_twin.twin_sq_threshold(ctypes.c_float(float(t))))
        around = y
        for _ in range(4):
            around = np.nextafter(around, np.float32(-np.inf))
        for _ in range(9):
            assert (np.sqrt(around) > t) == (around > y), (t, y, around)
            around = np.nextafter(around, np.float32(np.inf))
    assert np.isnan(host_twin.twin_sq_threshold(ctypes.c_float(float("nan"))))
    assert host_twin.twin_sq_threshold(ctypes.c_float(-1.0)) == -1.0
    assert host_twin.twin_sq_threshold(ctypes.c_float(float("inf"))) == float("inf")
