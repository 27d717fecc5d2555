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
                ("later_done_reward", ctypes.c_float), ("reward_scale", ctypes.c_float)]


def _cparams(d):
    return CParams(d["mass"], d["load_mass"], d["dt"], d["gravity"], d["tether_len"], d["pos_done"], d["vel_done"],
                   d["thrust_scale"], d["thrust_clamp"], d["first_done_reward"], d["later_done_reward"],
                   d.get("reward_scale", 1.0))


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


def test_atan2_bitwise_and_accurate(host_twin):
    """The pinned float32 atan2 of the planar controllers (csrc/rmv_math.cuh atan2_pinned == oracle/atan32.py) and its
    distance from the float64 libm the reference calls (quadrotor2d.py:128)."""
    from oracle.atan32 import atan2_32
    rng = np.random.RandomState(1)
    y = np.concatenate([rng.uniform(-60, 60, 400000), rng.normal(size=100000) * 1e-3, rng.normal(size=100000),
                        [0.0, -0.0, 0.0, 1.0, -1.0, 3.0, 1e-30, 1e30, -2.0, 0.0]]).astype(np.float32)
    x = np.concatenate([rng.uniform(-60, 60, 400000), rng.normal(size=100000), rng.normal(size=100000) * 1e-3,
                        [1.0, -1.0, 0.0, 0.0, 0.0, -0.0, 1e30, 1e-30, -2.0, -0.0]]).astype(np.float32)
    out = np.empty_like(x)
    host_twin.twin_atan2(ctypes.c_long(len(x)), _p(y), _p(x), _p(out))
    want = atan2_32(y, x)
    assert same_bits(out, want)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    ok = ~((x == 0) & (y == 0))                     # libm returns +-0 / +-pi by the signs of the zeros; the controllers never get there
    assert np.abs(out[ok] - ref[ok]).max() < 3e-7


@pytest.mark.parametrize("env_id", list(KIND_OF_ID))
def test_control_law_bitwise_and_within_1e5_of_float64(host_twin, env_id):
    """control_law<KIND> (the arithmetic of k_control and of the fused control+step kernels) equals the float32
    model's control() bit for bit, and the float32 model is within 1e-5 * max(1, |a|) of the float64 restatement of
    the reference's controllers — on reset-like and wide states, default and moved set-points."""
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    n = 200000
    s0 = wide_states(kind, n, np.random.RandomState(3))
    for ref_pos in (None, [0.3, -0.2] if A == 2 else [0.3, -0.2, 1.5]):
        o32 = BatchOracle(kind, n, dtype=np.float32)
        o32.state = s0.T.copy()
        o64 = BatchOracle(kind, n, dtype=np.float64)
        o64.state = s0.T.astype(np.float64)
        want32, want64 = o32.control(ref_pos=ref_pos), o64.control(ref_pos=ref_pos)
        rp = list(ref_pos if ref_pos is not None else BatchOracle.REF_POS[kind]) + [0.0] * 3
        ctrl = np.array(rp[:3] + [0.0, 0.0, 0.0, 1.0, 9.8], dtype=np.float32)
        got = np.empty((n, A), np.float32)
        assert host_twin.twin_control(kind, _p(ctrl), ctypes.c_long(n), _p(s0), _p(got)) == 0
        finite = np.isfinite(want64).all(axis=1)
        assert finite.mean() > 0.99
        assert same_bits(got[finite], want32[finite].astype(np.float32))
        err = np.abs(got.astype(np.float64) - want64) / np.maximum(1.0, np.abs(want64))
        # states the reference's own loop visits (reset draws U[-1,1)^S and what the controller makes of them): 1e-5
        # (planar kinds: 1.5e-5 — the rate command is -10 * (theta - att), three float32 roundings of an angle times the gain)
        home = finite & (np.abs(s0).max(axis=1) <= 1.0)
        assert home.sum() > 0.25 * n and err[home].max() <= (1e-5 if A == 4 else 1.5e-5), err[home].max()
        # wide states (3x the reset range): the commands are gains of 6.7-10 on differences of float32 numbers of
        # magnitude up to 5 (spacing 4.8e-7) and sums of products of magnitude 40, and the 3D attitude construction is
        # ill-conditioned where the desired acceleration is almost parallel to the y axis (|xb| -> 0): a handful of
        # states in 2e5 land above 1e-5 (worst 3.8e-5).  That is float32 itself, not the pinned atan2 (<= 2.8e-7).
        assert err[finite].max() <= 1e-4, err[finite].max()
        assert np.quantile(err[finite], 0.9999) <= 5e-6 and (err[finite] > 1e-5).sum() <= 4


def test_float64_asin_atan2_of_reinmav_match_libm(host_twin):
    """rmv_f64.cuh (the Euler angles of k_reinmav_step, reinmav_env.py:341-346): a few 1e-16 from NumPy's libm over
    every octant, the pi/8 split, the |x| = 1/2 switch of asin, signed zeros, infinities and NaN."""
    rng = np.random.RandomState(3)
    n = 400000
    ang = rng.uniform(-np.pi, np.pi, n)
    rad = 10.0 ** rng.uniform(-6, 6, n)
    y, x = rad * np.sin(ang), rad * np.cos(ang)
    edge = [0.0, -0.0, 1.0, -1.0, np.sqrt(2) - 1, 1 - np.sqrt(2), 1e-300, -1e-300, 5e-324, 1e300, np.inf, -np.inf, np.nan]
    ey, ex = np.meshgrid(edge, edge)
    y, x = np.concatenate([y, ey.ravel(), np.tan(np.pi / 8) * np.abs(x[:1000])]), np.concatenate([x, ex.ravel(), np.abs(x[:1000])])
    out = np.empty_like(y)
    host_twin.twin_atan2_f64(ctypes.c_long(len(y)), _p(y), _p(x), _p(out))
    with np.errstate(all="ignore"):
        ref = np.arctan2(y, x)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(np.signbit(out[ok]), np.signbit(ref[ok]))
    assert np.abs(out[ok] - ref[ok]).max() <= 4.5e-16            # one ulp at pi
    rel = np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert rel.max() <= 4 * np.finfo(np.float64).eps

    v = np.concatenate([rng.uniform(-1, 1, n), rng.uniform(-0.5, 0.5, n), 10.0 ** rng.uniform(-12, 0, 1000),
                        [0.0, -0.0, 0.5, -0.5, np.nextafter(0.5, 1), 1.0, -1.0, 1.0000001, np.nan, 1e-310]])
    out = np.empty_like(v)
    host_twin.twin_asin_f64(ctypes.c_long(len(v)), _p(v), _p(out))
    with np.errstate(all="ignore"):
        ref = np.arcsin(v)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(np.signbit(out[ok]), np.signbit(ref[ok]))
    rel = np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert rel.max() <= 4 * np.finfo(np.float64).eps
