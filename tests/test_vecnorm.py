"""VecNormalize: the restated baselines algorithm against its own algebra (CPU), and the device
implementation against the restatement (GPU)."""
import numpy as np
import pytest

from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle
from oracle.vecnorm import RunningMeanStd, VecNormalizeOracle


def test_running_mean_std_merges_like_concatenation():
    """update_mean_var_count_from_moments (Chan et al.): feeding batches one by one gives the mean / population
    variance of everything seen so far (up to the 1e-4 prior count of baselines' RunningMeanStd)."""
    rng = np.random.RandomState(0)
    rms = RunningMeanStd(shape=(5,))
    seen = []
    for n in (1000, 17, 4096, 1):
        x = rng.normal(loc=rng.uniform(-3, 3, 5), scale=rng.uniform(0.1, 4, 5), size=(n, 5))
        rms.update(x)
        seen.append(x)
    allx = np.concatenate(seen)
    np.testing.assert_allclose(rms.mean, allx.mean(0), rtol=0, atol=1e-6)       # prior: mean 0 with weight 1e-4 / 5114
    np.testing.assert_allclose(rms.var, allx.var(0), rtol=1e-6)
    assert abs(rms.count - (len(allx) + 1e-4)) < 1e-9


def test_vecnormalize_oracle_protocol():
    o = VecNormalizeOracle(4, 3)
    obs = o.reset(np.ones((4, 3)))
    assert obs.shape == (4, 3) and np.abs(obs).max() <= 10 and o.ret.sum() == 0
    obs2, rew = o.step(np.arange(12.0).reshape(4, 3), np.array([1.0, -2.0, 3.0, 0.5]), np.array([False, True, False, False]))
    assert o.ret[1] == 0 and o.ret[0] == 1.0 and np.abs(rew).max() <= 10
    o.training = False
    before = (o.ob_rms.mean.copy(), o.ret_rms.var)
    o.step(np.zeros((4, 3)), np.ones(4), np.zeros(4, bool))
    assert np.array_equal(before[0], o.ob_rms.mean) and before[1] == o.ret_rms.var


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["quadrotor2d-v0", "quadrotor3d-v0", "quadrotor3d-slungload-v0"])
def test_device_vecnormalize_matches_the_restated_baselines_algorithm(env_id):
    """rmv_norm_* (two kernels per step, float64 statistics in HBM) against oracle/vecnorm.py on the raw outputs of the
    same envs: normalised observations / rewards within 2e-6 (float32 outputs of float64 arithmetic), running
    statistics within 1e-9 relative, discounted returns cleared on done, frozen statistics in evaluation mode,
    checkpoint / restore of the statistics."""
    import torch
    import gym_reinmav_b200 as rmv
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, seed, T = 128 * 23 + 9, 6, 150
    venv = rmv.VecNormalize(rmv.QuadVecEnv(env_id, n, seed=seed), gamma=0.99)
    raw = BatchOracle(kind, n, seed=seed, dtype=np.float32)
    o = VecNormalizeOracle(n, S, gamma=0.99)
    got = venv.reset_tensor().cpu().numpy()
    want = o.reset(raw.reset().astype(np.float64))
    assert np.abs(got - want).max() <= 2e-6 * 10
    rng = np.random.RandomState(2)
    n_done = 0
    for t in range(T):
        if t == T - 20:
            venv.training = o.training = False                      # evaluation: statistics frozen
        a = rng.uniform(lo, hi, size=(n, A)).astype(np.float32)
        obs, rew, done = venv.step_tensor(torch.as_tensor(a, device="cuda:0"))
        r_obs, r_rew, r_done = raw.step(a)
        w_obs, w_rew = o.step(r_obs.astype(np.float64), r_rew.astype(np.float64), r_done)
        assert np.array_equal(done.cpu().numpy().astype(bool), r_done)
        assert np.abs(obs.cpu().numpy() - w_obs).max() <= 2e-5, t             # |normalised obs| <= 10, float32 output
        assert np.abs(rew.cpu().numpy() - w_rew).max() <= 2e-5, t
        n_done += int(r_done.sum())
    st = venv.statistics()
    np.testing.assert_allclose(st["ob_mean"], o.ob_rms.mean, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st["ob_var"], o.ob_rms.var, rtol=1e-9)
    np.testing.assert_allclose([st["ret_mean"], st["ret_var"]], [o.ret_rms.mean, o.ret_rms.var], rtol=1e-9)
    assert abs(st["ob_count"] - o.ob_rms.count) < 1e-6 and abs(st["ret_count"] - o.ret_rms.count) < 1e-6 and n_done > 20
    # host protocol + checkpoint / restore
    other = rmv.VecNormalize(rmv.QuadVecEnv(env_id, n, seed=seed))
    other.load_statistics(st)
    st2 = other.statistics()
    assert all(np.array_equal(np.asarray(st[k]), np.asarray(st2[k])) for k in st)
    h_obs, h_rew, h_done, infos = venv.step(rng.uniform(lo, hi, size=(n, A)).astype(np.float32))
    assert h_obs.shape == (n, S) and h_obs.dtype == np.float32 and h_rew.shape == (n,) and h_done.dtype == bool and len(infos) == n
    venv.close(); other.close()


@pytest.mark.gpu
def test_norm_step_with_unaligned_matrices_takes_the_scalar_walk():
    """rmv_norm_step walks the observation matrix four floats at a time when both matrices are 16-byte aligned and one at
    a time otherwise: both walks against the oracle, on a batch whose n * S is not a multiple of 4."""
    import ctypes
    import torch
    from gym_reinmav_b200._capi import check, lib
    n, S = 1237, 9
    results = []
    for offset in (0, 1):                                       # offset 1 float: 4-byte aligned only
        h = ctypes.c_void_p()
        check(lib().rmv_norm_create(ctypes.byref(h), n, S, 0, 1, 1, 10.0, 10.0, 0.99, 1e-8))
        o = VecNormalizeOracle(n, S, gamma=0.99)
        buf_in = torch.zeros(n * S + 4, dtype=torch.float32, device="cuda:0")
        buf_out = torch.zeros(n * S + 4, dtype=torch.float32, device="cuda:0")
        rew_out = torch.zeros(n, dtype=torch.float32, device="cuda:0")
        obs_in, obs_out = buf_in[offset:offset + n * S].view(n, S), buf_out[offset:offset + n * S].view(n, S)
        assert obs_in.data_ptr() % 16 == 4 * offset
        rs = np.random.RandomState(9)
        for t in range(6):
            x = rs.normal(2.0, 3.0, size=(n, S)).astype(np.float32)
            r = rs.normal(-1.0, 2.0, size=n).astype(np.float32)
            d = rs.uniform(size=n) < 0.1
            obs_in.copy_(torch.as_tensor(x))
            rew, done = torch.as_tensor(r, device="cuda:0"), torch.as_tensor(d.astype(np.uint8), device="cuda:0")
            check(lib().rmv_norm_step(h, obs_in.data_ptr(), rew.data_ptr(), done.data_ptr(), obs_out.data_ptr(), rew_out.data_ptr(), 1, None))
            w_obs, w_rew = o.step(x.astype(np.float64), r.astype(np.float64), d)
            assert np.abs(obs_out.cpu().numpy() - w_obs).max() <= 2e-5 and np.abs(rew_out.cpu().numpy() - w_rew).max() <= 2e-5
        assert float(buf_out[:offset].abs().sum()) == 0.0 and float(buf_out[offset + n * S:].abs().sum()) == 0.0   # nothing outside
        results.append(obs_out.cpu().numpy().copy())
        check(lib().rmv_norm_destroy(h))
    assert np.abs(results[0] - results[1]).max() <= 1e-6


@pytest.mark.gpu
def test_host_protocol_equals_the_device_path():
    """VecNormalize.reset() / step() (NumPy in, NumPy out through the env's page-locked buffers, asynchronous copies, one
    synchronisation) returns the bits of reset_tensor() / step_tensor(); infos carry Monitor's record for finished envs."""
    import torch
    import gym_reinmav_b200 as rmv
    n, seed = 128 * 9 + 40, 3
    host = rmv.VecNormalize(rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed))
    dev = rmv.VecNormalize(rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed))
    o_h, o_d = host.reset(), dev.reset_tensor().cpu().numpy()
    assert o_h.dtype == np.float32 and np.array_equal(o_h.view(np.int32), o_d.view(np.int32))
    rng = np.random.RandomState(1)
    finished = 0
    for t in range(60):
        a = rng.uniform(0, 25, size=(n, 4)).astype(np.float32)
        oh, rh, dh, infos = host.step(a)
        od, rd, dd = dev.step_tensor(torch.as_tensor(a, device="cuda:0"))
        assert np.array_equal(oh.view(np.int32), od.cpu().numpy().view(np.int32)), t
        assert np.array_equal(rh.view(np.int32), rd.cpu().numpy().view(np.int32)) and np.array_equal(dh, dd.cpu().numpy().astype(bool))
        for i in np.flatnonzero(dh)[:5]:
            assert set(infos[int(i)]["episode"]) == {"r", "l", "t"}
        finished += int(dh.sum())
    assert finished > 10
    host.close(); dev.close()
