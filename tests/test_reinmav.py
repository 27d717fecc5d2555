"""reinmav-v0: oracle pinned to the reference's ReinmavEnv (tests/golden/reinmav-v0.npz,
made by oracle/make_golden_reinmav.py from /root/reference), CUDA path checked against both."""
import numpy as np
import pytest

from oracle.reinmav64 import INIT_STATE, ReinmavBatch, substep_times
from tests.util import GOLDEN


@pytest.fixture(scope="module")
def golden():
    import os
    return np.load(os.path.join(GOLDEN, "reinmav-v0.npz"))


def test_substep_grid_matches_numpy_arange():
    t = 0.0
    lengths = []
    for _ in range(400):
        ref = np.arange(t, t + 1 / 100, 1 / 5000)
        mine = substep_times(t)
        assert len(mine) == len(ref) and np.array_equal(np.array(mine), ref)
        lengths.append(len(mine))
        t = t + 1 / 100
    assert set(lengths) == {50, 51}                      # the float arange length flips


def test_oracle_follows_the_reference(golden):
    T, E, _ = golden["traj"].shape
    b = ReinmavBatch(E, state=golden["init"])
    for t in range(T):
        s, r, d = b.step()
        np.testing.assert_allclose(s, golden["traj"][t], rtol=0, atol=1e-11)
        assert (r == 90.0).all() and d.all()
    assert np.array_equal(golden["init"][0], INIT_STATE)
    assert 0.15 < golden["traj"][-1, 0, 0] < 0.18         # env 0 tracks the quintic: x(1.2 s) = 0.163


@pytest.mark.gpu
def test_cuda_path_matches_reference_and_oracle(golden):
    import torch
    import gym_reinmav_b200 as rmv
    T, E, _ = golden["traj"].shape
    b = rmv.QuadBatch("reinmav-v0", E)
    assert (b.state_dim, b.action_dim) == (13, 0)
    assert np.array_equal(b.get_state(), np.tile(INIT_STATE.astype(np.float32), (E, 1)))
    b.set_state(golden["init"].astype(np.float32))
    o = ReinmavBatch(E, state=golden["init"])
    rew = torch.empty(E, device="cuda:0")
    done = torch.empty(E, dtype=torch.uint8, device="cuda:0")
    obs = torch.empty((E, 13), device="cuda:0")
    for t in range(T):
        b.step(None, rew, done, obs=obs)
        s_ref, _, _ = o.step(round_to_float32=True)
        got = obs.cpu().numpy().astype(np.float64)
        # float64 substeps, float32 state between steps: 1e-5 against the reference, 2e-6 against the
        # oracle with the same storage rule
        assert np.abs(got - golden["traj"][t]).max() <= 1e-5, t
        assert np.abs(got - s_ref).max() <= 2e-6, t
        assert (rew.cpu().numpy() == 90.0).all() and (done.cpu().numpy() == 1).all()
    assert abs(b.sim_time - T * 0.01) < 1e-12
    b.close()


@pytest.mark.gpu
def test_reference_test_loop_400_steps():
    """test/test_reinmav.py:12-25: 400 steps of reinmav-v0 (4 s): x, y, z and yaw reach the reference 1.0."""
    import gym_reinmav_b200 as rmv
    env = rmv.make("reinmav-v0")
    assert env.reset().shape == (13,)
    o = ReinmavBatch(1)
    worst = 0.0
    for i in range(400):
        s, r, d, info = env.step()
        so, _, _ = o.step(round_to_float32=True)
        worst = max(worst, float(np.abs(s - so[0]).max()))
        assert r == 90.0 and d is True and info == {}
    assert worst < 1e-5
    assert np.abs(s[:3] - 1.0).max() < 0.02 and abs(env.t - 4.0) < 1e-9
    before = env.state.copy()
    assert np.array_equal(env.reset(), before)            # reset() returns the current state (reinmav_env.py:348-351)
    with pytest.raises(TypeError):
        env.control()
    env.close()


@pytest.mark.gpu
def test_many_envs_and_step_many():
    import torch
    import gym_reinmav_b200 as rmv
    n = 5000
    rng = np.random.RandomState(0)
    init = np.tile(INIT_STATE, (n, 1)) + rng.uniform(-1, 1, (n, 13)) * np.array([0.02] * 3 + [0.05] * 3 + [0.01] * 4 + [0.1] * 3)
    init = init.astype(np.float32)
    a, b = rmv.QuadBatch("reinmav-v0", n), rmv.QuadBatch("reinmav-v0", n)
    a.set_state(init)
    b.set_state(init)
    rew = torch.empty((10, n), device="cuda:0")
    done = torch.empty((10, n), dtype=torch.uint8, device="cuda:0")
    a.step_many(10, None, rew, done, out_ld=n)
    for t in range(10):
        b.step(None, rew[0], done[0])
    assert np.array_equal(a.get_state(), b.get_state())
    o = ReinmavBatch(n, state=init.astype(np.float64))
    for t in range(10):
        s_ref, _, _ = o.step(round_to_float32=True)
    assert np.abs(a.get_state() - s_ref).max() < 2e-6
    a.close()
    b.close()


@pytest.mark.gpu
def test_every_attitude_octant_against_the_oracle():
    """Attitudes over all of SO(3): both atan2 of RotToRPY see every octant and the pi/8 split, asin both of its
    ranges (|R12| above 1/2 takes the library path of rmv_f64.cuh).  Near gimbal lock the reference divides by
    cos(phi) ~ 0 and amplifies its own rounding, so those envs are held to the looser bound only."""
    import torch
    import gym_reinmav_b200 as rmv
    n = 20000
    rng = np.random.RandomState(5)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    init = np.concatenate([rng.uniform(-0.5, 0.5, (n, 3)), rng.uniform(-1, 1, (n, 3)), q, rng.uniform(-2, 2, (n, 3))], axis=1)
    init = init.astype(np.float32)
    b = rmv.QuadBatch("reinmav-v0", n)
    b.set_state(init)
    o = ReinmavBatch(n, state=init.astype(np.float64))
    qw, qx, qy, qz = init[:, 6].astype(np.float64), init[:, 7].astype(np.float64), init[:, 8].astype(np.float64), init[:, 9].astype(np.float64)
    r12 = 2 * (qy * qz - qw * qx) / (qw * qw + qx * qx + qy * qy + qz * qz)
    assert (np.abs(r12) > 0.5).sum() > n // 4 and (np.abs(r12) <= 0.5).sum() > n // 4
    rew = torch.empty(n, device="cuda:0")
    done = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for t in range(3):
        b.step(None, rew, done)
        s_ref, _, _ = o.step(round_to_float32=True)
    diff = np.abs(b.get_state().astype(np.float64) - s_ref).max(axis=1)
    scale = np.maximum(1.0, np.abs(s_ref).max(axis=1))                 # body rates reach tens of rad/s under saturated motors
    regular = np.abs(r12) < 0.99
    assert (diff[regular] / scale[regular]).max() <= 2e-6
    assert (diff / scale).max() <= 1e-4
    b.close()


@pytest.mark.gpu
def test_history_for_plot_state_follows_the_reference_bookkeeping(golden):
    """The reference's own test ends with ``env.plot_state()`` (test/test_reinmav.py:24), which draws the history
    ``step()`` keeps (reinmav_env.py:119-124): cum_t, cum_state = stateToQd(state), cum_desired_state = trj_gen(t + dt)."""
    import gym_reinmav_b200 as rmv
    from oracle.reinmav64 import quat_to_matrix, trajectory
    env = rmv.make("reinmav-v0")
    assert env.cum_t == [0.0] and env.cum_state.shape == (12,) and env.cum_state[:6].tolist() == [0.0] * 6
    states = [env.step()[0] for _ in range(25)]
    assert np.allclose(env.cum_t, np.arange(26) * 0.01, atol=1e-12)
    assert env.cum_state.shape == (26, 12) and env.cum_desired_state.shape == (26, 11)
    for k, s in enumerate(states, start=1):
        qd = env.cum_state[k]
        assert np.array_equal(qd[:6], s[:6]) and np.array_equal(qd[9:], s[10:13])
        R = quat_to_matrix(s[6], s[7], s[8], s[9])
        phi = np.arcsin(R[1][2])
        want = (phi, np.arctan2(-R[0][2] / np.cos(phi), R[2][2] / np.cos(phi)), np.arctan2(-R[1][0] / np.cos(phi), R[1][1] / np.cos(phi)))
        assert np.allclose(qd[6:9], want, rtol=0, atol=1e-12)
        pos, vel, acc = trajectory(k * 0.01)
        assert np.allclose(env.cum_desired_state[k], [pos] * 3 + [vel] * 3 + [acc] * 3 + [pos, vel], rtol=0, atol=1e-12)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            env.plot_state()
    env.close()


@pytest.mark.reference
def test_history_helpers_equal_the_reference_class():
    """stateToQd / trj_gen of the N = 1 view against the reference's own ReinmavEnv (run from /root/reference under the
    shim), on random states and over the history its step() keeps."""
    from oracle.make_golden_reinmav import load_reinmav_class
    from gym_reinmav_b200.envs.native import ReinmavEnv
    ref = load_reinmav_class()()
    rng = np.random.RandomState(11)
    for _ in range(200):
        s = rng.normal(size=13)
        s[6:10] /= np.linalg.norm(s[6:10]) * rng.uniform(0.8, 1.2)
        np.testing.assert_allclose(ReinmavEnv.stateToQd(s), ref.stateToQd(s), rtol=0, atol=1e-12)
    for t in (-1.0, 0.0, 0.37, 1.3, 3.999, 4.0, 7.0):
        np.testing.assert_allclose(ReinmavEnv.trj_gen(t), ref.trj_gen(t), rtol=0, atol=1e-15)
    for _ in range(3):
        ref.step()
    assert len(ref.cum_t) == 4 and np.asarray(ref.cum_state).shape == (4, 12) and np.asarray(ref.cum_desired_state).shape == (4, 11)
    np.testing.assert_allclose(ref.cum_desired_state[3], ReinmavEnv.trj_gen(0.03), rtol=0, atol=1e-15)
    np.testing.assert_allclose(ref.cum_state[3], ReinmavEnv.stateToQd(np.asarray(ref.state, dtype=float).reshape(-1)), rtol=0, atol=1e-12)
