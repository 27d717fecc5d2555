"""The launcher seam: ``make_vec_env`` exactly as /root/reference/gym_reinmav/run.py:89 calls it, and the ppo2
Runner loop (run.py:63-68 -> baselines ``ppo2.Runner.run``; --play loop run.py:192-211) consuming the VecEnv.

CPU part: the call shape (no device needed: QuadVecEnv is replaced by a recorder).  GPU part: the real thing,
checked against the float32 oracle.
"""
import inspect

import numpy as np
import pytest

from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle
from tests.util import same_bits

# baselines.common.cmd_util.make_vec_env (README.md:20,27 pins baselines@90d6677; later revisions append
# env_kwargs / initializer / force_dummy): positional order and defaults a caller may rely on
BASELINES_SIGNATURE = ["env_id", "env_type", "num_env", "seed", "wrapper_kwargs", "env_kwargs", "start_index",
                       "reward_scale", "flatten_dict_observations", "gamestate", "initializer", "force_dummy"]


class _Recorder:
    last = None

    def __init__(self, env_id, num_envs, **kw):
        _Recorder.last = dict(env_id=env_id, num_envs=num_envs, **kw)


@pytest.fixture()
def seam(monkeypatch):
    import gym_reinmav_b200 as rmv
    monkeypatch.setattr(rmv, "QuadVecEnv", _Recorder)
    return rmv


def test_make_vec_env_has_the_baselines_signature(seam):
    names = list(inspect.signature(seam.make_vec_env).parameters)
    assert names[:len(BASELINES_SIGNATURE)] == BASELINES_SIGNATURE
    p = inspect.signature(seam.make_vec_env).parameters
    assert p["reward_scale"].default == 1.0 and p["flatten_dict_observations"].default is True
    assert p["start_index"].default == 0


def test_make_vec_env_accepts_the_reference_call_verbatim(seam):
    """run.py:89: make_vec_env(env_id, env_type, args.num_env or 1, seed, reward_scale=args.reward_scale,
    flatten_dict_observations=flatten_dict_observations) — common_arg_parser defaults: seed None, reward_scale 1.0."""
    env_id, env_type = "quadrotor2d-v0", "native"
    seam.make_vec_env(env_id, env_type, None or 1, None, reward_scale=1.0, flatten_dict_observations=True)
    assert _Recorder.last == dict(env_id=env_id, num_envs=1, seed=0, env_id_offset=0, params=None)
    seam.make_vec_env("quadrotor3d-v0", "native", 4096, 7, reward_scale=0.1, flatten_dict_observations=True)
    assert _Recorder.last["num_envs"] == 4096 and _Recorder.last["seed"] == 7
    assert _Recorder.last["params"] == {"reward_scale": 0.1}
    # start_index -> global env id of env 0; QuadVecEnv's own options pass through; params merge
    seam.make_vec_env("quadrotor3d-v0", "native", 8, 1, start_index=128, reward_scale=2.0, device=3, monitor=False,
                      params={"vel_done": 2.0})
    assert _Recorder.last["env_id_offset"] == 128 and _Recorder.last["device"] == 3 and _Recorder.last["monitor"] is False
    assert _Recorder.last["params"] == {"vel_done": 2.0, "reward_scale": 2.0}
    seam.make_vec_env("quadrotor3d-v0", "native", 8, 1, wrapper_kwargs={}, env_kwargs=None, force_dummy=True)


def test_make_vec_env_refuses_what_it_cannot_honour(seam):
    with pytest.raises(ValueError, match="native"):
        seam.make_vec_env("quadrotor3d-v0", "mujoco", 1, 0)
    with pytest.raises(ValueError, match="wrapper_kwargs"):
        seam.make_vec_env("quadrotor3d-v0", "native", 1, 0, wrapper_kwargs={"frame_stack": True})
    with pytest.raises(ValueError, match="initializer"):
        seam.make_vec_env("quadrotor3d-v0", "native", 1, 0, initializer=lambda **kw: None)


def test_spaces_follow_the_reference_dtypes():
    """quadrotor3d.py:70-71 declares float64 Boxes (dtype=np.float); the other kinds float32."""
    from gym_reinmav_b200 import _capi
    from gym_reinmav_b200.vec_env import make_spaces
    for kind, want in ((_capi.QUAD2D, np.float32), (_capi.QUAD2D_SLUNG, np.float32), (_capi.QUAD3D, np.float64),
                       (_capi.QUAD3D_SLUNG, np.float32)):
        obs, act = make_spaces(kind)
        assert obs.dtype == want and act.dtype == want
        assert obs.contains(np.zeros(obs.shape, np.float32).astype(obs.dtype))


# ------------------------------------------------------------------------------------------ GPU
class _RandomModel:
    """Stand-in for the ppo2 model: ``step(obs) -> actions, values, states, neglogpacs``."""
    initial_state = None

    def __init__(self, space, n, seed):
        self.space, self.n, self.rng = space, n, np.random.RandomState(seed)

    def step(self, obs, **_):
        lo, hi = float(self.space.low[0]), float(self.space.high[0])
        a = self.rng.uniform(lo, hi, size=(self.n,) + self.space.shape).astype(np.float32)
        return a, np.zeros(self.n, np.float32), None, np.zeros(self.n, np.float32)


def _runner_run(env, model, nsteps, obs, dones):
    """baselines ppo2.Runner.run (behind learn(env=env, ...) at run.py:63-68), minus the GAE arithmetic."""
    mb_obs, mb_rewards, mb_actions, mb_dones, epinfos = [], [], [], [], []
    for _ in range(nsteps):
        actions, values, _, neglogpacs = model.step(obs, S=None, M=dones)
        mb_obs.append(obs.copy())
        mb_actions.append(actions)
        mb_dones.append(dones)
        obs[:], rewards, dones, infos = env.step(actions)
        for info in infos:
            maybeepinfo = info.get("episode")
            if maybeepinfo:
                epinfos.append(maybeepinfo)
        mb_rewards.append(rewards)
    return (np.asarray(mb_obs, dtype=obs.dtype), np.asarray(mb_rewards, dtype=np.float32), np.asarray(mb_actions),
            np.asarray(mb_dones, dtype=bool), epinfos, dones)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,reward_scale", [("quadrotor2d-v0", 1.0), ("quadrotor3d-v0", 0.25),
                                                 ("quadrotor3d-slungload-v0", 1.0), ("quadrotor2d-slungload-v0", 3.0)])
def test_ppo2_runner_loop_through_the_reference_call(env_id, reward_scale):
    """build_env (run.py:73-94) -> Runner: the verbatim make_vec_env call, obs buffer of the space's dtype, nsteps x
    ``obs[:], rewards, dones, infos = env.step(actions)``, ``info.get('episode')``; everything bitwise against the
    float32 oracle, reward_scale included (scaled rewards, unscaled episode returns — RewardScaler wraps Monitor)."""
    import gym_reinmav_b200 as rmv
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    nenv, nsteps, seed = 1536, 128, 3
    env = rmv.make_vec_env(env_id, "native", nenv or 1, seed, reward_scale=reward_scale, flatten_dict_observations=True)
    assert env.num_envs == nenv and env.observation_space.shape == (S,) and env.action_space.shape == (A,)
    model = _RandomModel(env.action_space, nenv, seed=5)
    obs = np.zeros((nenv,) + env.observation_space.shape, dtype=env.observation_space.dtype.name)   # Runner.__init__
    obs[:] = env.reset()
    dones = [False for _ in range(nenv)]
    mb_obs, mb_rew, mb_act, mb_dones, epinfos, dones = _runner_run(env, model, nsteps, obs, dones)
    assert mb_obs.shape == (nsteps, nenv, S) and mb_rew.shape == (nsteps, nenv) and mb_dones.shape == (nsteps, nenv)

    o = BatchOracle(kind, nenv, seed=seed, dtype=np.float32, params={"reward_scale": reward_scale})
    ref_obs = o.reset()
    ret, length, want_eps = np.zeros(nenv, np.float32), np.zeros(nenv, np.int64), []
    for t in range(nsteps):
        assert same_bits(mb_obs[t].astype(np.float32), ref_obs), t
        ref_obs, rew, done = o.step(mb_act[t])
        assert same_bits(mb_rew[t], rew), t
        if t + 1 < nsteps:
            assert np.array_equal(mb_dones[t + 1], done), t
        ret = (ret + o.raw_reward).astype(np.float32)
        length += 1
        for i in np.flatnonzero(done):
            want_eps.append((float(ret[i]), int(length[i])))
        ret[done], length[done] = 0, 0
    assert np.array_equal(np.asarray(dones, dtype=bool), done)
    assert len(want_eps) > 10 and [(e["r"], e["l"]) for e in epinfos] == want_eps
    assert all(set(e) == {"r", "l", "t"} and e["t"] >= 0.0 for e in epinfos)          # Monitor's record
    if reward_scale != 1.0:
        live = ~mb_dones[1:].ravel()
        assert np.any(mb_rew[:-1].ravel()[live] != 0)

    # --play loop (run.py:192-211): obs = env.reset(); step; rew[0]; done.any(); env.reset()
    obs = env.reset()
    episode_rew = 0.0
    for _ in range(20):
        actions, _, _, _ = model.step(obs)
        obs, rew, done, _ = env.step(actions)
        episode_rew += rew[0]
        env.render()
        done = done.any() if isinstance(done, np.ndarray) else done
        if done:
            obs = env.reset()
    env.close()
