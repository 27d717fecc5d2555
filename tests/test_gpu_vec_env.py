"""The reference-facing Python surface on a GPU: baselines-style VecEnv, N=1 gym.Env views, registry."""
import numpy as np
import pytest

from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle
from tests.util import same_bits

pytestmark = pytest.mark.gpu
ENV_IDS = list(KIND_OF_ID)


@pytest.fixture(scope="module")
def rmv():
    import torch
    torch.cuda.init()
    import gym_reinmav_b200
    return gym_reinmav_b200


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_vec_env_protocol_and_monitor(rmv, env_id):
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, steps, seed = 3000, 250, 4
    venv = rmv.make_vec_env(env_id, "native", num_env=n, seed=seed)
    assert venv.num_envs == n and venv.observation_space.shape == (S,) and venv.action_space.shape == (A,)
    assert float(venv.action_space.low[0]) == lo and float(venv.action_space.high[0]) == hi
    o = BatchOracle(kind, n, seed=seed, dtype=np.float32)
    obs = venv.reset()
    assert obs.shape == (n, S) and obs.dtype == np.float32 and same_bits(obs, o.reset())
    rng = np.random.RandomState(0)
    ret, length = np.zeros(n, np.float32), np.zeros(n, np.int64)
    seen = 0
    for t in range(steps):
        a = rng.uniform(lo, hi, size=(n, A)).astype(np.float32)
        obs, rew, done, infos = venv.step(a)
        obs_ref, rew_ref, done_ref = o.step(a)
        assert done.dtype == bool and np.array_equal(done, done_ref)
        assert same_bits(rew, rew_ref) and same_bits(obs, obs_ref)          # done envs report the RESET observation
        ret = (ret + rew_ref).astype(np.float32)
        length += 1
        assert len(infos) == n
        for i in np.flatnonzero(done_ref):
            ep = infos[int(i)]["episode"]                                    # baselines Monitor's info['episode']
            assert ep["l"] == length[i] and np.float32(ep["r"]) == ret[i]
            seen += 1
        assert infos[int(np.flatnonzero(~done_ref)[0])] == {}
        ret[done_ref], length[done_ref] = 0, 0
    assert seen > 20
    stats = venv.episode_stats()
    assert stats["episodes"] == seen and stats["env_steps"] == n * steps
    with pytest.raises(ValueError):
        venv.step(np.zeros((n, A + 1), np.float32))
    venv.close()
    venv.close()                                                             # idempotent


def test_vec_env_tensor_path_matches_host_path(rmv):
    import torch
    n, seed = 4096, 8
    a_env = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed)
    b_env = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed)
    oa = a_env.reset()
    ob = b_env.reset_tensor()
    assert same_bits(oa, ob.cpu().numpy())
    rng = np.random.RandomState(1)
    for t in range(120):
        act = rng.uniform(0, 10, size=(n, 4)).astype(np.float32)
        oa, ra, da, _ = a_env.step(act)
        ob, rb, db = b_env.step_tensor(torch.as_tensor(act, device="cuda:0"))
        assert same_bits(oa, ob.cpu().numpy()) and same_bits(ra, rb.cpu().numpy())
        assert np.array_equal(da, db.cpu().numpy().astype(bool))
    assert b_env.state_layout == "rows"                # the 3D kinds keep the observation matrix as their state
    rows = b_env.state_tensor()[:n]                    # [n_pad, S] view of the handle's own buffer
    assert same_bits(rows.cpu().numpy(), oa) and rows.data_ptr() == ob.data_ptr()      # step_tensor's obs IS the state
    c_env = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed, state_layout="tiles")
    c_env.reset_tensor()
    rng = np.random.RandomState(1)
    for t in range(120):
        oc, _, _ = c_env.step_tensor(torch.as_tensor(rng.uniform(0, 10, size=(n, 4)).astype(np.float32), device="cuda:0"))
    tiles = c_env.state_tensor()                       # [n/128, S, 128]
    assert same_bits(tiles.permute(0, 2, 1).reshape(-1, 10)[:n].cpu().numpy(), oa) and same_bits(oc.cpu().numpy(), oa)
    c_env.close()
    a_env.close()
    b_env.close()


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_single_env_view(rmv, env_id):
    """The gym.Env surface of the reference classes (quadrotor3d.py:42-185 and twins), N = 1."""
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    env = rmv.make(env_id)
    assert type(env).__module__.endswith("envs.native")          # run.py:36 derives env_type from this tail
    assert env.metadata == {"render.modes": ["human"]}
    assert env.dt == np.float32(0.01) and env.mass == 1.0 and env.g[-1] == -np.float32(9.8)
    assert env.observation_space.shape == (S,) and env.action_space.shape == (A,)
    assert env.steps_beyond_done is None
    # the viewer handles the reference classes carry before their first render() (quadrotor2d.py:57-59 and twins)
    handles = {5: ["quadtrans", "reftrans"], 9: ["quadtrans", "reftrans", "loadtrans"],
               10: ["render_" + k for k in ("quad1", "quad2", "rotor1", "rotor2", "rotor3", "rotor4", "velocity", "ref")]}
    handles[16] = handles[10] + ["render_load", "render_tether"]
    assert env.x_range == 1.0 and env.viewer is None and all(getattr(env, h) is None for h in handles[S])
    assert env.seed(123) == [123]
    obs = env.reset()
    assert obs.shape == (S,) and obs.dtype == np.float64 and (np.abs(obs) <= 1).all()
    o = BatchOracle(kind, 1, seed=123, dtype=np.float32)
    o.episode[:] = 1                                               # the constructor already reset once
    assert np.array_equal(obs, o.reset()[0].astype(np.float64))
    a = env.action_space.sample()
    obs, rew, done, info = env.step(a)
    obs_ref, rew_ref, done_ref = o.step(a[None], auto_reset=False)
    assert isinstance(rew, float) and isinstance(done, bool) and info == {}
    assert np.array_equal(obs, obs_ref[0].astype(np.float64)) and rew == float(rew_ref[0]) and done == bool(done_ref[0])
    # drive it out of bounds: first done pays +1 and does NOT reset; the next done pays 0 (steps_beyond_done)
    far = np.zeros(S)
    far[0] = 50.0
    if S in (9, 16):
        far[5 if S == 9 else 10] = 50.0
    env.state = far
    assert np.array_equal(env.state, far)
    _, r1, d1, _ = env.step(np.zeros(A))
    _, r2, d2, _ = env.step(np.zeros(A))
    assert (d1, r1, d2, r2) == (True, 1.0, True, 0.0)
    assert env.steps_beyond_done is not None and abs(env.state[0]) > 40
    env.reset()
    env.state = far
    assert env.step(np.zeros(A))[1] == 0.0                          # reset() does not clear it (quadrotor3d.py:182-185)
    assert env.render() is None
    env.close()


def test_unknown_id_and_registry(rmv):
    assert sorted(rmv.ENTRY_POINTS) == sorted(list(KIND_OF_ID) + ["reinmav-v0"])      # gym_reinmav/__init__.py:3-26
    with pytest.raises(KeyError):
        rmv.make("quadrotor4d-v0")
    with pytest.raises(KeyError):
        rmv.QuadBatch("nope", 4)
    with pytest.raises(rmv.RmvError):
        rmv.QuadBatch("quadrotor3d-v0", 0)


def test_params_override_repair_b(rmv):
    """quadrotor2d.py:95-98 repair B = velocity threshold 2.0 through rmv_params."""
    import torch
    b = rmv.QuadBatch("quadrotor2d-v0", 2, params={"vel_done": 2.0})
    assert b.params["vel_done"] == 2.0 and b.params["thrust_scale"] == 10.0
    b.set_state(np.array([[0, 0, 0, 0, 0], [0, 0, 0, 2.5, 0]], np.float32))
    rew = torch.empty(2, device="cuda:0")
    done = torch.empty(2, dtype=torch.uint8, device="cuda:0")
    b.step(torch.zeros((2, 2), device="cuda:0"), rew, done, auto_reset=False)
    assert done.cpu().tolist() == [0, 1]
    with pytest.raises(KeyError):
        rmv.QuadBatch("quadrotor2d-v0", 2, params={"no_such_field": 1.0})
    b.close()


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_batched_control_kernel_vs_oracle(rmv, env_id):
    """rmv_control: bit for bit the float32 model of the reference controllers (pinned operation order,
    pinned atan2 — tests/test_host_twin.py measures that model against float64: 1e-5 on the states the reference
    visits), on reset-like and wide states, both action layouts, default and moved set-points."""
    import torch
    from tests.util import wide_states
    kind = KIND_OF_ID[env_id]
    A = ACTION_DIM[kind]
    n = 50000
    s0 = wide_states(kind, n, np.random.RandomState(3))
    b = rmv.QuadBatch(env_id, n)
    b.set_state(s0)
    o32 = BatchOracle(kind, n, dtype=np.float32)
    o32.state = s0.T.copy()
    o64 = BatchOracle(kind, n, dtype=np.float64)
    o64.state = s0.T.astype(np.float64)
    for ref_pos in (None, [0.3, -0.2] if A == 2 else [0.3, -0.2, 1.5]):
        want32, want64 = o32.control(ref_pos=ref_pos).astype(np.float32), o64.control(ref_pos=ref_pos)
        act = torch.empty((n, A), dtype=torch.float32, device="cuda:0")
        b.control(act, ref_pos=ref_pos)
        got = act.cpu().numpy()
        ok = np.isfinite(want64).all(axis=1)
        assert ok.mean() > 0.99 and same_bits(got[ok], want32[ok])
        err = np.abs(got[ok].astype(np.float64) - want64[ok]) / np.maximum(1.0, np.abs(want64[ok]))
        assert np.quantile(err, 0.9999) <= 5e-6 and err.max() <= 1e-4
        soa = torch.empty((A, b.ld), dtype=torch.float32, device="cuda:0")
        b.control(soa, ref_pos=ref_pos, layout=1, act_ld=b.ld)
        assert torch.equal(soa[:, :n].t().contiguous().view(torch.int32), act.view(torch.int32))
    b.close()


@pytest.mark.parametrize("env_id,repair", [(e, "A") for e in ENV_IDS] + [("quadrotor2d-v0", "B")])
def test_reference_test_loop_control_step_reset(rmv, env_id, repair):
    """BASELINE config 0: the reference's own test loop (test/test_quadrotor2d.py:12-22 and twins) —
    control() -> step() -> reset on done, 400 steps — run on the N=1 gym.Env view and compared with the
    trajectory the reference itself produced (tests/golden ctrl_*).  Closed loop amplifies float32
    rounding, so the comparison is per step from the reference's state (teacher-forced)."""
    from tests.util import load_golden
    g = load_golden(env_id, repair)
    T, A = g["ctrl_a"].shape
    env = rmv.make(env_id, seed=0)
    if repair == "B":
        env.close()
        pytest.skip("repair B is a params override of the batch handle (test_params_override_repair_b)")
    states = np.concatenate([g["ctrl_obs0"][None], g["ctrl_obs"][:-1]], axis=0)
    S = states.shape[1]
    n_done = compared = 0
    for t in range(T):
        if S in (9, 16):
            # a state on the slung-load switching surface (|tether| = L +- rounding, i.e. every state a
            # taut step produced) takes either branch once it is rounded to float32: not comparable
            d, lp, L = (2, 5, 0.5) if S == 9 else (3, 10, 1.5)
            if abs(np.linalg.norm(states[t][lp:lp + d] - states[t][0:d]) - L) < 1e-6:
                continue
        compared += 1
        env.state = states[t]
        a = env.control()
        ctol = 1.5e-5 if env_id == "quadrotor2d-v0" else 1e-5       # see tests/test_oracle_golden.py
        assert (np.abs(a - g["ctrl_a"][t]) <= ctol * np.maximum(1.0, np.abs(g["ctrl_a"][t]))).all(), t
        obs, rew, done, _ = env.step(g["ctrl_a"][t])
        assert done == bool(g["ctrl_done"][t]), t
        if done:
            n_done += 1
        else:                                       # on done the fixture holds the reset observation
            assert (np.abs(obs - g["ctrl_obs"][t]) <= 1e-5 * np.maximum(1.0, np.abs(g["ctrl_obs"][t]))).all(), t
            assert abs(rew - g["ctrl_r"][t]) <= 1e-5 * max(1.0, abs(g["ctrl_r"][t])), t
    assert compared > 40 and (S in (9, 16) or n_done == int(g["ctrl_done"].sum()))
    env.close()


@pytest.mark.parametrize("layout", ["tiles", "rows"])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_fused_control_step_equals_control_then_step(rmv, env_id, layout):
    """rmv_control_step (state -> controller -> dynamics -> reward/done -> reset, one kernel, no action
    buffer) against rmv_control followed by rmv_step, and against the float32 model's control() + step(): actions,
    state, rewards, done flags, observation rows, episode words and statistics bit for bit, closed loop with
    reset-on-done, ragged batch; then T steps in ONE launch (rmv_control_step_many) against the same loop."""
    import torch
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    n, seed, dev, T = 128 * 37 + 21, 31, "cuda:0", 60
    two = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True, state_layout=layout)
    one = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True, state_layout=layout)
    many = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True, state_layout=layout)
    o = BatchOracle(kind, n, seed=seed, dtype=np.float32)
    start = np.random.RandomState(9).uniform(-1.5, 1.5, (n, S)).astype(np.float32)
    for b in (two, one, many):
        b.reset(); b.set_state(start)
    o.reset(); o.state = start.T.copy()
    act2 = torch.empty((n, A), dtype=torch.float32, device=dev); act1 = torch.empty_like(act2)
    r2 = torch.empty(n, dtype=torch.float32, device=dev); r1 = torch.empty_like(r2)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev); d1 = torch.empty_like(d2)
    o2 = torch.empty((n, S), dtype=torch.float32, device=dev); o1 = torch.empty_like(o2)
    rews, dones, acts = [], [], []
    for t in range(T):
        two.control(act2)
        two.step(act2, r2, d2, obs=o2)
        one.control_step(r1, d1, obs=o1 if t % 2 == 0 else None, actions_out=act1 if t % 3 else None)
        a_ref = o.control().astype(np.float32)
        obs_ref, rew_ref, done_ref = o.step(a_ref)
        if t % 3:
            assert torch.equal(act1.view(torch.int32), act2.view(torch.int32)), t
        fin = np.isfinite(a_ref).all(axis=1)
        assert fin.all() and same_bits(act2.cpu().numpy(), a_ref), t
        assert torch.equal(r1.view(torch.int32), r2.view(torch.int32)) and torch.equal(d1, d2), t
        assert same_bits(r2.cpu().numpy(), rew_ref) and np.array_equal(d2.cpu().numpy().astype(bool), done_ref), t
        if t % 2 == 0:
            assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)) and same_bits(o2.cpu().numpy(), obs_ref), t
        rews.append(r2.clone()); dones.append(d2.clone()); acts.append(act2.clone())
    assert same_bits(one.get_state(), two.get_state()) and same_bits(one.get_state(), o.obs())
    assert np.array_equal(one.get_episode_words(), two.get_episode_words())
    sa, sb = one.episode_stats(), two.episode_stats()
    assert all(sa[k] == sb[k] for k in ("episodes", "length_sum", "env_steps"))
    assert abs(sa["return_sum"] - sb["return_sum"]) <= 1e-9 * max(1.0, abs(sb["return_sum"]))      # atomic double sums
    assert int(torch.stack(dones).sum()) > 0
    # T steps, one launch
    rm = torch.empty((T, n), dtype=torch.float32, device=dev)
    dm = torch.empty((T, n), dtype=torch.uint8, device=dev)
    am = torch.empty((T, n, A), dtype=torch.float32, device=dev)
    many.control_step_many(T, rm, dm, out_ld=n, actions_out=am)
    assert torch.equal(rm.view(torch.int32), torch.stack(rews).view(torch.int32)) and torch.equal(dm, torch.stack(dones))
    assert torch.equal(am.view(torch.int32), torch.stack(acts).view(torch.int32))
    assert same_bits(many.get_state(), two.get_state()) and np.array_equal(many.get_episode_words(), two.get_episode_words())
    a, b = many.episode_stats(), two.episode_stats()
    assert a["episodes"] == b["episodes"] and a["length_sum"] == b["length_sum"] and a["env_steps"] == b["env_steps"]
    many.control_step_many(5)                          # no outputs at all: only the state moves
    for b in (two, one, many):
        b.close()


@pytest.mark.parametrize("env_id,repair", [(e, "A") for e in ENV_IDS] + [("quadrotor2d-v0", "B")])
def test_fused_control_step_on_the_reference_loop_fixtures(rmv, env_id, repair):
    """The ctrl_* fixtures are the reference's own test loop (test/test_quadrotor3d.py:16-22: control() -> step() ->
    reset on done, 400 steps).  Teacher-forced: the T states the reference visited become a batch of T envs, ONE fused
    control+step launch produces the T actions and the T next states; actions within 1e-5 (1.5e-5 quadrotor2d, see
    tests/test_oracle_golden.py), next state / reward within 1e-5, done flags identical."""
    import torch
    from tests.util import load_golden
    g = load_golden(env_id, repair)
    T, A = g["ctrl_a"].shape
    states = np.concatenate([g["ctrl_obs0"][None], g["ctrl_obs"][:-1]], axis=0)
    S = states.shape[1]
    keep = np.ones(T, dtype=bool)
    if S in (9, 16):                                   # states on the slung-load switching surface take either branch in float32
        d, lp, L = (2, 5, 0.5) if S == 9 else (3, 10, 1.5)
        keep = np.abs(np.linalg.norm(states[:, lp:lp + d] - states[:, 0:d], axis=1) - L) >= 1e-6
    b = rmv.QuadBatch(env_id, T, params={"vel_done": 2.0} if repair == "B" else None)
    b.set_state(states.astype(np.float32))
    dev = "cuda:0"
    rew = torch.empty(T, dtype=torch.float32, device=dev); done = torch.empty(T, dtype=torch.uint8, device=dev)
    obs = torch.empty((T, S), dtype=torch.float32, device=dev); act = torch.empty((T, A), dtype=torch.float32, device=dev)
    b.control_step(rew, done, obs=obs, actions_out=act, auto_reset=False)
    a, o_, r, dn = act.cpu().numpy().astype(np.float64), obs.cpu().numpy().astype(np.float64), rew.cpu().numpy(), done.cpu().numpy().astype(bool)
    ctol = 1.5e-5 if env_id == "quadrotor2d-v0" else 1e-5
    assert (np.abs(a - g["ctrl_a"]) <= ctol * np.maximum(1.0, np.abs(g["ctrl_a"])))[keep].all()
    # the step consumed the float32 action, the fixture the float64 one: the closed loop amplifies that difference by
    # the same factor it amplifies state rounding, so the state bar is the controller's
    assert np.array_equal(dn[keep], g["ctrl_done"][keep])
    live = keep & ~g["ctrl_done"]                      # on done the fixture holds the reset observation
    assert live.sum() > 40
    assert (np.abs(o_ - g["ctrl_obs"]) <= 1e-5 * np.maximum(1.0, np.abs(g["ctrl_obs"])))[live].all()
    assert (np.abs(r - g["ctrl_r"]) <= 1e-5 * np.maximum(1.0, np.abs(g["ctrl_r"])))[live].all()
    b.close()


def test_pipelined_host_step_equals_single_shot(rmv):
    """rmv_step_host cuts large batches into chunks over internal streams; results must not change."""
    n, A, S = 70000 + 77, 4, 10
    rng = np.random.RandomState(2)
    outs = []
    for chunks in (1, 0, 5):                       # single shot, auto (8 chunks), 5 ragged chunks
        b = rmv.QuadBatch("quadrotor3d-v0", n, seed=6, track_episodes=True)
        b.set_host_chunks(chunks)
        obs = np.empty((n, S), np.float32)
        rew, done = np.empty(n, np.float32), np.empty(n, np.uint8)
        b.reset_host(obs)
        r = np.random.RandomState(2)
        hist = []
        for t in range(150):
            b.step_host(r.uniform(0, 10, size=(n, A)).astype(np.float32), rew, done, obs)
            hist.append((rew.copy(), done.copy()))
        outs.append((obs.copy(), hist, b.get_episode_words(), b.episode_stats()))
        b.close()
    ref = outs[0]
    assert sum(int(d.sum()) for _, d in ref[1]) > 100
    for other in outs[1:]:
        assert same_bits(ref[0], other[0]) and np.array_equal(ref[2], other[2]) and ref[3] == other[3]
        for (r0, d0), (r1, d1) in zip(ref[1], other[1]):
            assert same_bits(r0, r1) and np.array_equal(d0, d1)


def test_ppo2_shaped_rollout_buffers(rmv):
    """BASELINE config 4 shape: quadrotor2d-v0, obs [N,5] -> actions [N,2] ~ N(0,1) (baselines' initial
    diagonal Gaussian) -> step, T steps written straight into [T, N, ...] device buffers."""
    import torch
    n, T, seed = 8192, 96, 12
    venv = rmv.QuadVecEnv("quadrotor2d-v0", n, seed=seed, monitor=False)
    buf = rmv.RolloutBuffer(venv, T)
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    buf.collect(lambda obs: torch.randn((n, 2), generator=gen, device="cuda:0"))
    o = BatchOracle("quadrotor2d-v0", n, seed=seed, dtype=np.float32)
    assert same_bits(buf.obs[0].cpu().numpy(), o.reset())
    acts = buf.actions.cpu().numpy()
    for t in range(T):
        obs_ref, rew_ref, done_ref = o.step(acts[t])
        assert same_bits(buf.obs[t + 1].cpu().numpy(), obs_ref), t
        assert same_bits(buf.rewards[t].cpu().numpy(), rew_ref), t
        assert np.array_equal(buf.dones[t].cpu().numpy().astype(bool), done_ref), t
    assert int(buf.dones.sum()) > 50
    buf.roll()
    assert torch.equal(buf.obs[0], buf.obs[T])
    assert same_bits(venv.batch.get_state(), buf.obs[T].cpu().numpy())
    venv.close()


@pytest.mark.parametrize("env_id", ["quadrotor2d-v0", "quadrotor2d-slungload-v0", "quadrotor3d-v0", "quadrotor3d-slungload-v0"])
def test_travelling_state_rollout_equals_in_place_rollout(rmv, env_id):
    """rmv_step_rows (state read from obs[t], written to obs[t+1]; the handle's tiled state untouched) against
    rmv_step with observation rows: same obs, rewards, dones, episode words, per-env episode statistics and final
    handle state, bit for bit; ragged batch, tracking on, two consecutive rollouts; plus the in-place form
    (rows_in == rows_out)."""
    import torch
    n, T, seed = 128 * 29 + 57, 80, 21
    outs = []
    for travelling in (False, True):
        venv = rmv.QuadVecEnv(env_id, n, seed=seed, monitor=True)
        buf = rmv.RolloutBuffer(venv, T, travelling=travelling)
        assert buf.travelling == travelling
        A = venv.batch.action_dim
        lo, hi = float(venv.action_space.low[0]), float(venv.action_space.high[0])
        for rollout in range(2):
            for t in range(T):
                rmv.fill_uniform_actions(buf.actions[t], n, A, 2 * lo, 2 * hi, seed, step_index=rollout * T + t)
            buf.collect(lambda obs, _it=iter(range(T)): buf.action_views[next(_it)])
            if rollout == 0:
                first = (buf.obs.cpu().numpy().copy(), buf.rewards.cpu().numpy().copy(), buf.dones.cpu().numpy().copy())
                buf.roll()
        outs.append((first, buf.obs.cpu().numpy(), buf.rewards.cpu().numpy(), buf.dones.cpu().numpy(),
                     venv.batch.get_state(), venv.batch.get_episode_words(), venv.batch.episode_stats()["episodes"]))
        venv.close()
    a, b = outs
    assert int(a[3].sum()) > 50
    for x, y in zip(a[0][:2], b[0][:2]):
        assert same_bits(x, y)
    for i in (1, 2, 4):
        assert same_bits(a[i], b[i]), i
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[0][2], b[0][2])
    assert np.array_equal(a[5], b[5]) and a[6] == b[6]

    # in place: one [N, S] buffer is both source and destination
    b1 = rmv.QuadBatch(env_id, n, seed=seed)
    b2 = rmv.QuadBatch(env_id, n, seed=seed)
    S, A = b1.state_dim, b1.action_dim
    rows = torch.empty((n, S), dtype=torch.float32, device="cuda:0")
    obs = torch.empty((n, S), dtype=torch.float32, device="cuda:0")
    b1.reset(obs=rows)
    b2.reset()
    act = torch.empty((n, A), dtype=torch.float32, device="cuda:0")
    r1 = torch.empty(n, dtype=torch.float32, device="cuda:0"); r2 = torch.empty_like(r1)
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda:0"); d2 = torch.empty_like(d1)
    for t in range(60):
        rmv.fill_uniform_actions(act, n, A, 2 * lo, 2 * hi, seed, step_index=t)
        b1.step_rows(rows, act, r1, d1, rows)
        b2.step(act, r2, d2, obs=obs)
        assert torch.equal(rows.view(torch.int32), obs.view(torch.int32)) and torch.equal(r1.view(torch.int32), r2.view(torch.int32))
        assert torch.equal(d1, d2)
    b1.close(); b2.close()


@pytest.mark.parametrize("travelling", [False, True])
def test_second_collect_continues_without_an_explicit_roll(rmv, travelling):
    """ppo2's Runner keeps self.obs between run() calls; collect() carries obs[T] into obs[0] itself (ADVICE r1:
    in travelling mode a forgotten roll() used to restart every env from the previous rollout's first observation)."""
    import torch
    n, T, seed = 2048 + 33, 40, 5
    outs = []
    for explicit_roll in (True, False):
        venv = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=seed, monitor=True)
        buf = rmv.RolloutBuffer(venv, T, travelling=travelling)
        for rollout in range(3):
            for t in range(T):
                rmv.fill_uniform_actions(buf.actions[t], n, 4, 0.0, 10.0, seed, step_index=rollout * T + t)
            buf.collect(lambda obs, _it=iter(range(T)): buf.action_views[next(_it)])
            if explicit_roll:
                buf.roll()
                buf.roll()                                      # idempotent
        outs.append((buf.obs.cpu().numpy(), buf.rewards.cpu().numpy(), buf.dones.cpu().numpy(),
                     venv.batch.get_state(), venv.batch.get_episode_words(), venv.batch.episode_stats()))
        venv.close()
    a, b = outs
    assert same_bits(a[0][1:], b[0][1:]) and same_bits(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert same_bits(a[3], b[3]) and np.array_equal(a[4], b[4]) and a[5] == b[5]
    assert int(a[2].sum()) > 20


def test_misaligned_action_rows_are_refused(rmv):
    """A row-per-env action is one 16-byte (A = 4) / 8-byte (A = 2) vector load: a view at an odd float offset used to
    fault with 'misaligned address' and poison the context; now RMV_EALIGN."""
    import torch
    n = 1000
    b = rmv.QuadBatch("quadrotor3d-v0", n)
    b.reset()
    big = torch.zeros(n * 4 + 8, dtype=torch.float32, device="cuda:0")
    rew = torch.empty(n, dtype=torch.float32, device="cuda:0")
    done = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for off in (1, 2, 3):
        with pytest.raises(rmv.RmvError) as exc:
            b.step(big[off:off + n * 4], rew, done)
        assert exc.value.code == -4 and "aligned" in str(exc.value)
    b.step(big[4:4 + n * 4], rew, done)                     # 16-byte offset: fine
    with pytest.raises(TypeError):
        b.step(big[:n * 4].double(), rew, done)
    with pytest.raises(ValueError):
        b.step(big[:n * 4].cpu(), rew, done)
    torch.cuda.synchronize()
    b.close()


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_row_state_handle_equals_tiled_handle(rmv, env_id):
    """RMV_STATE_ROWS: the handle keeps [N, S] rows and steps them in place.  Every entry point against a tiled handle,
    bit for bit: reset (all / masked, with obs), step (obs == NULL, obs == the state pointer, obs == another buffer),
    step_host (single shot and chunked), step_many, control, set/get state, episode words and statistics."""
    import torch
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, seed, dev = 70000 + 45, 17, "cuda:0"
    t_ = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True)
    r_ = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True, state_layout="rows")
    rows = torch.as_tensor(r_.state_view(), device=dev)
    assert rows.shape == (r_.ld, S)
    obs_t = torch.empty((n, S), dtype=torch.float32, device=dev); obs_r = torch.empty_like(obs_t)
    rew_t = torch.empty(n, dtype=torch.float32, device=dev); rew_r = torch.empty_like(rew_t)
    done_t = torch.empty(n, dtype=torch.uint8, device=dev); done_r = torch.empty_like(done_t)
    act = torch.empty((n, A), dtype=torch.float32, device=dev)

    def same():
        torch.cuda.synchronize()
        assert same_bits(t_.get_state(), r_.get_state()) and same_bits(r_.get_state(), rows[:n].cpu().numpy())
        assert np.array_equal(t_.get_episode_words(), r_.get_episode_words())

    t_.reset(obs=obs_t); r_.reset(obs=obs_r)
    assert torch.equal(obs_t.view(torch.int32), obs_r.view(torch.int32)); same()
    mask = torch.as_tensor((np.arange(n) % 5 == 0).astype(np.uint8), device=dev)
    t_.reset(mask=mask); r_.reset(mask=mask); same()
    n_done = 0
    for t in range(90):
        rmv.fill_uniform_actions(act, n, A, 2 * lo, 2 * hi, seed, step_index=t)
        t_.step(act, rew_t, done_t, obs=obs_t)
        r_.step(act, rew_r, done_r, obs=[None, rows, obs_r][t % 3])
        assert torch.equal(rew_t.view(torch.int32), rew_r.view(torch.int32)) and torch.equal(done_t, done_r), t
        assert torch.equal(obs_t.view(torch.int32), rows[:n].view(torch.int32)), t
        if t % 3 == 2:
            assert torch.equal(obs_t.view(torch.int32), obs_r.view(torch.int32)), t
        n_done += int(done_t.sum())
    assert n_done > 100; same()
    # controllers read the rows
    a_t = torch.empty((n, A), dtype=torch.float32, device=dev); a_r = torch.empty_like(a_t)
    t_.control(a_t); r_.control(a_r)
    assert torch.equal(a_t.view(torch.int32), a_r.view(torch.int32))
    # host path: single shot and chunked; observations leave from the state itself
    ho_t, ho_r = np.empty((n, S), np.float32), np.empty((n, S), np.float32)
    hr_t, hr_r, hd_t, hd_r = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.uint8), np.empty(n, np.uint8)
    rng = np.random.RandomState(4)
    for chunks in (1, 0, 3):
        t_.set_host_chunks(chunks); r_.set_host_chunks(chunks)
        for _ in range(4):
            a = rng.uniform(lo, hi, size=(n, A)).astype(np.float32)
            t_.step_host(a, hr_t, hd_t, ho_t); r_.step_host(a, hr_r, hd_r, ho_r)
            assert same_bits(ho_t, ho_r) and same_bits(hr_t, hr_r) and np.array_equal(hd_t, hd_r)
    same()
    t_.reset_host(ho_t); r_.reset_host(ho_r)
    assert same_bits(ho_t, ho_r); same()
    # step_many, every mode (row-state handles: the stream with row transposes, the rollout kernel on rows, one row-state launch per step)
    T = 12
    acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=1000 + t)
    rt = torch.empty((T, n), dtype=torch.float32, device=dev); rr = torch.empty_like(rt)
    dt = torch.empty((T, n), dtype=torch.uint8, device=dev); dr = torch.empty_like(dt)
    for mode in (0, 1, 2):
        t_.step_many(T, acts, rt, dt, out_ld=n, layout=0, fused=mode); r_.step_many(T, acts, rr, dr, out_ld=n, layout=0, fused=mode)
        assert torch.equal(rt.view(torch.int32), rr.view(torch.int32)) and torch.equal(dt, dr); same()
        assert r_.last_kernel == ("k_step_stream", "k_rollout", "k_step_rows")[mode]       # each mode has its row-state kernel
    # set_state from host and from device rows
    st = np.random.RandomState(5).uniform(-1, 1, (n, S)).astype(np.float32)
    t_.set_state(st); r_.set_state(st); same()
    r_.set_state_device(torch.as_tensor(st[::-1].copy(), device=dev)); t_.set_state(st[::-1].copy()); same()
    a, b = t_.episode_stats(), r_.episode_stats()
    assert a["episodes"] == b["episodes"] > 0 and a["length_sum"] == b["length_sum"] and a["env_steps"] == b["env_steps"]
    t_.close(); r_.close()
    with pytest.raises(rmv.RmvError):
        rmv.QuadBatch("reinmav-v0", 8, state_layout="rows")


def test_vec_env_zero_copy_mode(rmv):
    n = 2048
    venv = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=1, copy_outputs=False, monitor=False)
    ref = rmv.QuadVecEnv("quadrotor3d-v0", n, seed=1, monitor=False)
    o1, o2 = venv.reset(), ref.reset()
    assert same_bits(o1, o2) and o1 is venv._obs.array
    rng = np.random.RandomState(0)
    for t in range(20):
        a = rng.uniform(0, 10, size=(n, 4)).astype(np.float32)
        venv.action_buffer[...] = a                         # fill the pinned input buffer in place
        o1, r1, d1, _ = venv.step(venv.action_buffer)
        o2, r2, d2, _ = ref.step(a)
        assert same_bits(o1, o2) and same_bits(r1, r2) and np.array_equal(d1, d2)
    venv.close()
    ref.close()


@pytest.mark.parametrize("env_id", ["quadrotor2d-v0", "quadrotor3d-v0", "quadrotor3d-slungload-v0"])
def test_step_many_tensor_equals_repeated_step_tensor(rmv, env_id):
    """QuadVecEnv.step_many_tensor (T steps from an action block in one call: the step stream, on the row-state handle the
    3D kinds default to as well) against T step_tensor calls: observations, rewards, done flags and Monitor statistics."""
    import torch
    kind = KIND_OF_ID[env_id]
    A = ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, T, dev = 128 * 77 + 3, 48, "cuda:0"
    acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, A, 2 * lo, 2 * hi, 5, step_index=t)
    one, many = rmv.QuadVecEnv(env_id, n, seed=5), rmv.QuadVecEnv(env_id, n, seed=5)
    one.reset_tensor(); many.reset_tensor()
    start = np.random.RandomState(6).uniform(-2.5, 2.5, (n, STATE_DIM[kind])).astype(np.float32)    # wide: episodes end in every kind
    one.batch.set_state(start); many.batch.set_state(start)
    rews, dones = [], []
    for t in range(T):
        obs1, r, d = one.step_tensor(acts[t])
        rews.append(r.clone()); dones.append(d.clone())
    obs2, rew2, done2 = many.step_many_tensor(acts)
    assert many.batch.last_kernel == "k_step_stream" and many.batch.launches < one.batch.launches
    assert torch.equal(torch.stack(rews).view(torch.int32), rew2.view(torch.int32)) and torch.equal(torch.stack(dones), done2)
    assert torch.equal(obs1.view(torch.int32), obs2.view(torch.int32))
    a, b = one.episode_stats(), many.episode_stats()
    assert a["episodes"] == b["episodes"] > 0 and a["length_sum"] == b["length_sum"] and abs(a["return_sum"] - b["return_sum"]) < 1e-6 * abs(a["return_sum"])
    with pytest.raises(ValueError):
        many.step_many_tensor(acts[:, :-1])
    one.close(); many.close()
