"""Parity of the CUDA path (through the C-ABI) with the oracle.  Needs a GPU.

Bars (BASELINE.json north_star):
  * done masks bit-exact, state / reward bit-exact against the float32 model T1
    (same formulas, same operation order, exact fma where the kernel fuses) — every kind, every launch variant;
  * against the float64 reference arithmetic T0 (pinned to the reference source by
    tests/golden): |x_gpu - x_ref| <= TOL * max(1, |x_ref|), TOL = 1e-5, over 1000 free-running
    steps with identical (seed, action) sequences, done masks identical;
  * against the committed reference trajectories tests/golden/*.npz.
"""
import numpy as np
import pytest

from oracle import philox
from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle
from tests.util import load_golden, random_actions, same_bits, wide_states

pytestmark = pytest.mark.gpu

TOL = 1e-5
ENV_IDS = list(KIND_OF_ID)


@pytest.fixture(scope="module")
def torch():
    import torch
    torch.cuda.init()
    return torch


@pytest.fixture(scope="module")
def rmv():
    import gym_reinmav_b200
    return gym_reinmav_b200


def _buffers(torch, n, S):
    dev = "cuda:0"
    return (torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev),
            torch.empty((n, S), dtype=torch.float32, device=dev))


def _close(scale_ref, got, ref, tol=TOL):
    return np.abs(got.astype(np.float64) - ref) <= tol * np.maximum(1.0, np.abs(scale_ref))


# ------------------------------------------------------------------------------------------ reset
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_reset_draws_match_oracle(torch, rmv, env_id):
    n, off, seed = 1000, 12345678901, 77
    b = rmv.QuadBatch(env_id, n, seed=seed, env_id_offset=off)
    o = BatchOracle(env_id, n, seed=seed, env_id_offset=off, dtype=np.float32)
    assert (b.get_state() == 0).all()
    b.reset()
    assert same_bits(b.get_state(), o.reset())
    mask = (np.arange(n) % 3 == 0)
    obs = torch.empty((n, b.state_dim), dtype=torch.float32, device="cuda:0")
    b.reset(mask=torch.as_tensor(mask.astype(np.uint8), device="cuda:0"), obs=obs)
    want = o.reset(mask=mask)
    assert same_bits(b.get_state(), want) and same_bits(obs.cpu().numpy(), want)
    words = b.get_episode_words()
    assert np.array_equal(words, o.episode) and (words[mask] == 2).all() and (words[~mask] == 1).all()
    b.close()


# ------------------------------------------------------------------------------------------ one step, every variant
@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("vec", [1, 2, 4])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_single_step_bitwise(torch, rmv, env_id, vec, layout):
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    rng = np.random.RandomState(kind * 10 + vec)
    n = 20000 + 3                                         # ragged tail: not a multiple of the vector width
    s0 = wide_states(kind, n, rng)
    a = random_actions(kind, (n,), rng)
    o = BatchOracle(kind, n, dtype=np.float32)
    o.state = s0.T.copy()
    obs_ref, rew_ref, done_ref = o.step(a, auto_reset=False)

    b = rmv.QuadBatch(env_id, n)
    b.set_tuning(vec=vec)
    b.set_state(s0)
    rew, done, obs = _buffers(torch, n, S)
    if layout == "aos":
        act = torch.as_tensor(a, device="cuda:0")
        b.step(act, rew, done, obs=obs, auto_reset=False)
    else:
        ld = b.ld
        act = torch.zeros((A, ld), dtype=torch.float32, device="cuda:0")
        act[:, :n] = torch.as_tensor(a.T.copy(), device="cuda:0")
        b.step(act, rew, done, obs=obs, auto_reset=False, layout=1, act_ld=ld)
    torch.cuda.synchronize()
    got = b.get_state()
    nan_both = np.isnan(got) & np.isnan(obs_ref)
    assert ((got.view(np.uint32) == obs_ref.view(np.uint32)) | nan_both).all()
    assert np.array_equal(done.cpu().numpy().astype(bool), done_ref)
    assert same_bits(rew.cpu().numpy(), rew_ref)
    o_np = obs.cpu().numpy()
    assert ((o_np.view(np.uint32) == obs_ref.view(np.uint32)) | nan_both).all()
    assert 0.02 < done_ref.mean() < 0.6
    # gym.Env semantics: nothing was reset, the ever-done bit is set exactly on the done envs
    words = b.get_episode_words()
    assert np.array_equal((words >> 31).astype(bool), done_ref) and ((words & 0x7FFFFFFF) == 0).all()
    b.close()


# ------------------------------------------------------------------------------------------ rollouts vs T1 (bit-exact)
def _rollout_bitwise(torch, rmv, env_id, n, steps, vec, seed, check_state_every, tma=False, track=False, block=0,
                     prefetch=-1):
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    off = 5 * n
    b = rmv.QuadBatch(env_id, n, seed=seed, env_id_offset=off, track_episodes=track)
    b.set_tuning(vec=vec, block=block)
    b.set_prefetch(prefetch)
    b.set_kernel(tma=tma)
    o = BatchOracle(kind, n, seed=seed, env_id_offset=off, dtype=np.float32)
    b.reset()
    o.reset()
    rew, done, obs = _buffers(torch, n, S)
    act = torch.empty((n, A), dtype=torch.float32, device="cuda:0")
    n_done = 0
    for t in range(steps):
        rmv.fill_uniform_actions(act, n, A, lo, hi, seed, env_id_offset=off, step_index=t)
        want_obs = (t % 7 == 0) or (t + 1) % check_state_every == 0      # both kernel variants (with / without the obs copy)
        b.step(act, rew, done, obs=obs if want_obs else None)
        a_ref = philox.uniform_f32(seed, o.env_ids, np.full(n, t), philox.STREAM_ACTION, A, lo, hi).T
        if t % 97 == 0:
            assert same_bits(act.cpu().numpy(), np.ascontiguousarray(a_ref)), "action generator, step %d" % t
        _, rew_ref, done_ref = o.step(a_ref)
        assert np.array_equal(done.cpu().numpy().astype(bool), done_ref), "done mask, step %d" % t
        assert same_bits(rew.cpu().numpy(), rew_ref), "reward, step %d" % t
        n_done += int(done_ref.sum())
        if (t + 1) % check_state_every == 0 or t == steps - 1:
            assert same_bits(b.get_state(), o.obs()), "state, step %d" % t
            if want_obs:                                   # VecEnv convention: a done env reports its RESET state
                assert same_bits(obs.cpu().numpy(), o.obs()), "obs rows, step %d" % t
    words = b.get_episode_words()
    assert np.array_equal(words & 0x7FFFFFFF, o.episode) and np.array_equal((words >> 31).astype(bool), o.ever_done)
    if track:
        assert b.episode_stats()["episodes"] == n_done
    b.close()
    return n_done


def test_quadrotor3d_65536_envs_1000_steps_bitwise(torch, rmv):
    """BASELINE config 2: quadrotor3d-v0, N = 65536, random-policy actions, 1000 steps."""
    n_done = _rollout_bitwise(torch, rmv, "quadrotor3d-v0", 65536, 1000, vec=0, seed=0, check_state_every=100)
    assert 0.005 < n_done / (65536 * 1000) < 0.03        # ~1.3 % of env-steps end an episode


def test_quadrotor3d_slungload_65536_envs_bitwise(torch, rmv):
    """BASELINE config 3: slung-load variant, N = 65536 (16-dim state, both tether branches)."""
    _rollout_bitwise(torch, rmv, "quadrotor3d-slungload-v0", 65536, 300, vec=0, seed=1, check_state_every=100)


@pytest.mark.parametrize("env_id", ["quadrotor2d-v0", "quadrotor2d-slungload-v0"])
def test_planar_rollouts_bitwise(torch, rmv, env_id):
    _rollout_bitwise(torch, rmv, env_id, 65536, 300, vec=0, seed=2, check_state_every=100)


@pytest.mark.parametrize("vec", [1, 4])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_small_ragged_rollouts_bitwise(torch, rmv, env_id, vec):
    _rollout_bitwise(torch, rmv, env_id, 1000 + 1, 200, vec=vec, seed=3, check_state_every=50)


@pytest.mark.parametrize("prefetch", [0, 1])
@pytest.mark.parametrize("block", [128, 256, 512])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_every_cta_size_and_prefetch_mode_bitwise(torch, rmv, env_id, block, prefetch):
    """The launch-shape knobs of the step kernel (threads per CTA, episode-word prefetch) never change a bit:
    ragged batch (last CTA partial for every size), obs rows, episode tracking on the prefetch=1 runs."""
    _rollout_bitwise(torch, rmv, env_id, 128 * 37 + 77, 120, vec=1, seed=5, check_state_every=40, block=block,
                     prefetch=prefetch, track=bool(prefetch))


@pytest.mark.parametrize("vec", [1, 2])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_fused_rollout_equals_single_steps(torch, rmv, env_id, vec):
    """rmv_step_many(mode = RMV_MANY_ROLLOUT) — T steps in one launch, state in registers, reset inside the loop — against T
    single-step launches: same state, rewards, done flags, episode words and episode statistics, bit for bit.
    Wide initial states make a third of the envs finish at once, the batch is ragged, tracking is on."""
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, T, seed = 128 * 41 + 18, 48, 8          # even: the two-env-per-thread variant needs out_ld % 2 == 0
    rng = np.random.RandomState(3)
    start = rng.uniform(-3, 3, (n, S)).astype(np.float32)
    dev = "cuda:0"
    acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=t)
    out = []
    for fused in (2, 1):                          # RMV_MANY_LAUNCHES, RMV_MANY_ROLLOUT
        b = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=True)
        b.set_tuning(vec=vec)
        b.reset()
        b.set_state(start)
        rew = torch.empty((T, n), dtype=torch.float32, device=dev)
        done = torch.empty((T, n), dtype=torch.uint8, device=dev)
        b.step_many(T, acts, rew, done, out_ld=n, layout=0, fused=fused)
        out.append((b.get_state(), rew.cpu().numpy(), done.cpu().numpy(), b.get_episode_words(), b.episode_stats()))
        b.close()
    (s0, r0, d0, w0, st0), (s1, r1, d1, w1, st1) = out
    assert d0[0].mean() > 0.1                     # the mass-finish case really happened
    assert np.array_equal(d0, d1) and same_bits(r0, r1) and same_bits(s0, s1) and np.array_equal(w0, w1)
    assert st0["episodes"] == st1["episodes"] > 0 and st0["length_sum"] == st1["length_sum"]
    assert abs(st0["return_sum"] - st1["return_sum"]) <= 1e-9 * max(1.0, abs(st0["return_sum"]))   # atomic double sums


STREAM_CASES = [
    # n, warps_per_sm cap, block, track, layout, action prefetch     (148 SMs: cap 1 -> 148 warps own everything)
    (128 * 41 + 18, 0, 0, True, "aos", -1),          # one unit per warp, ragged last unit, small-batch CTA size, auto (prefetch on)
    (128 * 41 + 18, 0, 0, False, "aos", 0),          # the same without the action prefetch
    (20000 + 5, 1, 32, False, "soa", -1),            # 5 units per warp (the last few warps own 4), 1-warp CTAs
    (20000 + 5, 2, 128, True, "aos", 1),             # 3 units per warp, next action from the next unit / the next step
    (20000 + 5, 2, 128, True, "aos", 0),
    (17, 0, 0, False, "aos", 1),                     # less than one unit
    ((1 << 19) + 77, 0, 0, False, "aos", -1),        # more units than resident warps: k >= 2 at full occupancy
    ((1 << 19) + 77, 0, 0, False, "aos", 0),
]


@pytest.mark.parametrize("case", range(len(STREAM_CASES)))
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_step_stream_bitwise(torch, rmv, env_id, case):
    """rmv_step_many(mode = RMV_MANY_STREAM): the persistent step stream (one launch, every warp walks its own units
    through all T steps, finished envs re-drawn by their own warp once per step) against T launches of the single-step
    kernel — state, rewards, done flags, episode words and episode statistics bit for bit; two consecutive calls (the
    second starts from the first one's state), wide initial states (a third of the envs finish at once: the per-warp
    re-draw list overflows and is flushed mid-step), every ownership shape."""
    n, cap, block, track, layout, apre = STREAM_CASES[case]
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    T, seed = (12 if n > 100000 else 40), 8
    start = np.random.RandomState(3).uniform(-3, 3, (n, S)).astype(np.float32)
    dev = "cuda:0"
    ld = ((n + 127) // 128) * 128
    if layout == "aos":
        acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
        for t in range(T):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=t)
        kw = dict(layout=0)
    else:
        acts = torch.zeros((T, A, ld), dtype=torch.float32, device=dev)
        for t in range(T):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=t, layout=1, act_ld=ld)
        kw = dict(layout=1, act_ld=ld)
    out = []
    for mode in (2, 0):                           # RMV_MANY_LAUNCHES, RMV_MANY_STREAM
        b = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=track)
        b.set_stream(warps_per_sm=cap, block=block, action_prefetch=apre)
        b.reset()
        b.set_state(start)
        rew = torch.empty((2, T, n), dtype=torch.float32, device=dev)
        done = torch.empty((2, T, n), dtype=torch.uint8, device=dev)
        for call in range(2):
            b.step_many(T, acts, rew[call], done[call], out_ld=n, fused=mode, **kw)
        torch.cuda.synchronize()
        out.append((b.get_state(), rew.cpu().numpy(), done.cpu().numpy(), b.get_episode_words(),
                    b.episode_stats() if track else None))
        b.close()
    (s0, r0, d0, w0, st0), (s1, r1, d1, w1, st1) = out
    assert d0[0, 0].mean() > 0.1                  # the mass-finish case really happened
    assert np.array_equal(d0, d1) and same_bits(r0, r1) and same_bits(s0, s1) and np.array_equal(w0, w1)
    if track:
        assert st0["episodes"] == st1["episodes"] > 0 and st0["length_sum"] == st1["length_sum"]
        assert abs(st0["return_sum"] - st1["return_sum"]) <= 1e-9 * max(1.0, abs(st0["return_sum"]))   # atomic double sums


ROW_STREAM_CASES = [(128 * 41 + 18, 0, 0, True, "aos", -1), (128 * 41 + 18, 0, 0, False, "aos", 0), (20000 + 5, 1, 32, False, "soa", -1),
                    (20000 + 5, 2, 128, True, "aos", 1), (17, 0, 0, False, "aos", 1), ((1 << 19) + 77, 0, 0, True, "aos", -1)]


@pytest.mark.parametrize("case", range(len(ROW_STREAM_CASES)))
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_step_stream_on_row_state_handles_bitwise(torch, rmv, env_id, case):
    """The step stream on a RMV_STATE_ROWS handle (the unit's 32 rows go through the warp's shared tile both ways, a
    finished env's row is stored and then overwritten by the re-draw) against T launches on a tiled handle: state,
    rewards, done flags, episode words and statistics bit for bit, mass finishes included; then the closed loop
    (rmv_control_step_many) the same way, actions included."""
    n, cap, block, track, layout, apre = ROW_STREAM_CASES[case]
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    T, seed = (10 if n > 100000 else 36), 21
    start = np.random.RandomState(5).uniform(-3, 3, (n, S)).astype(np.float32)
    dev = "cuda:0"
    ld = ((n + 127) // 128) * 128
    if layout == "aos":
        acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
        for t in range(T):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=t)
        kw = dict(layout=0)
    else:
        acts = torch.zeros((T, A, ld), dtype=torch.float32, device=dev)
        for t in range(T):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, seed, step_index=t, layout=1, act_ld=ld)
        kw = dict(layout=1, act_ld=ld)
    out = []
    for state_layout, mode in (("tiles", 2), ("rows", 0)):
        b = rmv.QuadBatch(env_id, n, seed=seed, track_episodes=track, state_layout=state_layout)
        b.set_stream(warps_per_sm=cap, block=block, action_prefetch=apre)
        b.reset()
        b.set_state(start)
        rew = torch.empty((3, T, n), dtype=torch.float32, device=dev)
        done = torch.empty((3, T, n), dtype=torch.uint8, device=dev)
        ctl = torch.zeros((T, n, A), dtype=torch.float32, device=dev)
        launches0 = b.launches
        for call in range(2):
            b.step_many(T, acts, rew[call], done[call], out_ld=n, fused=mode, **kw)
        if state_layout == "rows":
            # the library keeps one launch per step on a row-state handle above 64 MB per step (rmv_api.cu: stream_pays_for)
            big = n * (4 * (2 * S + A + 1) + 1) > (64 << 20) or not rmv._capi.lib().rmv_stream_pays(kind, n)
            assert b.last_kernel == ("k_step_rows" if big else "k_step_stream")
            assert b.launches - launches0 == (2 * T if big else 2)
            rows = torch.as_tensor(b.state_view(), device=dev)[:n].cpu().numpy()
            assert same_bits(rows, b.get_state())
        if mode == 0:
            b.control_step_many(T, rew[2], done[2], out_ld=n, actions_out=ctl)
        else:
            for t in range(T):
                b.control_step(rew[2, t], done[2, t], actions_out=ctl[t])
        torch.cuda.synchronize()
        out.append((b.get_state(), rew.cpu().numpy(), done.cpu().numpy(), b.get_episode_words(), ctl.cpu().numpy(),
                    b.episode_stats() if track else None))
        b.close()
    (s0, r0, d0, w0, c0, st0), (s1, r1, d1, w1, c1, st1) = out
    assert d0[0, 0].mean() > 0.1 or n < 100              # the mass-finish case really happened
    assert np.array_equal(d0, d1) and same_bits(r0, r1) and same_bits(s0, s1) and np.array_equal(w0, w1) and same_bits(c0, c1)
    if track:
        assert st0["episodes"] == st1["episodes"] > 0 and st0["length_sum"] == st1["length_sum"]
        assert abs(st0["return_sum"] - st1["return_sum"]) <= 1e-9 * max(1.0, abs(st0["return_sum"]))


@pytest.mark.parametrize("n,cap,block", [(128 * 41 + 18, 0, 0), (20000 + 5, 1, 32), (20000 + 6, 2, 128), (17, 0, 0), (63, 0, 0),
                                         ((1 << 19) + 77, 0, 0), ((1 << 20) + 2, 0, 0)])
def test_step_stream_pair_bitwise(torch, rmv, n, cap, block):
    """The two-envs-per-thread step stream on the packed float32 pipe (k_step_stream_pair: FFMA2 / FMUL2 / FADD2, 64-bit
    row accesses; quadrotor3d) against T launches of the one-env-per-thread kernel: state, rewards, done flags and
    episode words bit for bit — odd and even batch sizes (a lane with one env, a padding lane), less than one unit,
    several units per warp, wide initial states (mass finish: the warp-local flush), two consecutive calls."""
    env_id = "quadrotor3d-v0"
    T, seed, dev = (12 if n > 100000 else 40), 8, "cuda:0"
    start = np.random.RandomState(3).uniform(-3, 3, (n, 10)).astype(np.float32)
    acts = torch.empty((T, n, 4), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, 4, 0.0, 10.0, seed, step_index=t)
    out = []
    for mode, vec in ((2, 1), (0, 2)):
        b = rmv.QuadBatch(env_id, n, seed=seed)
        b.set_tuning(vec=vec)
        b.set_stream(warps_per_sm=cap, block=block)
        b.reset()
        b.set_state(start)
        ld = n + (n % 2)                              # even leading dimension: the pair kernel stores float2 / uchar2
        rew = torch.zeros((2, T, ld), dtype=torch.float32, device=dev)
        done = torch.zeros((2, T, ld), dtype=torch.uint8, device=dev)
        for call in range(2):
            b.step_many(T, acts, rew[call], done[call], out_ld=ld, layout=0, fused=mode)
        torch.cuda.synchronize()
        out.append((b.get_state(), rew[:, :, :n].cpu().numpy(), done[:, :, :n].cpu().numpy(), b.get_episode_words(), b.last_kernel))
        b.close()
    (s0, r0, d0, w0, k0), (s1, r1, d1, w1, k1) = out
    assert d0[0, 0].mean() > 0.1 and (k0, k1) == ("k_step", "k_step_stream_pair")
    assert np.array_equal(d0, d1) and same_bits(r0, r1) and same_bits(s0, s1) and np.array_equal(w0, w1)


def test_step_stream_without_outputs_and_vs_model(torch, rmv):
    """reward_dev / done_dev may be NULL (only the state matters), and the stream equals the float32 model directly."""
    env_id, n, T, seed = "quadrotor3d-slungload-v0", 9000, 30, 2
    kind = KIND_OF_ID[env_id]
    lo, hi = ACTION_BOUNDS[kind]
    acts = torch.empty((T, n, 4), dtype=torch.float32, device="cuda:0")
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, 4, lo, hi, seed, step_index=t)
    b = rmv.QuadBatch(env_id, n, seed=seed)
    o = BatchOracle(kind, n, seed=seed, dtype=np.float32)
    b.reset(); o.reset()
    b.step_many(T, acts, None, None, layout=0, fused=0)
    a = acts.cpu().numpy()
    for t in range(T):
        o.step(a[t])
    assert same_bits(b.get_state(), o.obs())
    words = b.get_episode_words()
    assert np.array_equal(words & 0x7FFFFFFF, o.episode) and np.array_equal((words >> 31).astype(bool), o.ever_done)
    b.close()


def test_step_is_cuda_graph_capturable(torch, rmv):
    """rmv_step only enqueues kernels on the caller's stream (no host sync, no allocation), so a training loop can
    capture it in a CUDA graph together with its policy: 8 captured steps replayed 3 times == 24 plain steps.
    (Untracked handle: with RMV_TRACK_EPISODES the step counter is a launch argument and would go stale in replays.)"""
    env_id, n, seed, K = "quadrotor3d-v0", 128 * 50 + 3, 11, 8
    dev = "cuda:0"
    acts = torch.empty((K, n, 4), dtype=torch.float32, device=dev)
    for t in range(K):
        rmv.fill_uniform_actions(acts[t], n, 4, 0.0, 10.0, seed, step_index=t)
    rew = torch.empty((K, n), dtype=torch.float32, device=dev)
    done = torch.empty((K, n), dtype=torch.uint8, device=dev)
    obs = torch.empty((n, 10), dtype=torch.float32, device=dev)

    def steps(b):
        for t in range(K):
            b.step(acts[t], rew[t], done[t], obs=obs if t == K - 1 else None)

    ref = rmv.QuadBatch(env_id, n, seed=seed)
    ref.reset()
    for _ in range(3):
        steps(ref)
    torch.cuda.synchronize()
    want = (ref.get_state(), rew.cpu().numpy().copy(), done.cpu().numpy().copy(), obs.cpu().numpy().copy(), ref.get_episode_words())
    ref.close()

    b = rmv.QuadBatch(env_id, n, seed=seed)
    b.reset()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            steps(b)
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    got = (b.get_state(), rew.cpu().numpy(), done.cpu().numpy(), obs.cpu().numpy(), b.get_episode_words())
    b.close()
    for x, y in zip(want, got):
        assert same_bits(x, y)


def test_step_many_graph_replay_bitwise(torch, rmv):
    """rmv_step_many(mode = RMV_MANY_LAUNCHES) captures a call shape in a CUDA graph the second time it sees it (consecutively or not) and
    replays it from then on: plain launch, capture + launch, replays, a second call shape (LRU), and a re-seed in between (the cached
    graphs carry the old Philox key and must not be used) — all against single rmv_step launches, bit for bit."""
    env_id, n, seed, K = "quadrotor3d-v0", 128 * 60 + 11, 13, 16
    dev = "cuda:0"
    acts = torch.empty((K, n, 4), dtype=torch.float32, device=dev)
    for t in range(K):
        rmv.fill_uniform_actions(acts[t], n, 4, 0.0, 14.0, seed, step_index=t)      # wide: plenty of resets
    rew_a = torch.empty((K, n), dtype=torch.float32, device=dev); done_a = torch.empty((K, n), dtype=torch.uint8, device=dev)
    rew_b = torch.empty_like(rew_a); done_b = torch.empty_like(done_a)
    a = rmv.QuadBatch(env_id, n, seed=seed)
    b = rmv.QuadBatch(env_id, n, seed=seed)
    a.reset(); b.reset()
    plan = [(K, None), (K, None), (K, None), (6, None), (6, None), (6, None), (K, None), (K, 99), (K, None), (K, None), (K, None),
            (K, 7), (6, None), (K, None), (6, None), (K, None), (6, None)]      # alternating shapes are captured as well
    for m, reseed in plan:
        if reseed is not None:
            a.set_seed(reseed); b.set_seed(reseed)
        a.step_many(m, acts, rew_a, done_a, out_ld=n, layout=0, fused=2)        # RMV_MANY_LAUNCHES
        for t in range(m):
            b.step(acts[t], rew_b[t], done_b[t])
        torch.cuda.synchronize()
        assert torch.equal(rew_a[:m].view(torch.int32), rew_b[:m].view(torch.int32)) and torch.equal(done_a[:m], done_b[:m])
        assert same_bits(a.get_state(), b.get_state())
        assert np.array_equal(a.get_episode_words(), b.get_episode_words())
    assert int(done_b.sum()) > 0
    a.close(); b.close()


def test_l2_persist_window_bitwise_and_refusal(torch, rmv):
    """Opt-in L2 access-policy window on the state: same bits; a state larger than the device's persisting
    carve-out is refused (a partial window thrashes) instead of silently running slower."""
    env_id, n, seed = "quadrotor3d-slungload-v0", 128 * 301 + 9, 6
    kind = KIND_OF_ID[env_id]
    lo, hi = ACTION_BOUNDS[kind]
    states = []
    for mode in (0, 1):
        b = rmv.QuadBatch(env_id, n, seed=seed)
        b.set_l2_persist(mode)
        b.reset()
        rew, done, _ = _buffers(torch, n, STATE_DIM[kind])
        act = torch.empty((n, 4), dtype=torch.float32, device="cuda:0")
        for t in range(60):
            rmv.fill_uniform_actions(act, n, 4, lo, hi, seed, step_index=t)
            b.step(act, rew, done)
        states.append((b.get_state(), rew.cpu().numpy(), b.get_episode_words()))
        b.set_l2_persist(0)
        b.close()
    for x, y in zip(*states):
        assert same_bits(x, y)
    big = rmv.QuadBatch(env_id, 1 << 22, seed=seed)            # 268 MB of state: larger than any carve-out of a 126 MB L2
    with pytest.raises(Exception, match="carve-out"):
        big.set_l2_persist(1)
    big.close()


@pytest.mark.parametrize("n,track", [(1000 + 1, False), (127, True), (128 * 700 + 5, True), (65536, False)])
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_tma_pipelined_kernel_bitwise(torch, rmv, env_id, n, track):
    """The persistent TMA-bulk / mbarrier-pipelined variant of the step kernel: same bits as the float32
    model, with more tiles than resident CTAs (pipeline wrap-around), ragged last tiles, obs rows, tracking."""
    _rollout_bitwise(torch, rmv, env_id, n, 200, vec=1, seed=4, check_state_every=50, tma=True, track=track)


# ------------------------------------------------------------------------------------------ vs float64 reference arithmetic
@pytest.mark.parametrize("env_id,n", [("quadrotor3d-v0", 16384), ("quadrotor2d-v0", 16384),
                                      ("quadrotor3d-slungload-v0", 8192), ("quadrotor2d-slungload-v0", 8192)])
def test_free_running_1000_steps_vs_float64(torch, rmv, env_id, n):
    """north_star: state/reward within 1e-5 and done masks bit-exact vs the reference over 1000 steps — all four
    kinds, free running, same (seed, action) sequence on both sides, every disagreement ACCOUNTED for
    (tests/accounting.py): a done flag may differ only where the reference's own norm is within 1e-5 of the threshold,
    a slung-load step only where the reference's tether margin is on the switching surface and the kernel's result is
    the reference's OTHER branch; each such event is counted and re-synchronised, the counts are bounded cumulatively
    (tests/test_accounting_cpu.py: check_counts), everything else is within 1e-5 with identical done flags."""
    from tests import accounting
    from tests.test_accounting_cpu import check_counts
    c = accounting.free_running(accounting.CudaStepper, env_id, n=n, steps=1000, seed=3)
    print("%s: %s" % (env_id, c))
    check_counts(env_id, c)
    if "slung" not in env_id:
        assert c["done_flips"] <= 2                      # cumulative over 1.6e7 env-steps (0 with this seed)


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_teacher_forced_vs_float64(torch, rmv, env_id):
    """Each step the float64 oracle starts from the kernel's own float32 state (exact in float64).
    For the slung-load kinds this is the meaningful comparison: free-running, the reference itself
    decides the taut/slack branch on 1-ulp noise after every taut step."""
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    n, steps, seed = 8192, 300, 11
    b = rmv.QuadBatch(env_id, n, seed=seed)
    o = BatchOracle(kind, n, seed=seed, dtype=np.float64)
    b.reset()
    o.reset()
    rew, done, obs = _buffers(torch, n, S)
    state = b.get_state()
    taut_steps = branch_mismatch = done_mismatch = 0
    for t in range(steps):
        a = philox.uniform_f32(seed, o.env_ids, np.full(n, t), philox.STREAM_ACTION, A, lo, hi).T.copy()
        o.state = state.T.astype(np.float64)
        words = b.get_episode_words()
        o.episode, o.ever_done = (words & 0x7FFFFFFF).astype(np.uint32), (words >> 31).astype(bool)
        b.step(torch.as_tensor(a, device="cuda:0"), rew, done, obs=obs)
        obs_ref, rew_ref, done_ref = o.step(a)
        d, got, r = done.cpu().numpy().astype(bool), obs.cpu().numpy(), rew.cpu().numpy()
        mism = d != done_ref
        assert (o.done_margin[mism] < 1e-5).all()
        done_mismatch += int(mism.sum())
        ok = ~mism
        good = _close(obs_ref, got, obs_ref, tol=2e-6).all(axis=1) & _close(rew_ref, r, rew_ref, tol=2e-6)
        if o.tether_margin is not None:
            taut_steps += int((o.tether_margin >= 0).sum())
            # the only admissible disagreement: exactly on the switching surface (float64 tie)
            bad = ok & ~good
            assert (np.abs(o.tether_margin[bad]) < 1e-12).all()
            branch_mismatch += int(bad.sum())
        else:
            assert good[ok].all(), "step %d" % t
        state = got
    assert done_mismatch <= 2
    if o.tether_margin is not None:
        assert taut_steps > 0.1 * n * steps and branch_mismatch <= 5, (taut_steps, branch_mismatch)
    b.close()


# ------------------------------------------------------------------------------------------ reference fixtures
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_golden_single_steps(torch, rmv, env_id):
    """tests/golden kat_*: outputs of the reference's own step() (float64) from float32-exact inputs."""
    g = load_golden(env_id)
    k, S = g["kat_s0"].shape
    b = rmv.QuadBatch(env_id, k)
    b.set_state(g["kat_s0"].astype(np.float32))
    rew, done, obs = _buffers(torch, k, S)
    b.step(torch.as_tensor(g["kat_a"].astype(np.float32), device="cuda:0"), rew, done, obs=obs, auto_reset=False)
    assert np.array_equal(done.cpu().numpy().astype(bool), g["kat_done"])
    assert _close(g["kat_s1"], obs.cpu().numpy(), g["kat_s1"], tol=2e-6).all()
    assert _close(g["kat_r"], rew.cpu().numpy(), g["kat_r"], tol=2e-6).all()
    b.close()


@pytest.mark.parametrize("env_id,repair", [("quadrotor3d-v0", "A"), ("quadrotor2d-v0", "A"), ("quadrotor2d-v0", "B")])
def test_golden_trajectories_free_running(torch, rmv, env_id, repair):
    """400 reference steps x 6 envs with reset-on-done: same seed, same actions, free running."""
    g = load_golden(env_id, repair)
    T, E, A = g["traj_a"].shape
    S = g["traj_obs"].shape[2]
    params = {"vel_done": 2.0} if repair == "B" else None
    b = rmv.QuadBatch(env_id, E, seed=int(g["traj_seed"]), params=params)
    b.reset()
    assert np.array_equal(b.get_state().astype(np.float64), g["traj_obs0"])
    rew, done, obs = _buffers(torch, E, S)
    for t in range(T):
        b.step(torch.as_tensor(g["traj_a"][t].astype(np.float32), device="cuda:0"), rew, done, obs=obs)
        assert np.array_equal(done.cpu().numpy().astype(bool), g["traj_done"][t]), t
        assert _close(g["traj_obs"][t], obs.cpu().numpy(), g["traj_obs"][t]).all(), t
        assert _close(g["traj_r"][t], rew.cpu().numpy(), g["traj_r"][t]).all(), t
    term = g["traj_r"][g["traj_done"]]
    assert (term == 1.0).sum() >= 1 and (term == 0.0).sum() >= 1     # first-ever done pays +1, later ones 0
    b.close()


@pytest.mark.parametrize("env_id", ["quadrotor2d-slungload-v0", "quadrotor3d-slungload-v0"])
def test_golden_trajectories_slung_accounted(torch, rmv, env_id):
    """The slung-load fixtures (400 reference steps x 6 envs, reset-on-done) on the GPU, the way
    tests/test_oracle_golden.py runs them on the oracle: every step starts from the reference's previous state; a step
    that misses the fixture must start on the switching surface (|tether margin| < 1e-6) and equal one of the reference
    arithmetic's two branch results; done flags equal the fixture's on every step."""
    from tests import accounting
    compared, flips = accounting.golden_trajectory(accounting.CudaStepper, env_id, load_golden(env_id))
    assert compared == 2400 and flips < 0.25 * compared
    print("%s: %d of %d fixture steps took the other branch of a surface state" % (env_id, flips, compared))


# ------------------------------------------------------------------------------------------ properties at full size
def test_full_size_properties_quadrotor3d_2pow20(torch, rmv):
    """N = 2^20 (the metric's size): size-independent properties instead of a per-element oracle."""
    n, A, S, seed = 1 << 20, 4, 10, 5
    dev = "cuda:0"
    T = 96                                                # long enough for episodes to end (mean length ~78)
    acts = torch.empty((T, A, n), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, A, 0.0, 10.0, seed, step_index=t, layout=1, act_ld=n)
    rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)

    def run(fused, vec, shards):
        outs = []
        per = n // shards
        for k in range(shards):
            b = rmv.QuadBatch("quadrotor3d-v0", per, seed=seed, env_id_offset=k * per)
            b.set_tuning(vec=vec)
            b.reset()
            r = torch.empty((T, per), dtype=torch.float32, device=dev)
            d = torch.empty((T, per), dtype=torch.uint8, device=dev)
            a = acts[:, :, k * per:(k + 1) * per].contiguous()
            b.step_many(T, a, r, d, out_ld=per, layout=1, act_ld=per, fused=fused)
            torch.cuda.synchronize()
            outs.append((torch.as_tensor(b.state_view(), device=dev).clone(), r, d))
            b.close()
        return (torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 1), torch.cat([o[2] for o in outs], 1))

    s1, r1, d1 = run(fused=False, vec=2, shards=1)
    # (1) T launches == one fused launch, bit for bit; (2) independent of the vector width;
    # (3) independent of how the envs are sharded (global env ids key the reset stream)
    for fused, vec, shards in [(True, 2, 1), (False, 1, 1), (True, 4, 1), (False, 2, 4)]:
        s2, r2, d2 = run(fused, vec, shards)
        assert torch.equal(s1.view(torch.int32), s2.view(torch.int32)), (fused, vec, shards)
        assert torch.equal(r1.view(torch.int32), r2.view(torch.int32)) and torch.equal(d1, d2), (fused, vec, shards)
    # (4) reward is -|p'| where not done, and a first-ever done pays exactly +1
    frac = d1.float().mean().item()
    assert 0.0002 < frac < 0.03
    assert ((r1 <= 0) | (d1 == 1)).all() and (r1[d1 == 1] >= 0).all()
    # (5) a done env was re-drawn: its state is inside the reset cube right after
    last = d1[-1].bool()
    rows = s1.permute(0, 2, 1).reshape(n, S)             # tiles [n/128, S, 128] -> rows [n, S]
    assert (rows[last].abs() <= 1).all()
    assert torch.isfinite(s1).all()


# ------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n", [1, 2, 31, 127, 128, 129, 255, 257, 1025])
def test_ragged_sizes_all_variants(torch, rmv, n):
    """Tile / warp / vector-width boundaries: every kind, every V, obs rows on, auto-reset on."""
    for env_id in ENV_IDS:
        kind = KIND_OF_ID[env_id]
        S, A = STATE_DIM[kind], ACTION_DIM[kind]
        for vec in (1, 2, 4):
            rng = np.random.RandomState(n + vec)
            o = BatchOracle(kind, n, seed=n, dtype=np.float32)
            b = rmv.QuadBatch(env_id, n, seed=n)
            b.set_tuning(vec=vec)
            b.reset()
            o.reset()
            rew, done, obs = _buffers(torch, n, S)
            s0 = wide_states(kind, max(n, 8), rng)[:n]
            b.set_state(s0)
            o.state = s0.T.copy()
            for t in range(6):
                a = random_actions(kind, (n,), rng)
                b.step(torch.as_tensor(a, device="cuda:0"), rew, done, obs=obs)
                obs_ref, rew_ref, done_ref = o.step(a)
                assert np.array_equal(done.cpu().numpy().astype(bool), done_ref), (env_id, vec, t)
                assert same_bits(rew.cpu().numpy(), rew_ref) and same_bits(obs.cpu().numpy(), obs_ref), (env_id, vec, t)
            assert same_bits(b.get_state(), o.obs())
            b.close()


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_non_finite_and_extreme_inputs(torch, rmv, env_id):
    """NaN / inf / huge actions and states: NaN norms are 'not done' (strict '>' on the norm, as the
    reference's comparisons), infinities are done; the kernel and the float32 model agree bit for bit."""
    kind = KIND_OF_ID[env_id]
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    n = 512
    rng = np.random.RandomState(9)
    s0 = wide_states(kind, n, rng)
    a = random_actions(kind, (n,), rng)
    specials = np.array([np.nan, np.inf, -np.inf, 1e38, -1e38, 1e-40, -0.0, 3e5, -7e9], np.float32)
    for i in range(n):
        if i % 3 == 0:
            a[i, rng.randint(A)] = specials[rng.randint(len(specials))]
        if i % 5 == 0:
            s0[i, rng.randint(S)] = specials[rng.randint(len(specials))]
    s0[7] = 0.0                                            # zero quaternion / zero tether vector
    o = BatchOracle(kind, n, dtype=np.float32)
    o.state = s0.T.copy()
    obs_ref, rew_ref, done_ref = o.step(a, auto_reset=False)
    b = rmv.QuadBatch(env_id, n)
    b.set_state(s0)
    rew, done, obs = _buffers(torch, n, S)
    b.step(torch.as_tensor(a, device="cuda:0"), rew, done, obs=obs, auto_reset=False)
    got, r, d = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool)
    assert np.array_equal(d, done_ref)
    both_nan = np.isnan(got) & np.isnan(obs_ref)
    assert ((got.view(np.uint32) == obs_ref.view(np.uint32)) | both_nan).all()
    rn = np.isnan(r) & np.isnan(rew_ref)
    assert ((r.view(np.uint32) == rew_ref.view(np.uint32)) | rn).all()
    assert np.isnan(got).any() and done_ref.any() and not done_ref[np.isnan(rew_ref)].any()
    b.close()


def test_deterministic_and_seed_sensitive(torch, rmv):
    def run(seed):
        b = rmv.QuadBatch("quadrotor3d-slungload-v0", 5000, seed=seed)
        b.reset()
        act = torch.empty((5000, 4), device="cuda:0")
        rew = torch.empty(5000, device="cuda:0")
        done = torch.empty(5000, dtype=torch.uint8, device="cuda:0")
        for t in range(120):
            rmv.fill_uniform_actions(act, 5000, 4, -10, 10, 3, step_index=t)
            b.step(act, rew, done)
        out = b.get_state()
        b.close()
        return out
    a, b_, c = run(1), run(1), run(2)
    assert same_bits(a, b_) and not np.array_equal(a, c)
