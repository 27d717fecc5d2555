"""Free-running / fixture comparisons of a float32 stepper with the float64 reference arithmetic, with every
disagreement ACCOUNTED for — shared by the CPU suite (stepper = the float32 model, oracle/batch.py dtype=float32)
and the GPU suite (stepper = the CUDA path through the C-ABI, which equals that model bit for bit).

Why accounting is needed (DESIGN.md §4 "Switching surface"): a float32 trajectory cannot follow a float64 one
through a discontinuity that the float64 trajectory itself decides on rounding noise.

* ``done``: ``norm > threshold`` (quadrotor3d.py:106-110).  A disagreement is admissible only when the reference's
  own norm is within TOL of the threshold (``done_margin``).
* slung-load branch: ``norm(tether) >= L`` (quadrotor3d_slungload.py:104, quadrotor2d_slungload.py:96).  After every
  taut step the projection (:126-127) leaves ``|tether| = L(1 +- ulp)``, so on the NEXT step the branch is a coin
  toss between any two arithmetics — about a quarter of all slung env-steps start there.  A disagreement is admissible
  only when the reference's ``tether_margin`` is within the window AND the float32 result is within TOL of what the
  reference computes from the same state with the OTHER branch.

Every admissible event is counted and the float32 side is re-synchronised to the reference state; everything else
must be within TOL with identical done flags.  The callers assert cumulative bounds on the counts.
"""
import numpy as np

from oracle import philox
from oracle.batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM, BatchOracle

TOL = 1e-5            # north_star: state / reward within 1e-5 (relative to max(1, |x|))
SURFACE = 1e-6        # |tether| - L after a taut step, in either arithmetic
EVER_DONE = 0x80000000


class ModelStepper:
    """The float32 model (tier T1) behind the stepper interface — the CPU stand-in for the CUDA path."""

    def __init__(self, env_id, n, seed, params=None):
        self.o = BatchOracle(env_id, n, seed=seed, dtype=np.float32, params=params)

    def reset(self):
        return self.o.reset()

    def step(self, a):
        return self.o.step(a)

    def resync(self, mask, rows, episode, ever_done):
        self.o.state[:, mask] = rows[mask].T.astype(np.float32)
        self.o.episode[mask] = episode[mask]
        self.o.ever_done[mask] = ever_done[mask]

    def set_rows(self, rows):
        self.o.state = np.ascontiguousarray(rows.T.astype(np.float32))

    def close(self):
        pass


class CudaStepper:
    """The CUDA path (QuadBatch -> ctypes -> librmv.so)."""

    def __init__(self, env_id, n, seed, params=None):
        import torch
        import gym_reinmav_b200 as rmv
        self.torch = torch
        self.b = rmv.QuadBatch(env_id, n, seed=seed, params=params)
        dev = "cuda:0"
        self.rew = torch.empty(n, dtype=torch.float32, device=dev)
        self.done = torch.empty(n, dtype=torch.uint8, device=dev)
        self.obs = torch.empty((n, self.b.state_dim), dtype=torch.float32, device=dev)

    def reset(self):
        self.b.reset(obs=self.obs)
        return self.obs.cpu().numpy()

    def step(self, a):
        act = self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device="cuda:0")
        self.b.step(act, self.rew, self.done, obs=self.obs)
        return self.obs.cpu().numpy(), self.rew.cpu().numpy(), self.done.cpu().numpy().astype(bool)

    def resync(self, mask, rows, episode, ever_done):
        st = self.b.get_state()
        st[mask] = rows[mask].astype(np.float32)
        self.b.set_state(st)
        w = self.b.get_episode_words()
        w[mask] = (episode[mask] | (ever_done[mask].astype(np.uint32) << 31)).astype(np.uint32)
        self.b.set_episode_words(w)

    def set_rows(self, rows):
        self.b.set_state(np.ascontiguousarray(rows, dtype=np.float32))

    def close(self):
        self.b.close()


def _off(got, r, d, ref, rr, dr, tol):
    """Per env: does (got, r, d) differ from (ref, rr, dr) beyond tol / in the done flag?"""
    return ((np.abs(got.astype(np.float64) - ref) > tol * np.maximum(1.0, np.abs(ref))).any(axis=1)
            | (np.abs(r.astype(np.float64) - rr) > tol * np.maximum(1.0, np.abs(rr))) | (d != dr))


def free_running(stepper_cls, env_id, n, steps, seed, tol=TOL):
    """``steps`` free-running steps of ``n`` envs from the same (seed, action) sequence on both sides.  Returns the
    event counts; raises AssertionError on the first disagreement that is not admissible."""
    kind = KIND_OF_ID[env_id]
    A = ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    g = stepper_cls(env_id, n, seed)
    o = BatchOracle(kind, n, seed=seed, dtype=np.float64)
    alt = BatchOracle(kind, n, seed=seed, dtype=np.float64)          # the reference with the other tether branch
    first = g.reset()
    assert np.array_equal(first.astype(np.float64), o.reset())
    c = dict(env_steps=n * steps, near_surface=0, branch_flips=0, drift_flips=0, done_flips=0, dones=0, taut=0, worst=0.0)
    for t in range(steps):
        a = philox.uniform_f32(seed, o.env_ids, np.full(n, t), philox.STREAM_ACTION, A, lo, hi).T.copy()
        prev, episode, ever = o.state.copy(), o.episode.copy(), o.ever_done.copy()
        got, r, d = g.step(a)
        ref, rr, dr = o.step(a)
        c["dones"] += int(dr.sum())
        bad = _off(got, r, d, ref, rr, dr, tol)
        margin = o.tether_margin
        if margin is not None:
            c["near_surface"] += int((np.abs(margin) < SURFACE).sum())
            c["taut"] += int((margin >= 0).sum())
        if bad.any():
            explained = np.zeros(n, dtype=bool)
            if margin is not None:
                alt.state, alt.episode, alt.ever_done = prev, episode.copy(), ever.copy()
                alt.taut_override = np.where(bad, ~(margin >= 0), margin >= 0)
                ref2, rr2, dr2 = alt.step(a)
                alt.taut_override = None
                # window: on the surface (SURFACE), or within the drift the state tolerance itself allows (tol)
                explained = bad & ~_off(got, r, d, ref2, rr2, dr2, tol) & (np.abs(margin) < tol)
                c["branch_flips"] += int((explained & (np.abs(margin) < SURFACE)).sum())
                c["drift_flips"] += int((explained & ~(np.abs(margin) < SURFACE)).sum())
            rest = bad & ~explained
            done_flip = rest & (d != dr) & (o.done_margin < tol)
            c["done_flips"] += int(done_flip.sum())
            rest &= ~done_flip
            if rest.any():
                i = int(np.flatnonzero(rest)[0])
                raise AssertionError("%s step %d env %d: |err| %.3g, done %s vs %s, tether margin %s, done margin %.3g"
                                     % (env_id, t, i, np.abs(got[i] - ref[i]).max(), d[i], dr[i],
                                        None if margin is None else margin[i], o.done_margin[i]))
            g.resync(bad, o.obs(), o.episode, o.ever_done)
        good = ~bad
        c["worst"] = max(c["worst"], float((np.abs(got[good] - ref[good]) / np.maximum(1.0, np.abs(ref[good]))).max()))
    g.close()
    return c


def golden_trajectory(stepper_cls, env_id, g, tol=TOL):
    """tests/golden traj_*: the reference's own 400-step x 6-env run with reset-on-done, teacher-forced the way
    tests/test_oracle_golden.py does it for the oracle: every step starts from the reference's previous state
    (rounded to float32), and a step that misses the fixture must be the other tether branch of a state on the
    switching surface.  Returns (compared, flips)."""
    kind = KIND_OF_ID[env_id]
    T, E, _ = g["traj_a"].shape
    seed = int(g["traj_seed"])
    s = stepper_cls(env_id, E, seed)
    o = BatchOracle(kind, E, seed=seed, dtype=np.float64)
    assert np.array_equal(s.reset().astype(np.float64), g["traj_obs0"])
    o.reset()
    prev = g["traj_obs0"]
    flips = 0
    for t in range(T):
        a = g["traj_a"][t]
        ep, ev = o.episode.copy(), o.ever_done.copy()          # the fixture's bookkeeping before this step
        s.set_rows(prev)
        got, r, d = s.step(a)
        ref, rr, dr = g["traj_obs"][t], g["traj_r"][t], g["traj_done"][t]
        bad = _off(got, r, d, ref, rr, dr, tol)
        if bad.any():
            assert STATE_DIM[kind] in (9, 16), "only the slung-load kinds have a switching surface"
            # the reference arithmetic at this state with either branch: the vectorised float64 restatement, the
            # reference's own np.linalg.norm on 2-/3-vectors and the float32 state may each land on a different
            # side of |tether| = L(1 +- ulp), so the stepper must reproduce ONE of the two branch results
            o.state = prev.T.astype(np.float64)
            ref1, rr1, dr1 = o.step(a)
            margin = o.tether_margin.copy()
            assert (np.abs(margin[bad]) < SURFACE).all(), (t, margin[bad])
            o.state, o.episode, o.ever_done = prev.T.astype(np.float64), ep.copy(), ev.copy()
            o.taut_override = np.where(bad, ~(margin >= 0), margin >= 0)
            ref2, rr2, dr2 = o.step(a)
            o.taut_override = None
            still = bad & _off(got, r, d, ref1, rr1, dr1, tol) & _off(got, r, d, ref2, rr2, dr2, tol)
            assert not still.any(), (t, np.flatnonzero(still))
            flips += int(bad.sum())
        assert np.array_equal(d, dr), "step %d: done flags differ from the fixture" % t
        o.episode, o.ever_done = ep + dr.astype(np.uint32), ev | dr
        prev = ref
    s.close()
    return T * E, flips
