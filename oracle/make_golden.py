"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN SOURCE (T0-ref).

Build-container only: needs /root/reference (read-only).  The fixtures travel
to the GPU box; the reference does not.  Run:  python -m oracle.make_golden

Per env id the file holds

  kat_s0 [K,S], kat_a [K,A]  ->  kat_s1 [K,S], kat_r [K], kat_done [K]
      single reference ``step()`` calls from scripted states: reset-like draws,
      wide draws (x3, to trip both done clauses) and, for the slung-load kinds,
      states on both sides of the taut/slack switch.
  traj_a [T,E,A], traj_obs [T,E,S], traj_r [T,E], traj_done [T,E], traj_obs0 [E,S]
      E reference env objects stepped T times with U(action_space) actions and
      reset-on-done the way a VecEnv does it (obs after a done is the reset obs),
      their ``np_random`` replaced by the Philox stream (seed, env_id, episode)
      so that resets are reproducible lane-parallel.  Pins ``steps_beyond_done``
      (first done pays +1.0, later ones 0.0) and the reset semantics.
  ctrl_a [T,A], ctrl_obs [T,S], ctrl_r [T], ctrl_done [T], ctrl_obs0 [S]
      the reference's own test loop (test/test_quadrotor3d.py:12-22 and twins):
      ``control()`` -> ``step()`` -> reset on done, 400 steps, no render.

quadrotor2d-v0 is stored for both repairs of its unparseable done expression
(quadrotor2d.py:95-98): ``quadrotor2d-v0.npz`` (A) and
``quadrotor2d-v0.repairB.npz``.
"""
import os
import sys

import numpy as np

from . import philox, refshim
from .batch import ACTION_BOUNDS, ACTION_DIM, KIND_OF_ID, STATE_DIM

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 20190312           # any fixed number
K, T, E = 192, 400, 6


def _scripted(env, seed, env_id, kind):
    ep = {"n": 0}

    def draws(n):
        ep["n"] += 1
        return philox.uniform_f32(seed, [env_id], [ep["n"]], philox.STREAM_RESET_BASE + kind, n, -1.0, 1.0)[:, 0]

    env.np_random = refshim.ScriptedUniform(draws)


def _kat(cls, kind, rng):
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    s0 = rng.uniform(-1, 1, size=(K, S))
    s0[K // 3:2 * K // 3] *= 3.0
    if S in (9, 16):                      # push a third of the loads near / beyond the tether length
        d = 2 if S == 9 else 3
        lp = slice(5, 7) if S == 9 else slice(10, 13)
        tether = 0.5 if S == 9 else 1.5
        dirs = rng.normal(size=(K // 3, d))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        radius = tether * rng.uniform(0.8, 1.3, size=(K // 3, 1))
        s0[2 * K // 3:, lp] = s0[2 * K // 3:, 0:d] + dirs * radius
    # values exactly representable in float32, so the f32 path sees identical inputs
    s0 = s0.astype(np.float32).astype(np.float64)
    a = rng.uniform(lo, hi, size=(K, A)).astype(np.float32).astype(np.float64)
    s1 = np.empty_like(s0)
    r = np.empty(K)
    done = np.empty(K, dtype=bool)
    for i in range(K):
        env = cls()                        # fresh object: steps_beyond_done is None
        env.state = s0[i].copy()
        o, r[i], done[i], _ = env.step(a[i].copy())
        s1[i] = np.asarray(o, dtype=float).reshape(-1)
    return dict(kat_s0=s0, kat_a=a, kat_s1=s1, kat_r=r, kat_done=done)


def _traj(cls, kind, rng):
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    lo, hi = ACTION_BOUNDS[kind]
    acts = rng.uniform(lo, hi, size=(T, E, A)).astype(np.float32).astype(np.float64)
    obs = np.empty((T, E, S))
    rew = np.empty((T, E))
    done = np.empty((T, E), dtype=bool)
    obs0 = np.empty((E, S))
    for e in range(E):
        env = cls()
        _scripted(env, SEED, e, kind)
        obs0[e] = env.reset()
        for t in range(T):
            o, rew[t, e], done[t, e], _ = env.step(acts[t, e].copy())
            if done[t, e]:
                o = env.reset()
            obs[t, e] = np.asarray(o, dtype=float).reshape(-1)
    return dict(traj_a=acts, traj_obs=obs, traj_r=rew, traj_done=done, traj_obs0=obs0,
                traj_seed=np.int64(SEED))


def _ctrl(cls, kind):
    S, A = STATE_DIM[kind], ACTION_DIM[kind]
    env = cls()
    _scripted(env, SEED, 1000, kind)
    obs0 = env.reset()
    acts, obs, rew, done = np.empty((T, A)), np.empty((T, S)), np.empty(T), np.empty(T, dtype=bool)
    for t in range(T):
        acts[t] = np.asarray(env.control(), dtype=float).reshape(-1)
        o, rew[t], done[t], _ = env.step(acts[t].copy())
        if done[t]:
            o = env.reset()
        obs[t] = np.asarray(o, dtype=float).reshape(-1)
    return dict(ctrl_a=acts, ctrl_obs=obs, ctrl_r=rew, ctrl_done=done, ctrl_obs0=obs0)


def main():
    if not refshim.available():
        sys.exit("reference tree not found; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    jobs = [(eid, "A", eid + ".npz") for eid in KIND_OF_ID] + [("quadrotor2d-v0", "B", "quadrotor2d-v0.repairB.npz")]
    for eid, repair, fname in jobs:
        kind = KIND_OF_ID[eid]
        cls = refshim.load_class(eid, repair)
        rng = np.random.RandomState(1234 + kind)
        data = {}
        data.update(_kat(cls, kind, rng))
        data.update(_traj(cls, kind, rng))
        data.update(_ctrl(cls, kind))
        np.savez_compressed(os.path.join(OUT, fname), **data)
        print("%-34s kat done=%d/%d  traj done=%d/%d  ctrl done=%d/%d" % (
            fname, data["kat_done"].sum(), K, data["traj_done"].sum(), T * E, data["ctrl_done"].sum(), T))


if __name__ == "__main__":
    main()
