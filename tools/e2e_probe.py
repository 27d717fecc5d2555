"""e2e (host-buffer) step time vs number of pipeline chunks (developer tool)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import numpy as np, torch
import gym_reinmav_b200 as rmv
from gym_reinmav_b200.vec_env import _Pinned
n, A, S = 1 << 20, 4, 10
for obs_on in (True, False):
    for chunks in (1, 2, 4, 8, 16, 32, 64):
        b = rmv.QuadBatch("quadrotor3d-v0", n, seed=0)
        b.set_host_chunks(chunks)
        act, obs = _Pinned((n, A), "float32"), _Pinned((n, S), "float32")
        rew, done = _Pinned((n,), "float32"), _Pinned((n,), "uint8")
        act.array[...] = np.random.RandomState(0).uniform(0, 10, (n, A)).astype(np.float32)
        b.reset_host(obs.array)
        for _ in range(5):
            b.step_host(act.array, rew.array, done.array, obs.array if obs_on else None)
        t0 = time.perf_counter()
        K = 50
        for _ in range(K):
            b.step_host(act.array, rew.array, done.array, obs.array if obs_on else None)
        ms = (time.perf_counter() - t0) / K * 1e3
        print("obs=%d chunks=%3d  %.3f ms/step  %.3g env-steps/s" % (obs_on, chunks, ms, n / ms * 1e3), flush=True)
        b.close()
        for p in (act, obs, rew, done):
            p.free()
