"""Developer probe: host-buffer step through rmv_step_host (staged, chunked DMA) vs the step kernel reading
actions from / writing obs, reward, done to PINNED HOST memory directly (UVA zero-copy over PCIe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import numpy as np, torch
import gym_reinmav_b200 as rmv
from gym_reinmav_b200.vec_env import _Pinned

for env_id, A, S, lo, hi in (("quadrotor3d-v0", 4, 10, 0, 10), ("quadrotor2d-v0", 2, 5, -10, 10)):
    n = 1 << 20
    act, obs = _Pinned((n, A), "float32"), _Pinned((n, S), "float32")
    rew, done = _Pinned((n,), "float32"), _Pinned((n,), "uint8")
    act.array[...] = np.random.RandomState(0).uniform(lo, hi, (n, A)).astype(np.float32)
    d_act = torch.empty((n, A), dtype=torch.float32, device="cuda:0")
    res = {}
    for mode in ("staged", "zero_copy_all", "zero_copy_out", "staged"):
        b = rmv.QuadBatch(env_id, n, seed=0)
        b.reset_host(obs.array)

        def one():
            if mode == "staged":
                b.step_host(act.array, rew.array, done.array, obs.array)
            elif mode == "zero_copy_all":
                b.step(act.array, rew.array, done.array, obs.array)
                torch.cuda.current_stream().synchronize()
            else:
                d_act.copy_(torch.from_numpy(act.array), non_blocking=True)
                b.step(d_act, rew.array, done.array, obs.array)
                torch.cuda.current_stream().synchronize()
        for _ in range(5):
            one()
        K = 50
        t0 = time.perf_counter()
        for _ in range(K):
            one()
        ms = (time.perf_counter() - t0) / K * 1e3
        chk = (float(rew.array[:4096].sum()), int(done.array.sum()), float(obs.array[:64].sum()))
        print("%-16s %-14s %.3f ms/step  %.3g env-steps/s  check=%s" % (env_id, mode, ms, n / ms * 1e3, chk), flush=True)
        b.close()
    for p in (act, obs, rew, done):
        p.free()
