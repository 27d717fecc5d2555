"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kind, every CTA size, both
prefetch modes, obs rows + episode tracking on, ragged batch, the TMA variant, the fused rollout, reset, control."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import numpy as np
import torch
import gym_reinmav_b200 as rmv
from tools.sweep import BOUNDS

n, dev = 128 * 9 + 53, "cuda:0"
for env_id, (lo, hi) in BOUNDS.items():
    for block, pref, tma in ((128, 0, 0), (256, 1, 0), (512, 1, 0), (128, 0, 1)):
        b = rmv.QuadBatch(env_id, n, seed=1, track_episodes=True)
        b.set_tuning(vec=1, block=block); b.set_prefetch(pref); b.set_kernel(tma=bool(tma))
        b.reset()
        A, S = b.action_dim, b.state_dim
        # wide states: about a third of the envs finish on the very first step, so the re-draw pass runs in every CTA
        b.set_state(np.random.RandomState(7).uniform(-3, 3, (n, S)).astype(np.float32))
        act = torch.empty((8, n, A), dtype=torch.float32, device=dev)
        for t in range(8):
            rmv.fill_uniform_actions(act[t], n, A, 3 * lo, 3 * hi, 1, step_index=t)      # wide actions: many resets
        rew = torch.empty((8, n), dtype=torch.float32, device=dev)
        done = torch.empty((8, n), dtype=torch.uint8, device=dev)
        obs = torch.empty((n, S), dtype=torch.float32, device=dev)
        for t in range(8):
            b.step(act[t], rew[t], done[t], obs=obs if t % 2 else None)
        b.step_many(8, act, rew, done, out_ld=n, layout=0, fused=True)
        b.step_many(8, act, rew, done, out_ld=n, layout=0, fused=2)
        for cap in (0, 1):
            b.set_stream(warps_per_sm=cap)
            b.step_many(8, act, rew, done, out_ld=n, layout=0, fused=0)
        b.get_state(); b.episode_stats()
        torch.cuda.synchronize()
        print(env_id, block, pref, tma, "dones in last ring:", int(done.sum()), flush=True)
        b.close()
print("sanitize probe finished")
