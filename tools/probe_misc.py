"""Launches the kernels that are not reachable through tools/sweep.py a few times each, for ncu captures:
k_control, the fused control+step kernel, k_rollout, k_reinmav_step, the row-state kernel with Monitor on.

    python tools/probe_misc.py <control|control_step|rollout|rows_track|reinmav> [n_envs] [env_id]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch  # noqa: E402

import gym_reinmav_b200 as rmv  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dev = "cuda:0"
if what == "reinmav":
    b = rmv.QuadBatch("reinmav-v0", n)
    rew = torch.empty(n, dtype=torch.float32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(6):
        b.step(None, rew, done)
else:
    env_id = sys.argv[3] if len(sys.argv) > 3 else "quadrotor3d-v0"
    rows = what == "rows_track"
    b = rmv.QuadBatch(env_id, n, seed=0, track_episodes=rows, state_layout="rows" if rows else "tiles")
    b.reset()
    A = b.action_dim
    lo, hi = {"quadrotor3d-v0": (0, 10)}.get(env_id, (-10, 10))
    T = 16
    acts = torch.empty((T, n, A), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
    rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)
    if what == "control":
        for t in range(6):
            b.control(acts[t])
            b.step(acts[t], rew[t], done[t])
    elif what == "control_step":
        for t in range(6):
            b.control_step(rew[t], done[t])
    elif what == "rollout":
        for _ in range(4):
            b.step_many(T, acts, rew, done, out_ld=n, layout=0, fused=1)
    elif what == "rows_track":
        for t in range(12):
            b.step(acts[t], rew[t], done[t])
    else:
        raise SystemExit("unknown probe %r" % what)
torch.cuda.synchronize()
print("done", what)
