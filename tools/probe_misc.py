"""Launches the kernels that are not reachable through tools/sweep.py a few times each, for ncu captures:
k_control, the fused control+step kernel, k_rollout, k_reinmav_step, the row-state kernel with Monitor on.

    python tools/probe_misc.py <control|control_step|rollout|rows_track|reinmav|vecnorm> [n_envs] [env_id]

vecnorm also prints the time of VecNormalize(QuadVecEnv).step_tensor against the bare venv.step_tensor.
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
elif what == "vecnorm":
    env_id = sys.argv[3] if len(sys.argv) > 3 else "quadrotor3d-v0"
    lo, hi = {"quadrotor3d-v0": (0, 10)}.get(env_id, (-10, 10))
    for wrap in (False, True):
        venv = rmv.make_vec_env(env_id, "native", n, 0)
        env = rmv.VecNormalize(venv) if wrap else venv
        env.reset_tensor() if wrap else venv.reset()
        A = venv.action_space.shape[0]
        acts = torch.empty((16, n, A), dtype=torch.float32, device=dev)
        for t in range(16):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
        for t in range(8):
            env.step_tensor(acts[t])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(128):
            env.step_tensor(acts[t % 16])
        e1.record()
        torch.cuda.synchronize()
        print("%s %s n=%d  %.2f us/step" % (env_id, "VecNormalize(venv).step_tensor" if wrap else "venv.step_tensor", n, e0.elapsed_time(e1) / 128 * 1e3))
        env.close()
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
