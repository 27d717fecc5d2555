"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kind, every CTA size, both
prefetch modes, obs rows + episode tracking on, ragged batch, the TMA variant, the fused rollout, the step stream (scalar
and packed pair), reset, the controllers (separate, fused with the step, inside the stream), a row-state handle, reinmav-v0,
the VecNormalize kernels (float4 and scalar walks)."""
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
# controllers, row-state handles, packed-pair stream
for env_id, (lo, hi) in BOUNDS.items():
    for layout in ("tiles", "rows"):
        if layout == "rows" and "3d" not in env_id:
            continue
        b = rmv.QuadBatch(env_id, n, seed=2, track_episodes=True, state_layout=layout)
        b.reset()
        A = b.action_dim
        act = torch.empty((4, n, A), dtype=torch.float32, device=dev)
        rew = torch.empty((4, n), dtype=torch.float32, device=dev)
        done = torch.empty((4, n), dtype=torch.uint8, device=dev)
        b.control(act[0]); b.step(act[0], rew[0], done[0])
        b.control_step(rew[1], done[1], actions_out=act[1])
        b.control_step_many(4, rew, done, out_ld=n, actions_out=act)
        b.step_many(4, act, rew, done, out_ld=n, layout=0, fused=1)      # rows: the rollout kernel on rows
        b.step_many(4, act, rew, done, out_ld=n, layout=0, fused=0)      # rows: the stream with the row transposes
        torch.cuda.synchronize()
        print(env_id, layout, "control paths ok, step_many ->", b.last_kernel, flush=True)
        b.close()
n2 = 128 * 9 + 64                                  # two envs per thread need even row offsets in the outputs
b = rmv.QuadBatch("quadrotor3d-v0", n2, seed=3)
b.reset()
act = torch.empty((8, n2, 4), dtype=torch.float32, device=dev)
for t in range(8):
    rmv.fill_uniform_actions(act[t], n2, 4, 0, 30, 1, step_index=t)
rew = torch.empty((8, n2), dtype=torch.float32, device=dev)
done = torch.empty((8, n2), dtype=torch.uint8, device=dev)
b.set_tuning(vec=2, block=128)                     # two envs per thread: the packed-pair stream kernel
b.step_many(8, act, rew, done, out_ld=n2, layout=0, fused=0)
torch.cuda.synchronize()
print("pair stream:", b.last_kernel, flush=True)
b.close()
# reinmav-v0
b = rmv.QuadBatch("reinmav-v0", n)
q = np.random.RandomState(3).normal(size=(n, 13)).astype(np.float32)
q[:, 6:10] /= np.linalg.norm(q[:, 6:10], axis=1, keepdims=True)
b.set_state(q)
r1 = torch.empty(n, dtype=torch.float32, device=dev)
d1 = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(2):
    b.step(None, r1, d1)
torch.cuda.synchronize()
b.close()
print("reinmav ok", flush=True)
# VecNormalize: float4 walk (aligned internal buffers) and the scalar walk (offset views through the C-ABI)
import ctypes
from gym_reinmav_b200._capi import check, lib
venv = rmv.VecNormalize(rmv.QuadVecEnv("quadrotor2d-slungload-v0", n, seed=4))
venv.reset_tensor()
a2 = torch.empty((n, 2), dtype=torch.float32, device=dev)
for t in range(3):
    rmv.fill_uniform_actions(a2, n, 2, -30, 30, 1, step_index=t)
    venv.step_tensor(a2)
venv.statistics()
venv.close()
h = ctypes.c_void_p()
S = 9
check(lib().rmv_norm_create(ctypes.byref(h), n, S, 0, 1, 1, 10.0, 10.0, 0.99, 1e-8))
buf_in = torch.randn(n * S + 4, device=dev)
buf_out = torch.zeros(n * S + 4, device=dev)
rw, ro = torch.randn(n, device=dev), torch.zeros(n, device=dev)
dn = (torch.rand(n, device=dev) < 0.2).to(torch.uint8)
for t in range(3):
    check(lib().rmv_norm_step(h, buf_in[1:].data_ptr(), rw.data_ptr(), dn.data_ptr(), buf_out[1:].data_ptr(), ro.data_ptr(), 1, None))
torch.cuda.synchronize()
check(lib().rmv_norm_destroy(h))
print("vecnormalize ok", flush=True)
print("sanitize probe finished")
