"""Developer probe: ONE step launch with a cold L2 (256 MiB written before every launch, one event pair per
launch) — the regime of an RL loop whose policy network flushes L2 between env steps.  Compares launch shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch
import gym_reinmav_b200 as rmv
from tools.sweep import BYTES, BOUNDS

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for env_id in ("quadrotor3d-v0", "quadrotor2d-v0"):
    for n in (1 << 20, 1 << 17):
        lo, hi = BOUNDS[env_id]
        for name, kw in (("default", {}), ("block256", dict(block=256)), ("block512", dict(block=512)), ("prefetch1", dict(pref=1)),
                         ("prefetch0", dict(pref=0)), ("vec2", dict(vec=2)), ("vec4", dict(vec=4)),
                         ("tma", dict(tma=1)), ("tma_8ctas", dict(tma=1, ctas=8)), ("tma_4ctas", dict(tma=1, ctas=4))):
            b = rmv.QuadBatch(env_id, n, seed=0)
            b.set_tuning(vec=kw.get("vec", 1), block=kw.get("block", 0))
            b.set_prefetch(kw.get("pref", -1))
            b.set_kernel(tma=bool(kw.get("tma", 0)), ctas_per_sm=kw.get("ctas", 0))
            b.reset()
            A = b.action_dim
            acts = torch.empty((4, n, A), dtype=torch.float32, device=dev)
            for t in range(4):
                rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
            rew = torch.empty((4, n), dtype=torch.float32, device=dev)
            done = torch.empty((4, n), dtype=torch.uint8, device=dev)
            for t in range(4):
                b.step(acts[t], rew[t], done[t])
            evs = []
            for t in range(30):
                flush.fill_(t & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); b.step(acts[t % 4], rew[t % 4], done[t % 4]); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(z) * 1e3 for a, z in evs)
            us = ts[len(ts) // 2]
            print("%-16s n=%-8d %-10s %7.2f us (min %.2f)  %.0f GB/s alg" % (env_id, n, name, us, ts[0], BYTES[env_id] * n / us / 1e3), flush=True)
            b.close()
