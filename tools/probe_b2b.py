"""Developer probe: per-launch time of the step kernel as a function of how many launches run back to back
(isolated launch vs sustained stream), with and without programmatic dependent launch (RMV_NO_PDL=1)."""
import argparse, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch
import gym_reinmav_b200 as rmv
from tools.sweep import BYTES, BOUNDS

ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="quadrotor2d-v0,quadrotor3d-v0")
ap.add_argument("--n", type=int, default=1 << 23)
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--ks", default="1,2,4,8,32,128,512")
ap.add_argument("--ring", type=int, default=8)
args = ap.parse_args()
dev = "cuda:0"
for env_id in args.envs.split(","):
    n = args.n
    b = rmv.QuadBatch(env_id, n, seed=0)
    b.set_tuning(vec=1, block=args.block)
    b.reset()
    A = b.action_dim
    lo, hi = BOUNDS[env_id]
    ring = args.ring
    acts = torch.empty((ring, n, A), dtype=torch.float32, device=dev)
    for t in range(ring):
        rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
    rew = torch.empty((ring, n), dtype=torch.float32, device=dev)
    done = torch.empty((ring, n), dtype=torch.uint8, device=dev)

    def run(k):
        left = k
        while left > 0:
            m = min(left, ring)
            b.step_many(m, acts, rew, done, out_ld=n, layout=0, fused=2)
            left -= m
    run(ring)
    torch.cuda.synchronize()
    for k in (int(x) for x in args.ks.split(",")):
        ts = []
        for rep in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(k); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / k)
        ts.sort()
        us = ts[len(ts) // 2]
        smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print("%-26s n=%d pdl=%s K=%4d  %8.2f us/launch (min %.2f)  %.0f GB/s alg   [%s]" % (
            env_id, n, "off" if os.environ.get("RMV_NO_PDL") else "on", k, us, ts[0], BYTES[env_id] * n / us / 1e3, smi), flush=True)
    b.close()
    del acts, rew, done
