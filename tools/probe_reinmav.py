"""Developer probe: reinmav-v0 step time (50/51 float64 Euler substeps per step, compute-bound)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch
import gym_reinmav_b200 as rmv
for n in (1 << 16, 1 << 20):
    b = rmv.QuadBatch("reinmav-v0", n)
    rew = torch.empty(n, dtype=torch.float32, device="cuda:0")
    done = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        b.step(None, rew, done)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        b.step(None, rew, done)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("reinmav-v0 n=%d  %.1f us/step  %.3g env-steps/s  %.3g substeps/s" % (n, ms * 1e3, n / ms * 1e3, 50 * n / ms * 1e3))
    b.close()
