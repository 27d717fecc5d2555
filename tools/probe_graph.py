"""Developer probe: K single-step launches issued from the C loop vs the same K launches captured once in a CUDA
graph and replayed (small batches are launch-bound)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch
import gym_reinmav_b200 as rmv
from tools.sweep import BYTES, BOUNDS

dev = "cuda:0"
K = 64
for env_id in ("quadrotor3d-v0", "quadrotor2d-v0"):
    for n in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
        lo, hi = BOUNDS[env_id]
        b = rmv.QuadBatch(env_id, n, seed=0)
        b.reset()
        A = b.action_dim
        acts = torch.empty((K, n, A), dtype=torch.float32, device=dev)
        for t in range(K):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
        rew = torch.empty((K, n), dtype=torch.float32, device=dev)
        done = torch.empty((K, n), dtype=torch.uint8, device=dev)

        def run():
            b.step_many(K, acts, rew, done, out_ld=n, layout=0, fused=2)
        run(); torch.cuda.synchronize()

        def timeit(fn, reps=20):
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * K))
            return best
        t_stream = timeit(run)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                run()
        torch.cuda.synchronize()
        t_graph = timeit(g.replay)
        os.environ["X"] = "1"
        print("%-16s n=%-8d stream %6.2f us/step   graph %6.2f us/step   fused-rollout " % (env_id, n, t_stream, t_graph), end="")
        def runf():
            b.step_many(K, acts, rew, done, out_ld=n, layout=0, fused=True)
        print("%6.2f us/step" % timeit(runf), flush=True)
        b.close()
