"""Developer probe: where the host time of one QuadVecEnv.step_tensor call goes (cProfile, small batch so the GPU is idle)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]
import torch
import gym_reinmav_b200 as rmv

for env_id in ("quadrotor3d-v0", "quadrotor2d-v0"):
    venv = rmv.make_vec_env(env_id, "native", 4096, 0)
    venv.reset()
    A = venv.action_space.shape[0]
    act = torch.zeros((4096, A), dtype=torch.float32, device="cuda:0")
    for _ in range(100):
        venv.step_tensor(act)
    torch.cuda.synchronize()
    K = 20000
    t0 = time.perf_counter()
    for _ in range(K):
        venv.step_tensor(act)
    torch.cuda.synchronize()
    print("%s: %.2f us per step_tensor call (host + device, N = 4096)" % (env_id, (time.perf_counter() - t0) / K * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(K):
        venv.step_tensor(act)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)
    venv.close()
