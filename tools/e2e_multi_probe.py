"""Host-buffer (e2e) ceiling of a multi-GPU box, and where rmv_step_host sits under it.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_multi_probe.py

Every rank owns one GPU and moves, per step, what the e2e step of bench.py moves: 16.8 MB pinned host -> device
(actions of 2^20 quadrotor3d envs) and 47.2 MB device -> host (obs + reward + done), both directions at once.

  ceiling   plain cudaMemcpyAsync of those bytes on two streams, all ranks at the same time (barrier), no kernel
  step      rmv_step_host (H2D | kernel | D2H pipelined over chunks), for several chunk counts
  pinned    per-rank CPU affinity (a disjoint slice of the cores) on / off

Rank 0 prints one JSON line per measurement; times are the max over ranks (CUDA events / perf_counter around the
loop, barrier on both sides).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gym_reinmav_b200 as rmv  # noqa: E402
from gym_reinmav_b200.vec_env import _Pinned  # noqa: E402

N, A, S = 1 << 20, 4, 10
H2D, D2H = N * A * 4, N * (4 * S + 4 + 1)


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def worst(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(**kw):
        if rank == 0:
            kw.update(n_gpus=world, cores=len(os.sched_getaffinity(0)))
            print(json.dumps(kw), flush=True)

    all_cores = sorted(os.sched_getaffinity(0))
    for pin in (False, True):
        if pin:
            per = max(1, len(all_cores) // world)
            os.sched_setaffinity(0, all_cores[local * per:(local + 1) * per] or all_cores)
        else:
            os.sched_setaffinity(0, all_cores)
        # ---- ceiling: the bytes of one e2e step, both directions at once, nothing else
        h_in, h_out = _Pinned((H2D,), np.uint8), _Pinned((D2H,), np.uint8)
        t_in, t_out = torch.from_numpy(h_in.array), torch.from_numpy(h_out.array)
        d_in, d_out = torch.empty(H2D, dtype=torch.uint8, device=dev), torch.empty(D2H, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def copies(k, both=True, up=True, down=True):
            for _ in range(k):
                if up:
                    with torch.cuda.stream(s1):
                        d_in.copy_(t_in, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        t_out.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()

        for name, kw in (("h2d_only", dict(down=False)), ("d2h_only", dict(up=False)), ("both", dict())):
            copies(3, **kw)
            barrier()
            t0 = time.perf_counter()
            copies(30, **kw)
            ms = worst((time.perf_counter() - t0) * 1e3 / 30)
            barrier()
            emit(what="ceiling", dirs=name, cpu_affinity=pin, ms_per_step=ms,
                 per_rank_h2d_GBps=(H2D / ms / 1e6) if kw.get("up", True) else 0.0,
                 per_rank_d2h_GBps=(D2H / ms / 1e6) if kw.get("down", True) else 0.0,
                 aggregate_GBps=world * ((H2D if kw.get("up", True) else 0) + (D2H if kw.get("down", True) else 0)) / ms / 1e6)
        h_in.free(); h_out.free()
        del t_in, t_out, d_in, d_out

        # ---- rmv_step_host
        for layout, chunks in (("tiles", 0), ("tiles", 1), ("tiles", 4), ("tiles", 32), ("rows", 0)):
            if True:
                b = rmv.QuadBatch("quadrotor3d-v0", N, device=local, seed=0, env_id_offset=rank * N, state_layout=layout)
                b.set_host_chunks(chunks)
                act, obs = _Pinned((N, A), "float32"), _Pinned((N, S), "float32")
                rew, done = _Pinned((N,), "float32"), _Pinned((N,), "uint8")
                act.array[...] = np.random.RandomState(rank).uniform(0, 10, (N, A)).astype(np.float32)
                b.reset_host(obs.array)
                for _ in range(3):
                    b.step_host(act.array, rew.array, done.array, obs.array)
                barrier()
                t0 = time.perf_counter()
                for _ in range(30):
                    b.step_host(act.array, rew.array, done.array, obs.array)
                ms = worst((time.perf_counter() - t0) * 1e3 / 30)
                barrier()
                emit(what="rmv_step_host", state_layout=layout, chunks=chunks, cpu_affinity=pin, ms_per_step=ms,
                     env_steps_per_s=world * N / ms * 1e3, per_rank_h2d_GBps=H2D / ms / 1e6, per_rank_d2h_GBps=D2H / ms / 1e6,
                     aggregate_GBps=world * (H2D + D2H) / ms / 1e6)
                b.close()
                for p in (act, obs, rew, done):
                    p.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
