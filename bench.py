#!/usr/bin/env python
"""bench.py — env-steps/s of the batched quadrotor3d-v0 step() on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                          (the reference CPU arm)

A "step" is ONE pass of the hot path (the fused step: dynamics + reward + done + reset-on-done, state read from and
written back to the state buffer) over one batch of N_ENVS = 2^20 quadrotor3d-v0 envs per GPU with random-policy
actions (BASELINE.json metric / configs[1] scaled to the metric's N = 2^20).  K steps are issued with ONE call of
rmv_step_many(RMV_MANY_STREAM): the persistent step-stream kernel (csrc/rmv_stream.cuh).

* value     whole-job env-steps/s, inputs resident in HBM when the timed region starts
            (state in the handle, actions in a 64-step ring written by the Philox fill kernel); the K-step block is
            repeated until the timed region is >= 60 ms.
* e2e       the same metric through the C-ABI call with HOST buffers (rmv_step_host: pinned
            host actions -> H2D -> kernel -> D2H of obs/reward/done every step).
* roofline  algorithmic bytes (101 B per env-step) / CUDA-event time per launch,
            against MEASURED_PEAKS.json hbm_gbs; the dict also carries the points that qualify the L2-assisted headline:
            hbm_bound (N = 2^23), cold_l2, the N = 65536 configs, and the same work as one launch per step.
* cpu_baseline  oracle/scalar.py (one Python object per env, the reference's execution model) on
            all host cores, a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]

ENV_ID = "quadrotor3d-v0"
N_ENVS = 1 << 20                 # per GPU (weak scaling)
STATE_DIM, ACTION_DIM = 10, 4
ALG_BYTES = 4 * (2 * STATE_DIM + ACTION_DIM + 1) + 1      # 101 B per env-step
ACT_LO, ACT_HI = 0.0, 10.0
RING = 64                        # pre-generated action steps (64 x 16 MiB = 1 GiB > L2)
SEED = 0
METRIC = "env-steps/sec quadrotor3d-v0 at N=2^20 per GPU"
UNIT = "env-steps/s"


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One worker process = a DummyVecEnv of `envs` reference-style env objects stepped `steps` times."""
    worker, envs, steps, warmup = args
    import numpy as np
    from oracle.scalar import ScalarEnv
    objs = [ScalarEnv(ENV_ID, seed=SEED, env_id=worker * envs + i) for i in range(envs)]
    for o in objs:
        o.reset()
    rng = np.random.RandomState(1000 + worker)
    acts = rng.uniform(ACT_LO, ACT_HI, size=(steps + warmup, envs, ACTION_DIM))

    def run(t0, t1):
        for t in range(t0, t1):
            for i, o in enumerate(objs):
                _, _, done = o.step(acts[t, i])
                if done:
                    o.reset()

    run(0, warmup)
    begin = time.perf_counter()
    run(warmup, warmup + steps)
    return time.perf_counter() - begin


def cpu_reference(steps, warmup, target_seconds=15.0, per_core_rate=15000.0):
    """Times oracle/scalar.py with every host core (process-per-worker, the reference's own scaling
    model: baselines SubprocVecEnv behind gym_reinmav/run.py:74-89)."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    envs = int(max(1, min(4096, round(target_seconds * per_core_rate / max(1, steps + warmup)))))
    with mp.get_context("fork").Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(w, envs, steps, warmup) for w in range(cores)])
    worst = max(times)
    total = cores * envs * steps
    return {"value": total / worst, "seconds": worst, "cores": cores, "envs_per_core": envs, "steps": steps,
            "sample": "%d processes x %d env objects x %d steps (+%d warm-up), %s, actions ~ U(action_space), "
                      "reset-on-done, float64 NumPy per env (oracle/scalar.py)" % (cores, envs, steps, warmup, ENV_ID)}


def cpu_vectorised(seconds=3.0, n=1 << 16):
    """A STRONGER CPU baseline than the reference's execution model, clearly not the reference: the vectorised float64
    NumPy restatement (oracle/batch.py) stepping N = 65536 envs per call on one core, reset-on-done included."""
    import numpy as np
    from oracle.batch import BatchOracle
    o = BatchOracle(ENV_ID, n, seed=SEED, dtype=np.float64)
    o.reset()
    acts = np.random.RandomState(7).uniform(ACT_LO, ACT_HI, size=(8, n, ACTION_DIM))
    o.step(acts[0])
    k, begin = 0, time.perf_counter()
    while time.perf_counter() - begin < seconds:
        o.step(acts[k % 8])
        k += 1
    dt = time.perf_counter() - begin
    return {"value": n * k / dt, "unit": UNIT, "cores": 1, "kind": "vectorised float64 NumPy restatement (not the reference)",
            "sample": "%d calls x %d envs in %.1f s, one core" % (k, n, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s, %d env objects per step over %d host cores (bounded sample of N=2^20)"
                               % (ENV_ID, res["cores"] * res["envs_per_core"], res["cores"]),
                   "env_id": ENV_ID, "envs_per_step": res["cores"] * res["envs_per_core"]},
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python and does not import here (SURVEY.md R8/R9); this arm runs the "
                "oracle port that reproduces its trajectories bit for bit (tests/test_oracle_golden.py)",
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, n_gpus):
        self.path = tempfile.mktemp(prefix="rmv_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", ",".join(str(i) for i in range(n_gpus))],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons, busy, watts = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx, util = float(f[1]), float(f[2]), float(f[4])
            except ValueError:
                continue
            sm.append(clk); smax.append(mx); busy.append(util > 0)
            try:
                watts.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        under = [c for c, b in zip(sm, busy) if b] or sm
        if under:
            under.sort()
            out.update(sm_mhz=under[len(under) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(under),
                       sm_min_mhz=under[0], power_w_max=max(watts) if watts else None)
        return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import gym_reinmav_b200 as rmv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
        args.gpus = world
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, A, S = args.envs, ACTION_DIM, STATE_DIM
    batch = rmv.QuadBatch(ENV_ID, n, device=local, seed=SEED, env_id_offset=rank * n)   # global env ids: shard-invariant
    batch.set_tuning(vec=args.vec, block=args.block)
    batch.reset()
    ring = max(2, min(RING, args.steps + args.warmup))
    acts = torch.empty((ring, n, A), dtype=torch.float32, device=dev)
    for t in range(ring):
        rmv.fill_uniform_actions(acts[t], n, A, ACT_LO, ACT_HI, SEED, env_id_offset=rank * n, step_index=t, device=local)
    rew = torch.empty((ring, n), dtype=torch.float32, device=dev)
    done = torch.empty((ring, n), dtype=torch.uint8, device=dev)

    def run_steps(k):
        """k single-step launches (state streams through HBM every step), actions from the ring."""
        left = k
        while left > 0:
            m = min(left, ring)
            batch.step_many(m, acts, rew, done, out_ld=n, layout=0, fused=False)
            left -= m

    sampler = ClockSampler(world) if rank == 0 else None

    # ---- value: device-resident inputs.  The driver asks for K = 20 steps (0.2 ms): a window that short is launch
    # skew, not throughput, so the K-step block is repeated back to back until the timed region is >= 60 ms — one pair
    # of events around the whole region (value), one event after every block (best block, reported beside it).
    run_steps(args.warmup)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); run_steps(args.steps); run_steps(args.steps); c1.record()
    torch.cuda.synchronize()
    est_block_ms = max_over_ranks(c0.elapsed_time(c1) / 2.0)
    blocks = int(max(args.repeats, min(20000, -(-args.min_region_ms // max(est_block_ms, 1e-3)))))
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
    launches0 = batch.launches
    marks[0].record()
    for i in range(blocks):
        run_steps(args.steps)
        marks[i + 1].record()
    barrier()
    launches = batch.launches - launches0
    headline_kernel = batch.last_kernel
    region_ms = max_over_ranks(marks[0].elapsed_time(marks[blocks]))
    best_block_ms = max_over_ranks(min(marks[i].elapsed_time(marks[i + 1]) for i in range(blocks)))
    ms_per_step = region_ms / (blocks * args.steps)
    value = world * n * args.steps * blocks / (region_ms * 1e-3)

    # ---- the same steps as one launch per step (RMV_MANY_LAUNCHES: what the stream kernel replaces)
    def run_launches(k):
        left = k
        while left > 0:
            m = min(left, ring)
            batch.step_many(m, acts, rew, done, out_ld=n, layout=0, fused=2)
            left -= m
    run_launches(2 * ring)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_l = max(args.steps, 200)
    c0.record(); run_launches(k_l); c1.record()
    barrier()
    launch_per_step_us = max_over_ranks(c0.elapsed_time(c1)) / k_l * 1e3

    # ---- e2e: host buffers through rmv_step_host (H2D + kernel + D2H every step)
    from gym_reinmav_b200.vec_env import _Pinned
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_ring = 4
    h_act = [_Pinned((n, A), "float32") for _ in range(h_ring)]
    for i, h in enumerate(h_act):
        h.array[...] = acts[i % ring].cpu().numpy()
    h_obs, h_rew, h_done = _Pinned((n, S), "float32"), _Pinned((n,), "float32"), _Pinned((n,), "uint8")
    for t in range(3):
        batch.step_host(h_act[t % h_ring].array, h_rew.array, h_done.array, h_obs.array)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(e2e_steps):
        batch.step_host(h_act[t % h_ring].array, h_rew.array, h_done.array, h_obs.array)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0) if world == 1 else 0.0))
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)
    checksum = float(h_rew.array[:1024].sum())            # the step's result is read on the host

    # ---- episode statistics through the collective the path has (logging only, never on the step path):
    # a tracked handle per rank, all_reduce_episode_stats over NCCL, checked against an all_gather of the per-rank numbers
    stats = measure_episode_stats(torch, dist, rmv, local, rank, world, dev)

    points, extras = {}, {}
    if world == 1:
        try:                                   # side measurements must never cost the contract line
            points = measure_roofline_points(torch, rmv, args, local)
        except Exception as exc:               # noqa: BLE001
            points = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.synchronize()
    if args.extras and world == 1:
        try:
            extras = measure_extras(torch, rmv, args, local)
        except Exception as exc:               # noqa: BLE001
            extras = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.synchronize()

    clocks = sampler.stop() if sampler else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        res = cpu_reference(args.cpu_steps, 20)
        cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]}
        try:
            cpu["vectorised_numpy"] = cpu_vectorised()
        except Exception as exc:               # noqa: BLE001
            cpu["vectorised_numpy"] = {"error": str(exc)}

    if rank == 0:
        peaks, peak_src = load_peak()
        achieved = ALG_BYTES * n / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s N=%d envs per GPU, random-policy actions U[0,10)^4, reset-on-done in-kernel"
                                   % (ENV_ID, n),
                       "env_id": ENV_ID, "envs_per_gpu": n, "state_dim": S, "action_dim": A,
                       "launch": "rmv_step_many(RMV_MANY_STREAM): ONE launch of the persistent step-stream kernel per K-step block; every "
                                 "step reads and writes the state in place (no state kept in registers across steps), actions from a "
                                 "%d-step ring; roofline.launch_per_step is the same work as one launch per step" % ring,
                       "l2": "no flush: per-step traffic %.0f MB (state %.0f MB read+written in place, re-read next step; "
                             "actions stream from a %.2f GiB ring > 126 MB L2); roofline.hbm_bound is the same kernel with a "
                             "working set that cannot be L2-resident" % (ALG_BYTES * n / 1e6, 4 * S * n / 1e6, ring * n * A * 4 / 2 ** 30),
                       "timing": "CUDA events on the launch stream around %d back-to-back blocks of exactly K = %d steps (%.1f ms "
                                 "region), max over ranks; best single block %.4f ms/step" % (blocks, args.steps, region_ms,
                                                                                              best_block_ms / args.steps)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * A * 4, "d2h_bytes_per_step": n * (4 * S + 4 + 1),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "per_rank": {"h2d_GBps": n * A * 4 / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                                 "d2h_GBps": n * (4 * S + 4 + 1) / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                                 "note": "both directions overlap inside rmv_step_host; rates = bytes per step / step time of the slowest rank"},
                    "path": "rmv_step_host (C-ABI, pinned host buffers): H2D actions, step kernel with obs rows, D2H obs+reward+done, stream sync"},
            "gpu_launches": launches * world,
            "roofline": dict({
                "bound": "hbm", "achieved": achieved, "peak": peaks, "unit": "GB/s", "frac": achieved / peaks,
                "traffic": (load_traffic("k_step_stream_quadrotor3d_n2pow20") or {}).get("dram_bytes_per_step"),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, divided by "
                                "the steps of the captured launch; a CITATION of profiles/traffic.json, not measured in this run",
                "kernel": "rmv::%s<QUAD3D> (%s)" % (headline_kernel, "persistent; one launch = K steps of N envs"
                                                     if "stream" in headline_kernel else "one launch per step"),
                "steps_per_launch": min(args.steps, ring),
                "algorithmic_bytes_per_step": ALG_BYTES * n,
                "algorithmic_bytes_per_launch": ALG_BYTES * n * min(args.steps, ring),
                "avg_launch_us": ms_per_step * 1e3 * min(args.steps, ring), "peak_source": peak_src,
                "l2_resident": True,
                "note": "frac > 1 at N=2^20 is NOT an HBM fraction: the 42 MB state stays in the 126 MB L2 between steps, only "
                        "actions / rewards / done flags (22 MB per step) move through DRAM.  hbm_bound is the same kernel at "
                        "N=2^23 (847 MB per step, cannot be L2-resident); cold_l2 is one isolated N=2^20 launch after an L2 flush",
                "launch_per_step": {"us_per_step": launch_per_step_us, "frac": ALG_BYTES * n / (launch_per_step_us * 1e-6) / 1e9 / peaks,
                                    "kernel": "rmv::k_step<QUAD3D> (one launch per step, PDL + CUDA-graph replay)"},
            }, **points),
            "episode_stats": stats,
            "cpu_baseline": cpu,
            "extras": extras,
            "result_checksum": checksum,
        }
        emit(line)
    for h in h_act + [h_obs, h_rew, h_done]:
        h.free()
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_episode_stats(torch, dist, rmv, local, rank, world, dev):
    """The only collective of the path: episode statistics for logging.  A tracked handle per rank steps a small batch,
    `dist.all_reduce_episode_stats` sums {episodes, return_sum, length_sum, env_steps} over the ranks (NCCL at N > 1),
    and the sum is checked against an all_gather of the per-rank numbers."""
    import gym_reinmav_b200.dist as rdist
    n, T = 1 << 16, 256
    venv = rmv.QuadVecEnv(ENV_ID, n, device=local, seed=SEED, env_id_offset=(1 << 40) + rank * n, monitor=True)
    venv.batch.reset()
    acts = torch.empty((T, n, ACTION_DIM), dtype=torch.float32, device=dev)
    for t in range(T):
        rmv.fill_uniform_actions(acts[t], n, ACTION_DIM, ACT_LO, ACT_HI, SEED, env_id_offset=rank * n, step_index=t, device=local)
    venv.batch.step_many(T, acts, None, None, layout=0, fused=0)
    mine = venv.episode_stats()
    torch.cuda.synchronize()
    rdist.all_reduce_episode_stats(venv)                       # warm-up (NCCL channel setup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    total = rdist.all_reduce_episode_stats(venv)
    torch.cuda.synchronize()
    us = 1e6 * (time.perf_counter() - t0)
    keys = ("episodes", "return_sum", "length_sum", "env_steps")
    ok = True
    if world > 1:
        t = torch.tensor([mine[k] for k in keys], dtype=torch.float64, device=dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        summed = torch.stack(parts).sum(0).tolist()
        ok = all(abs(total[k] - v) <= 1e-9 * max(1.0, abs(v)) for k, v in zip(keys, summed))
    venv.close()
    return {"episodes": total["episodes"], "mean_return": total["mean_return"], "mean_length": total["mean_length"],
            "env_steps": total["env_steps"], "allreduce_us": us, "sum_equals_per_rank_sums": ok,
            "backend": "nccl all_reduce (4 x float64)" if world > 1 else "single rank (no collective)",
            "sample": "%d tracked envs per rank x %d steps through the step stream" % (n, T)}


def _time_many(torch, batch, acts, rew, done, n, k, mode):
    ring = acts.shape[0]

    def run(kk):
        left = kk
        while left > 0:
            m = min(left, ring)
            batch.step_many(m, acts, rew, done, out_ld=n, layout=0, fused=mode)
            left -= m
    run(2 * ring)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(k); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / k)
    return best * 1e3            # us per step


def measure_roofline_points(torch, rmv, args, local):
    """The points that qualify the headline fraction, measured in the same run and carried INSIDE the roofline dict:
    the HBM-bound size (N = 2^23), a cold-L2 isolated launch at N = 2^20, and BASELINE configs 1-2 (N = 65536)."""
    dev = "cuda:0" if local == 0 else "cuda:%d" % local
    peak, _ = load_peak()
    out = {}
    bounds = {ENV_ID: (0, 10), "quadrotor3d-slungload-v0": (-10, 10)}
    bytes_of = {ENV_ID: 101, "quadrotor3d-slungload-v0": 149}

    def make(env_id, n, ring):
        b = rmv.QuadBatch(env_id, n, device=local, seed=SEED)
        b.reset()
        lo, hi = bounds[env_id]
        acts = torch.empty((ring, n, 4), dtype=torch.float32, device=dev)
        for t in range(ring):
            rmv.fill_uniform_actions(acts[t], n, 4, lo, hi, SEED, step_index=t, device=local)
        return b, acts, torch.empty((ring, n), dtype=torch.float32, device=dev), torch.empty((ring, n), dtype=torch.uint8, device=dev)

    def entry(env_id, n, us, mode, kernel=None):
        gbs = n * bytes_of[env_id] / (us * 1e-6) / 1e9
        return {"env_id": env_id, "n_envs": n, "us_per_step": us, "env_steps_per_s": n / (us * 1e-6), "achieved": gbs,
                "frac": gbs / peak, "mode": mode, "kernel": kernel}

    # HBM-bound: 847 MB per step
    n = 1 << 23
    b, acts, rew, done = make(ENV_ID, n, 8)
    us = _time_many(torch, b, acts, rew, done, n, 96, 0)
    out["hbm_bound"] = entry(ENV_ID, n, us, "rmv_step_many(RMV_MANY_STREAM), 8-step calls: at this size the library issues one launch per "
                             "step (rmv_stream_pays)", b.last_kernel)
    out["hbm_bound"]["dram_bytes_per_step_ncu"] = (load_traffic("k_step_quadrotor3d_n2pow23") or {}).get("dram_bytes_per_step")
    b.close()
    del acts, rew, done

    # cold L2: 256 MiB are read between launches (> 126 MB L2, leaves it clean), each single-step launch timed alone
    n = 1 << 20
    b, acts, rew, done = make(ENV_ID, n, 8)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    for t in range(8):
        b.step(acts[t], rew[t], done[t])
    evs, sink = [], []
    for t in range(40):
        sink.append(flush.max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step(acts[t % 8], rew[t % 8], done[t % 8]); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(z) for a, z in evs)
    out["cold_l2"] = entry(ENV_ID, n, ts[len(ts) // 2] * 1e3,
                           "ONE isolated k_step launch after an L2 flush, its own event pair (~5 us of the figure is the pair), median of 40",
                           "k_step")
    b.close()
    del acts, rew, done, flush, sink

    # BASELINE configs 1-2: N = 65536 (6.6 / 9.8 MB per step: latency-, not bandwidth-bound)
    for key, env_id in (("n65536_quadrotor3d", ENV_ID), ("n65536_quadrotor3d_slungload", "quadrotor3d-slungload-v0")):
        n = 1 << 16
        b, acts, rew, done = make(env_id, n, 64)
        e = entry(env_id, n, _time_many(torch, b, acts, rew, done, n, 4096, 0), "step stream, 64-step launches", "k_step_stream")
        e["launch_per_step_us"] = _time_many(torch, b, acts, rew, done, n, 2048, 2)
        e["rollout_in_registers_us"] = _time_many(torch, b, acts, rew, done, n, 4096, 1)
        out[key] = e
        b.close()
        del acts, rew, done
    return out


def measure_extras(torch, rmv, args, local):
    """Side measurements that explain the headline (not part of the contract line's value)."""
    dev = "cuda:%d" % local
    out = {}

    def time_steps(batch, acts, rew, done, n, k, fused=False, layout=0, act_ld=0):
        ring = acts.shape[0]

        def run(kk):
            left = kk
            while left > 0:
                m = min(left, ring)
                batch.step_many(m, acts, rew, done, out_ld=n, layout=layout, act_ld=act_ld, fused=fused)
                left -= m
        run(ring)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(k); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / k)
        return best

    peak, _ = load_peak()
    cases = [("fused_rollout", ENV_ID, 1 << 20, 64, 256, True, "tiles"),
             ("quadrotor3d-slungload-v0", "quadrotor3d-slungload-v0", 1 << 20, 32, 300, False, "tiles"),
             ("quadrotor2d-v0", "quadrotor2d-v0", 1 << 20, 32, 300, False, "tiles"),
             ("quadrotor2d-slungload-v0", "quadrotor2d-slungload-v0", 1 << 20, 32, 300, False, "tiles"),
             # a row-state handle (the state buffer is the observation matrix; the VecEnv default for the 3D kinds)
             ("row_state_stream_quadrotor3d_n65536", ENV_ID, 1 << 16, 64, 2048, False, "rows"),
             ("row_state_stream_quadrotor3d_n262144", ENV_ID, 1 << 18, 64, 1024, False, "rows")]
    bytes_of = {"quadrotor2d-v0": 53, "quadrotor2d-slungload-v0": 85, "quadrotor3d-v0": 101, "quadrotor3d-slungload-v0": 149}
    bounds = {"quadrotor2d-v0": (-10, 10), "quadrotor2d-slungload-v0": (-10, 10), "quadrotor3d-v0": (0, 10),
              "quadrotor3d-slungload-v0": (-10, 10)}
    for name, env_id, n, ring, k, fused, state_layout in cases:
        b = rmv.QuadBatch(env_id, n, device=local, seed=SEED, state_layout=state_layout)
        b.set_tuning(vec=args.vec, block=args.block)
        b.reset()
        A = b.action_dim
        lo, hi = bounds[env_id]
        acts = torch.empty((ring, n, A), dtype=torch.float32, device=dev)
        for t in range(ring):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, SEED, step_index=t, device=local)
        rew = torch.empty((ring, n), dtype=torch.float32, device=dev)
        done = torch.empty((ring, n), dtype=torch.uint8, device=dev)
        ms = time_steps(b, acts, rew, done, n, k, fused=fused)
        sps = n / (ms * 1e-3)
        out[name] = {"env_id": env_id, "n_envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": sps,
                     "algorithmic_GBps": sps * bytes_of[env_id] / 1e9, "frac_of_hbm_peak": sps * bytes_of[env_id] / 1e9 / peak,
                     "mode": "one launch per %d steps, state in registers" % ring if fused else "step stream (one launch per %d steps, state through memory every step)" % ring}
        if state_layout == "rows":
            out[name]["mode"] += "; RMV_STATE_ROWS handle: rows through the warp's shared tile both ways"
            out[name]["kernel"] = b.last_kernel
        b.close()
        del acts, rew, done

    # opt-in L2 access-policy window on the state (rmv_set_l2_persist): the band where the state fits the carve-out but
    # state + streams of one step do not fit L2
    for name, env_id, n in (("l2_persist_quadrotor3d_slungload", "quadrotor3d-slungload-v0", 1 << 20),
                            ("l2_persist_quadrotor3d_n21", ENV_ID, 1 << 21)):
        b = rmv.QuadBatch(env_id, n, device=local, seed=SEED)
        try:
            b.set_l2_persist(1)
        except Exception as exc:
            out[name] = {"refused": str(exc)}
            b.close()
            continue
        b.reset()
        lo, hi = bounds[env_id]
        acts = torch.empty((16, n, 4), dtype=torch.float32, device=dev)
        for t in range(16):
            rmv.fill_uniform_actions(acts[t], n, 4, lo, hi, SEED, step_index=t, device=local)
        rew = torch.empty((16, n), dtype=torch.float32, device=dev)
        done = torch.empty((16, n), dtype=torch.uint8, device=dev)
        ms = time_steps(b, acts, rew, done, n, 300)
        out[name] = {"env_id": env_id, "n_envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3),
                     "algorithmic_GBps": n * bytes_of[env_id] / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": n * bytes_of[env_id] / (ms * 1e-3) / 1e9 / peak,
                     "mode": "one launch per step, state pinned in L2 by an access-policy window (opt-in)"}
        b.set_l2_persist(0)
        b.close()
        del acts, rew, done

    # reinmav-v0 (closed-loop 13-state body, 50/51 float64 substeps per step, no action)
    n = 1 << 20
    b = rmv.QuadBatch("reinmav-v0", n, device=local)
    rew = torch.empty(n, dtype=torch.float32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        b.step(None, rew, done)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        b.step(None, rew, done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out["reinmav-v0"] = {"env_id": "reinmav-v0", "n_envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3),
                         "euler_substeps_per_s": 50 * n / (ms * 1e-3), "frac_of_hbm_peak": 109 * n / (ms * 1e-3) / 1e9 / peak,
                         "mode": "one launch per step; float64 substeps in registers (compute-bound, not HBM-bound)"}
    b.close()

    # BASELINE config 4: quadrotor2d-v0, N = 2^20, PPO2 rollout shape — obs [N,5] in HBM, actions [N,2] ~ N(0,1)
    # (baselines' initial diagonal Gaussian), one VecEnv.step per policy step, Monitor on (the VecEnv default).  Also for
    # quadrotor3d.  (a) QuadVecEnv.step_tensor: what a device-resident PPO loop calls; the 3D kinds keep their state as
    # the observation matrix (state_layout='rows'), so the step updates the obs tensor in place.  (b) RolloutBuffer:
    # step t reads obs[t] and writes obs[t+1] of a [T+1, N, S] buffer.
    import gym_reinmav_b200.rollout as rollout
    for env_id, key in (("quadrotor2d-v0", "ppo2_rollout_shape_quadrotor2d"), (ENV_ID, "ppo2_rollout_shape_quadrotor3d")):
        n, T = 1 << 20, 128
        Sd, Ad = {"quadrotor2d-v0": (5, 2), ENV_ID: (10, 4)}[env_id]
        venv = rmv.QuadVecEnv(env_id, n, device=local, seed=SEED, monitor=True)
        actions = torch.empty((T, n, Ad), dtype=torch.float32, device=dev).normal_()
        if env_id == ENV_ID:
            actions.mul_(2.5).add_(5.0)                  # around the middle of quadrotor3d's [0, 10) action box
        views = [actions[t] for t in range(T)]
        venv.reset_tensor()
        for t in range(T):
            venv.step_tensor(views[t])
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(T):
                venv.step_tensor(views[t])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / T)
        sps = n / (best * 1e-3)
        rows = venv.state_layout == "rows"
        nbytes = bytes_of[env_id] + (0 if rows else 4 * Sd) + 8          # + episode tracking (Monitor)
        out[key] = {"env_id": env_id, "n_envs": n, "n_steps": T, "us_per_step": best * 1e3, "env_steps_per_s": sps,
                    "bytes_per_env_step": nbytes, "algorithmic_GBps": sps * nbytes / 1e9, "frac_of_hbm_peak": sps * nbytes / 1e9 / peak,
                    "kernel": venv.batch.last_kernel,
                    "mode": "QuadVecEnv.step_tensor (one C-ABI call per step from Python), monitor=True, state_layout=%s" % venv.state_layout}
        stats = venv.episode_stats()
        out[key]["episodes_finished"] = stats["episodes"]
        if env_id == ENV_ID:
            # the same loop under VecNormalize (run.py:91-92): two more kernels per step (rmv_norm.cuh), running statistics
            # in HBM; bytes = the step's + 4(2S + 2) + 1 + 16 (obs in / out, reward in / out, done, the float64 return)
            norm = rmv.VecNormalize(venv)
            norm.reset_tensor()
            for t in range(16):
                norm.step_tensor(views[t])
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for t in range(T):
                    norm.step_tensor(views[t])
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / T)
            nb2 = nbytes + 4 * (2 * Sd + 2) + 1 + 16
            out["vecnormalize_quadrotor3d"] = {
                "env_id": env_id, "n_envs": n, "n_steps": T, "us_per_step": best * 1e3, "env_steps_per_s": n / (best * 1e-3),
                "bytes_per_env_step": nb2, "algorithmic_GBps": n / (best * 1e-3) * nb2 / 1e9,
                "frac_of_hbm_peak": n / (best * 1e-3) * nb2 / 1e9 / peak,
                "mode": "VecNormalize(QuadVecEnv).step_tensor: env step + k_norm_moments + k_norm_apply, statistics updated every step"}
            norm.close()                                  # closes venv too
            del norm
        venv.close()
        del venv
        venv = rmv.QuadVecEnv(env_id, n, device=local, seed=SEED, monitor=True)
        buf = rollout.RolloutBuffer(venv, T)
        buf.actions.copy_(actions)
        policy = lambda obs, _it=iter(range(10 ** 9)): buf.action_views[next(_it) % T]      # pre-drawn actions: env-side cost only
        buf.collect(policy)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); buf.collect(policy); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / T)
        sps = n / (best * 1e-3)
        nbytes = bytes_of[env_id] + (0 if buf.travelling else 4 * Sd) + 8
        out[key + "_rollout_buffer"] = {
            "env_id": env_id, "n_envs": n, "n_steps": T, "us_per_step": best * 1e3, "env_steps_per_s": sps,
            "bytes_per_env_step": nbytes, "algorithmic_GBps": sps * nbytes / 1e9, "frac_of_hbm_peak": sps * nbytes / 1e9 / peak,
            "mode": "RolloutBuffer.collect: [T+1,N,%d] obs buffer, %s, monitor=True"
                    % (Sd, "state travels through the buffer (rmv_step_rows: read obs[t], write obs[t+1])" if buf.travelling
                       else "tiled state in place + obs rows written into obs[t+1] (rmv_step)")}
        venv.close()
        del buf, venv, actions, views
    return out


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def load_traffic(key):
    """An entry of profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of committed ncu --set full
    captures); None if no capture is recorded under that key."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)[key]
    except Exception:
        return None


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under
    NCCL_DEBUG, warnings from C code), so everything else is sent to stderr and the JSON line is written
    to the original descriptor at the end."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS, help="envs per GPU")
    ap.add_argument("--vec", type=int, default=1)
    ap.add_argument("--block", type=int, default=0, help="threads per CTA of the step kernel (0 = the library's choice)")
    ap.add_argument("--repeats", type=int, default=5, help="minimum number of K-step blocks in the timed region")
    ap.add_argument("--min-region-ms", type=float, default=60.0, help="the K-step block is repeated until the timed region is this long")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-steps", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
