#!/usr/bin/env python
"""bench.py — env-steps/s of the batched quadrotor3d-v0 step() on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N > 1: launched by torchrun)
    python bench.py --impl reference ...                          (the reference CPU arm)

A "step" is ONE pass of the hot path (one fused step kernel: dynamics + reward + done +
reset-on-done) over one batch of N_ENVS = 2^20 quadrotor3d-v0 envs per GPU with random-policy
actions (BASELINE.json metric / configs[1] scaled to the metric's N = 2^20).

* value     whole-job env-steps/s, inputs resident in HBM when the timed region starts
            (state in the handle, actions in a 64-step ring written by the Philox fill kernel).
* e2e       the same metric through the C-ABI call with HOST buffers (rmv_step_host: pinned
            host actions -> H2D -> kernel -> D2H of obs/reward/done every step).
* roofline  algorithmic bytes (101 B per env-step) / CUDA-event time per launch,
            against MEASURED_PEAKS.json hbm_gbs.
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
        sm, smax, reasons, busy = [], [], set(), []
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
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        under = [c for c, b in zip(sm, busy) if b] or sm
        if under:
            under.sort()
            out.update(sm_mhz=under[len(under) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(under))
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

    # ---- value: K steps, device-resident inputs
    run_steps(args.warmup)
    times = []
    for _ in range(args.repeats):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = batch.launches
        e0.record()
        run_steps(args.steps)
        e1.record()
        barrier()
        times.append(max_over_ranks(e0.elapsed_time(e1)))
        launches = batch.launches - launches0
    ms_total = min(times)            # best of --repeats runs of exactly K steps (each already the max over ranks)
    ms_per_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)

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

    extras = {}
    if args.extras and world == 1:
        try:                                   # side measurements must never cost the contract line
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
                       "launch": "one fused step kernel per step (vec=%d, block=%s), %d-step action ring; rmv_step_many issues the launches of a "
                                 "ring pass from C and replays them from a CUDA graph once it has seen the same call twice"
                                 % (args.vec, args.block or "auto", ring),
                       "l2": "no flush: per-step traffic %.0f MB (state %.0f MB read+written in place, re-read next step; "
                             "actions stream from a %.2f GiB ring > 126 MB L2); extras.n23 is the same kernel with a "
                             "working set that cannot be L2-resident" % (ALG_BYTES * n / 1e6, 4 * S * n / 1e6, ring * n * A * 4 / 2 ** 30),
                       "timing": "CUDA events on the launch stream, best of %d runs of exactly K steps, max over ranks" % args.repeats},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * A * 4, "d2h_bytes_per_step": n * (4 * S + 4 + 1),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "path": "rmv_step_host (C-ABI, pinned host buffers): H2D actions, step kernel with obs rows, D2H obs+reward+done, stream sync"},
            "gpu_launches": launches * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks, "unit": "GB/s", "frac": achieved / peaks,
                         "traffic": (load_traffic() or {}).get("dram_bytes_per_launch"),
                         "traffic_steady_state": ((load_traffic() or {}).get("steady_state") or {}).get("dram_bytes_per_launch"),
                         "traffic_note": "ncu --set full per launch: cold-cache capture (L2 flushed per replay) / "
                                         "--cache-control none capture of a back-to-back run (42 MB state L2-resident)",
                         "kernel": "rmv::k_step<QUAD3D, vec=%d, block=%s>" % (args.vec, args.block or "auto(128)"),
                         "algorithmic_bytes_per_launch": ALG_BYTES * n, "avg_launch_us": ms_per_step * 1e3, "peak_source": peak_src,
                         "note": "frac > 1 at N=2^20 because the 42 MB state stays L2-resident between steps "
                                 "(traffic_steady_state << algorithmic bytes); the HBM-bound point of the same kernel is "
                                 "extras.n23 (847 MB per step)"},
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
    cases = [("n23", ENV_ID, 1 << 23, 8, 100, False), ("fused_rollout", ENV_ID, 1 << 20, 64, 256, True),
             ("quadrotor3d-slungload-v0", "quadrotor3d-slungload-v0", 1 << 20, 32, 300, False),
             ("quadrotor2d-v0", "quadrotor2d-v0", 1 << 20, 32, 300, False),
             ("quadrotor2d-slungload-v0", "quadrotor2d-slungload-v0", 1 << 20, 32, 300, False)]
    bytes_of = {"quadrotor2d-v0": 53, "quadrotor2d-slungload-v0": 85, "quadrotor3d-v0": 101, "quadrotor3d-slungload-v0": 149}
    bounds = {"quadrotor2d-v0": (-10, 10), "quadrotor2d-slungload-v0": (-10, 10), "quadrotor3d-v0": (0, 10),
              "quadrotor3d-slungload-v0": (-10, 10)}
    for name, env_id, n, ring, k, fused in cases:
        b = rmv.QuadBatch(env_id, n, device=local, seed=SEED)
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
                     "mode": "one launch per %d steps, state in registers" % ring if fused else "one launch per step"}
        b.close()
        del acts, rew, done

    # The headline kernel with a COLD L2: 256 MiB are written between launches (> 126 MB L2), each launch is timed
    # by its own pair of events.  Every state/action byte then comes from HBM; the isolated launch also pays its
    # own ramp and launch latency (no overlap with a predecessor).
    n = 1 << 20
    b = rmv.QuadBatch(ENV_ID, n, device=local, seed=SEED)
    b.set_tuning(vec=args.vec, block=args.block)
    b.reset()
    acts = torch.empty((8, n, ACTION_DIM), dtype=torch.float32, device=dev)
    for t in range(8):
        rmv.fill_uniform_actions(acts[t], n, ACTION_DIM, ACT_LO, ACT_HI, SEED, step_index=t, device=local)
    rew = torch.empty((8, n), dtype=torch.float32, device=dev)
    done = torch.empty((8, n), dtype=torch.uint8, device=dev)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)      # 256 MiB, READ between launches: L2 ends up clean
    for t in range(8):
        b.step(acts[t], rew[t], done[t])
    evs = []
    sink = []
    for t in range(40):
        sink.append(flush.max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step(acts[t % 8], rew[t % 8], done[t % 8]); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(z) for a, z in evs)
    ms = ts[len(ts) // 2]
    out["n20_l2_flushed"] = {"env_id": ENV_ID, "n_envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3),
                             "algorithmic_GBps": n * 101 / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": n * 101 / (ms * 1e-3) / 1e9 / peak,
                             "mode": "one launch per step, 256 MiB read between launches (L2 flushed, left clean), one event pair "
                                     "per launch (the pair itself costs ~5 us: an N = 2^17 launch measures 8 us this way), "
                                     "median of 40; all state/action bytes come from HBM, no overlap with a predecessor"}
    b.close()
    del acts, rew, done, flush, sink

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

    # BASELINE configs[1], [2]: N = 65536 (launch-latency-bound: one launch moves 6.6 / 9.8 MB)
    for name, env_id in (("n65536_quadrotor3d", ENV_ID), ("n65536_quadrotor3d_slungload", "quadrotor3d-slungload-v0")):
        n = 1 << 16
        b = rmv.QuadBatch(env_id, n, device=local, seed=SEED)
        b.reset()
        lo, hi = bounds[env_id]
        acts = torch.empty((64, n, 4), dtype=torch.float32, device=dev)
        for t in range(64):
            rmv.fill_uniform_actions(acts[t], n, 4, lo, hi, SEED, step_index=t, device=local)
        rew = torch.empty((64, n), dtype=torch.float32, device=dev)
        done = torch.empty((64, n), dtype=torch.uint8, device=dev)
        ms = time_steps(b, acts, rew, done, n, 2000)
        out[name] = {"env_id": env_id, "n_envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3),
                     "algorithmic_GBps": n * bytes_of[env_id] / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": n * bytes_of[env_id] / (ms * 1e-3) / 1e9 / peak, "mode": "one launch per step (launch-bound)"}
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

    # BASELINE config 4: quadrotor2d-v0, N = 2^20, PPO2 rollout shape — obs rows [N,5] out, actions [N,2]
    # ~ N(0,1) in (baselines' initial diagonal Gaussian), T = 128 steps into [T, N, ...] device buffers.
    # "travelling": the rollout buffer carries the state (rmv_step_rows: read obs[t], write obs[t+1]);
    # "in_place": tiled state in the handle + a row copy per step (rmv_step with obs rows).  Also for quadrotor3d.
    import gym_reinmav_b200.rollout as rollout
    for env_id, key in (("quadrotor2d-v0", "ppo2_rollout_shape_quadrotor2d"), (ENV_ID, "ppo2_rollout_shape_quadrotor3d")):
        n, T = 1 << 20, 128
        Sd = {"quadrotor2d-v0": 5, ENV_ID: 10}[env_id]
        for travelling in (False, True):
            venv = rmv.QuadVecEnv(env_id, n, device=local, seed=SEED, monitor=False)
            buf = rollout.RolloutBuffer(venv, T, travelling=travelling)
            buf.actions.normal_()
            if env_id == ENV_ID:
                buf.actions.mul_(2.5).add_(5.0)          # around the middle of quadrotor3d's [0, 10) action box
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
            nbytes = bytes_of[env_id] + (0 if travelling else 4 * Sd)
            out[key + ("_travelling_state" if travelling else "")] = {
                "env_id": env_id, "n_envs": n, "n_steps": T, "us_per_step": best * 1e3, "env_steps_per_s": sps,
                "bytes_per_env_step": nbytes, "algorithmic_GBps": sps * nbytes / 1e9,
                "frac_of_hbm_peak": sps * nbytes / 1e9 / peak,
                "mode": ("one launch per step from Python (ctypes); state read from obs[t], written to obs[t+1] of a [T+1,N,%d] buffer "
                         "(rmv_step_rows)" if travelling else
                         "one launch per step from Python (ctypes); tiled state in place + obs rows written into obs[t+1] of a "
                         "[T+1,N,%d] buffer (rmv_step)") % Sd}
            venv.close()
            del buf
    return out


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the committed
    ncu --set full capture (profiles/); None if no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)["k_step_quadrotor3d_n2pow20"]
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
    ap.add_argument("--repeats", type=int, default=5)
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
