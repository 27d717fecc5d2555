"""Launch-shape sweep of the step kernel (developer tool; prints a table).  Needs a GPU."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reinmav-gym_b200")]

import torch  # noqa: E402

import gym_reinmav_b200 as rmv  # noqa: E402

BYTES = {"quadrotor2d-v0": 53, "quadrotor2d-slungload-v0": 85, "quadrotor3d-v0": 101, "quadrotor3d-slungload-v0": 149}
BOUNDS = {"quadrotor2d-v0": (-10, 10), "quadrotor2d-slungload-v0": (-10, 10), "quadrotor3d-v0": (0, 10),
          "quadrotor3d-slungload-v0": (-10, 10)}


def time_cfg(env_id, n, vec, block, layout, fused, steps, ring, peak, tma=0, ctas=0, obs=0, track=0, l2=0, rows=0,
             warps=0, sblock=0, apre=-1):
    dev = "cuda:0"
    b = rmv.QuadBatch(env_id, n, seed=0, track_episodes=bool(track), state_layout="rows" if rows == 3 else "tiles")
    row_state, rows = rows == 3, (0 if rows == 3 else rows)          # 3: a row-state handle through step_many
    b.set_tuning(vec=vec, block=block)
    b.set_kernel(tma=bool(tma), ctas_per_sm=ctas)
    b.set_stream(warps_per_sm=warps, block=sblock, action_prefetch=apre)
    if l2:
        try:
            b.set_l2_persist(1)
        except Exception as exc:        # state larger than the carve-out: refused by the library
            print("# l2 window refused: %s" % exc)
    b.reset()
    A = b.action_dim
    lo, hi = BOUNDS[env_id]
    if layout == "soa":
        acts = torch.empty((ring, A, n), dtype=torch.float32, device=dev)
        for t in range(ring):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t, layout=1, act_ld=n)
    else:
        acts = torch.empty((ring, n, A), dtype=torch.float32, device=dev)
        for t in range(ring):
            rmv.fill_uniform_actions(acts[t], n, A, lo, hi, 0, step_index=t)
    rew = torch.empty((ring, n), dtype=torch.float32, device=dev)
    done = torch.empty((ring, n), dtype=torch.uint8, device=dev)
    lay = 1 if layout == "soa" else 0

    if rows:
        # travelling state (rmv_step_rows): state read from rows[t], written to rows[t+1] of a ring; handle state untouched
        R = 4
        row_ring = torch.empty((R, n, b.state_dim), dtype=torch.float32, device=dev)
        b.reset(obs=row_ring[0])
        ap, rp, dp = [acts[t].data_ptr() for t in range(ring)], [rew[t].data_ptr() for t in range(ring)], [done[t].data_ptr() for t in range(ring)]
        op = [row_ring[t].data_ptr() for t in range(R)]
        stream = torch.cuda.current_stream().cuda_stream
        pos = [0]

        def run(k):
            t0 = pos[0]
            for t in range(t0, t0 + k):
                b.step_rows(op[t % R], ap[t % ring], rp[t % ring], dp[t % ring], op[(t + 1) % R] if rows == 1 else op[t % R],
                            layout=lay, act_ld=n, stream=stream)
            pos[0] = t0 + k
    elif obs:
        # obs rows [N][S] written by the kernel every step (the VecEnv / rollout-buffer variant); one C-ABI call per step
        obs_ring = torch.empty((2, n, b.state_dim), dtype=torch.float32, device=dev)
        ap, rp, dp = [acts[t].data_ptr() for t in range(ring)], [rew[t].data_ptr() for t in range(ring)], [done[t].data_ptr() for t in range(ring)]
        op = [obs_ring[t].data_ptr() for t in range(2)]
        stream = torch.cuda.current_stream().cuda_stream

        def run(k):
            for t in range(k):
                b.step(ap[t % ring], rp[t % ring], dp[t % ring], obs=op[t & 1], layout=lay, act_ld=n, stream=stream)
    else:
        def run(k):
            left = k
            while left > 0:
                m = min(left, ring)
                b.step_many(m, acts, rew, done, out_ld=n, layout=lay, act_ld=n, fused=fused)
                left -= m

    run(ring)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    b_state_dim = b.state_dim
    b.close()
    sps = n / (best * 1e-3)
    gbs = sps * (BYTES[env_id] + (4 * b_state_dim if (obs and not rows) else 0) + (8 if track else 0)) / 1e9
    return best * 1e3, sps, gbs, gbs / peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="quadrotor3d-v0")
    ap.add_argument("--n", default="1048576")
    ap.add_argument("--vec", default="1,2,4")
    ap.add_argument("--block", default="0", help="comma list of threads per CTA (0 = the library default for the kind and size)")
    ap.add_argument("--layout", default="soa,aos")
    ap.add_argument("--fused", default="0", help="comma list of rmv_step_many modes: 0 = step stream, 1 = rollout in registers, 2 = one launch per step")
    ap.add_argument("--warps", default="0", help="comma list: step-stream resident warps per SM cap (0 = all that fit)")
    ap.add_argument("--apre", default="-1", help="comma list: step-stream action prefetch (-1 auto, 0, 1)")
    ap.add_argument("--sblock", default="0", help="comma list: step-stream threads per CTA (0 auto, 32, 64, 128)")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--ring", type=int, default=16)
    ap.add_argument("--peak", type=float, default=6458.4)
    ap.add_argument("--tma", default="0", help="comma list of 0/1: register-path / TMA-pipelined kernel")
    ap.add_argument("--ctas", default="0", help="comma list: persistent CTAs per SM for the TMA kernel (0 = max)")
    ap.add_argument("--obs", type=int, default=0, help="1: the kernel also writes row-per-env observations [N][S] every step")
    ap.add_argument("--track", type=int, default=0, help="1: episode tracking (Monitor) on")
    ap.add_argument("--rows", type=int, default=0, help="1: rmv_step_rows through a ring of row buffers, 2: rmv_step_rows in place, 3: a RMV_STATE_ROWS handle through step_many")
    ap.add_argument("--l2", type=int, default=0, help="1: opt-in L2 access-policy window on the state (rmv_set_l2_persist)")
    args = ap.parse_args()
    print("%-26s %9s %3s %5s %4s %5s %9s %11s %9s %6s" % ("env", "n", "vec", "block", "lay", "fused", "us/step", "steps/s", "GB/s(alg)", "frac"))
    for env_id in args.envs.split(","):
        for n in (int(x) for x in args.n.split(",")):
            for fused in (int(x) for x in args.fused.split(",")):
                for layout in args.layout.split(","):
                    for vec in (int(x) for x in args.vec.split(",")):
                        for block in (int(x) for x in args.block.split(",")):
                            ring = max(2, min(args.ring, int(6e9 // (n * 4 * 6))))
                            for tma in (int(x) for x in args.tma.split(",")):
                                for ctas in (int(x) for x in args.ctas.split(",")) if tma else (0,):
                                    stream_cfgs = [(w_, sb_, ap_) for w_ in (int(x) for x in args.warps.split(","))
                                                   for sb_ in (int(x) for x in args.sblock.split(","))
                                                   for ap_ in (int(x) for x in args.apre.split(","))] if fused == 0 else [(0, 0, -1)]
                                    for warps, sblock, apre in stream_cfgs:
                                        us, sps, gbs, frac = time_cfg(env_id, n, vec, block, layout, fused, args.steps, ring, args.peak, tma, ctas, args.obs, args.track, args.l2, args.rows, warps, sblock, apre)
                                        print("%-26s %9d %3d %5d %4s %5d %9.2f %11.4g %9.1f %6.3f  tma=%d ctas=%d obs=%d track=%d l2=%d rows=%d warps=%d sblock=%d apre=%d" % (env_id, n, vec, block, layout, fused, us, sps, gbs, frac, tma, ctas, args.obs, args.track, args.l2, args.rows, warps, sblock, apre), flush=True)


if __name__ == "__main__":
    main()
