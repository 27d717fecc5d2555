// rmv_stream.cuh — the persistent step STREAM: T consecutive steps of the fused step kernel without launch boundaries.
//
// k_step (rmv_kernels.cuh) is one launch per step.  Back to back, every launch re-synchronises the whole grid: step
// t+1 cannot touch memory before the LAST CTA of step t has drained (griddepcontrol.wait), although env e of step t+1
// depends on nothing but env e of step t.  At N = 2^20 that boundary costs ~27 % (12.2 us per launch against 8.8 us for the
// same arithmetic without boundaries, profiles/r1b_*), at N = 65536 it is two thirds of the step (3.06 us against ~1 us
// of traffic).
//
// Here the dependency is honoured exactly and nothing else: a WARP owns a fixed set of 32-env units for the whole call
// and walks (step, unit) in order.  Its state still makes the round trip through memory every step — read from the
// tiled state buffer, written back in place; the next step reads what this one wrote (L2 at N = 2^20, HBM at N = 2^23:
// the same 4(2S+A+1)+1 bytes per env-step as k_step) — but no warp ever waits for another one:
//   * no grid-wide barrier between steps (one launch per T steps);
//   * no CTA barrier either: the finished envs of a unit are re-drawn by the lanes of the SAME warp (ballot, then one
//     (env, Philox block) pair per lane), so there is no shared memory, no atomic and no __syncthreads in the kernel;
//   * the grid is sized so that every warp owns the same number of units (k = ceil(units / resident warps), warps =
//     ceil(units / k)) and all of them are co-resident: at N = 2^20 that is 8192 warps x 4 units = 2048 CTAs on 2368 slots.
// PIPE: the loads of the warp's NEXT work item (its next unit of this step, or its first unit of the next step) are
// issued before the current one is computed — actions always, state whenever the next item is a different unit (k >= 2) —
// so a warp keeps two items' worth of loads in flight without doubling the resident warps.
//
// Same arithmetic (EnvStep), same Philox re-draws, same results bit for bit as T launches of k_step
// (tests/test_gpu_parity.py::test_step_stream_bitwise).
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

struct StreamShape {
  uint32_t units_per_warp;   // k
  uint32_t warps;            // warps that own work; warp w owns units w, w + warps, ...
};

template <int KIND, bool TRACK, bool PREF>
struct UnitRegs {
  float s[Dims<KIND>::S];
  float act[Dims<KIND>::A];
  uint32_t word;             // PREF: the env's episode word
  float ep_ret;              // TRACK
  uint32_t ep_started;       // TRACK
};

template <int KIND>
__device__ __forceinline__ void stream_load_actions(const StepArgs& a, const float* act_t, uint32_t e, bool active,
                                                    float (&act)[Dims<KIND>::A]) {
  constexpr int A = Dims<KIND>::A;
  if (a.act_soa) {
#pragma unroll
    for (int c = 0; c < A; ++c) act[c] = active ? __ldcs(act_t + (size_t)c * a.act_ld + e) : 0.0f;
  } else if (active) {
    load_contig_stream<A>(act_t + (size_t)e * A, act);
  } else {
#pragma unroll
    for (int c = 0; c < A; ++c) act[c] = 0.0f;
  }
}

template <int KIND, bool TRACK, bool PREF>
__device__ __forceinline__ void stream_load_state(const StepArgs& a, uint32_t e, UnitRegs<KIND, TRACK, PREF>& r) {
  constexpr int S = Dims<KIND>::S;
  const float* const sp = a.state + ((size_t)(e >> 7) * (S * kTile) + (e & (kTile - 1)));
#pragma unroll
  for (int c = 0; c < S; ++c) r.s[c] = sp[c * kTile];           // tiles (and the per-env arrays) are padded: in bounds
  if (PREF) r.word = a.episode[e];
  if (TRACK) { r.ep_ret = a.ep_return[e]; r.ep_started = a.ep_start[e]; }
}

// dynamics + done + reward + state write-back + in-warp re-draw of one 32-env unit at step t
template <int KIND, bool TRACK, bool PREF>
__device__ __forceinline__ void stream_finish_unit(const StepArgs& a, uint32_t t, uint32_t unit_e0, uint32_t lane,
                                                   bool active, UnitRegs<KIND, TRACK, PREF>& r) {
  constexpr int S = Dims<KIND>::S;
  constexpr int NB = (S + 3) / 4;
  const uint32_t e = unit_e0 + lane;
  float* const sp = a.state + ((size_t)(e >> 7) * (S * kTile) + (e & (kTile - 1)));
  float pos_sq, vel_sq, n_rew;
  EnvStep<KIND>::run(a.P, r.s, r.act, pos_sq, vel_sq, n_rew);
  const bool done = is_done(a.P, pos_sq, vel_sq) && active;
  float rew = -n_rew;
  uint32_t count = 0u;
  bool redraw = false;
  if (done) {
    // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
    const uint32_t w = PREF ? r.word : a.episode[e];
    rew = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew);
    count = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
    a.episode[e] = count | kEverDone;
    redraw = a.auto_reset != 0;
    if (TRACK) {
      const uint32_t now = a.step_index + t + 1u;
      const float total = r.ep_ret + rew;
      a.last_return[e] = total;
      a.last_length[e] = now - r.ep_started;
      atomicAdd(a.cum_return + e, (double)total);
      atomicAdd(a.cum_length + e, (unsigned long long)(now - r.ep_started));
      atomicAdd(a.fin_count + e, 1u);
      a.ep_start[e] = now;
      r.ep_ret = -rew;                            // becomes 0 after the "+= rew" below
    }
  }
  if (active) {
    const size_t o = (size_t)t * a.out_ld + e;
    if (a.reward) __stcs(a.reward + o, scaled_reward(a.P, rew));
    if (a.done) __stcs(reinterpret_cast<unsigned char*>(a.done) + o, (unsigned char)(done ? 1 : 0));
    if (TRACK) a.ep_return[e] = r.ep_ret + rew;
    if (!redraw) {
#pragma unroll
      for (int c = 0; c < S; ++c) sp[c * kTile] = r.s[c];
    }
  }
  // re-draw inside the warp: pair i = (i / NB)-th finished env of the unit, Philox block i % NB
  const uint32_t mask = __ballot_sync(0xffffffffu, redraw);
  if (mask != 0u) {                                // warp-uniform
    const uint32_t pairs = (uint32_t)__popc(mask) * NB;
    for (uint32_t i0 = 0; i0 < pairs; i0 += 32u) {
      const uint32_t i = i0 + lane;
      const bool mine = i < pairs;
      const uint32_t slot = i / NB, b = i - slot * NB;
      uint32_t m = mask;
      if (mine) for (uint32_t k = 0; k < slot; ++k) m &= m - 1u;        // strip `slot` set bits (slot is 0 almost always)
      const uint32_t src = mine ? (uint32_t)__ffs((int)m) - 1u : 0u;
      const uint32_t episode = __shfl_sync(0xffffffffu, count, src);
      if (mine) {
        const uint32_t er = unit_e0 + src;
        const uint64_t id = (uint64_t)a.env_id_offset + er;
        uint32_t w4[4];
        philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), episode, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b, a.ks, w4);
        float* const row = a.state + ((size_t)(er >> 7) * (S * kTile) + (er & (kTile - 1))) + 4 * b * kTile;
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (S % 4 == 0 || 4 * b + l < S) row[l * kTile] = uniform_pm1(w4[l]);
      }
    }
    __syncwarp();                                  // the owner lane reads these rows at the next step
  }
}

constexpr int stream_min_ctas(int kind, bool track, bool pipe) {
  // CTAs of 128 threads per SM: 16 = 2048 threads (<= 32 registers), 12 = 1536 (<= 40), 10 = 1280 (<= 48), 8 = 1024 (<= 64)
  if (pipe) return (kind == QUAD3D_SLUNG) ? 6 : (kind == QUAD2D ? 12 : 10);
  if (kind == QUAD3D_SLUNG) return 10;
  if (kind == QUAD2D_SLUNG) return 12;
  return track ? 12 : 16;
}

template <int KIND, bool TRACK, bool PREF, bool PIPE>
__global__ void __launch_bounds__(128, stream_min_ctas(KIND, TRACK, PIPE))
k_step_stream(const StepArgs a, const StreamShape shape) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t units = (n + 31u) >> 5;
  const uint32_t W = shape.warps, K = shape.units_per_warp;
  const uint32_t T = (uint32_t)a.n_steps;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous launch on the stream has completed and flushed
  if (gw >= W) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); return; }
  // units this warp owns: gw, gw + W, ... (< units); the last one may be missing for some warps
  const uint32_t mine = (units > gw) ? (units - gw + W - 1u) / W : 0u;
  if (mine == 0u) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); return; }
  (void)K;

  if constexpr (!PIPE) {
    for (uint32_t t = 0; t < T; ++t) {
      if (t + 1u == T) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // next launch may become resident as we drain
      const float* const act_t = a.act + (size_t)t * a.act_step_stride;
      for (uint32_t j = 0, u = gw; j < mine; ++j, u += W) {
        const uint32_t e0 = u << 5, e = e0 + lane;
        const bool active = e < n;
        UnitRegs<KIND, TRACK, PREF> r;
        stream_load_state<KIND, TRACK, PREF>(a, e, r);
        stream_load_actions<KIND>(a, act_t, e, active, r.act);
        stream_finish_unit<KIND, TRACK, PREF>(a, t, e0, lane, active, r);
      }
    }
  } else {
    // software pipeline over the flat item sequence (t, j): while item i is computed, item i+1's loads are in flight.
    // Item i+1's state may be fetched early only if it is a different unit than item i (its last writer is then item
    // i+1-mine, already retired by this warp); with one unit per warp only the actions run ahead.
    const bool ahead = mine >= 2u;
    UnitRegs<KIND, TRACK, PREF> cur;
    {
      const uint32_t e = (gw << 5) + lane;
      stream_load_state<KIND, TRACK, PREF>(a, e, cur);
      stream_load_actions<KIND>(a, a.act, e, e < n, cur.act);
    }
    uint32_t t = 0, j = 0, u = gw;
    const uint32_t items = T * mine;
    for (uint32_t i = 0; i < items; ++i) {
      // next item
      uint32_t tn = t, jn = j + 1u, un = u + W;
      if (jn == mine) { jn = 0u; un = gw; tn = t + 1u; }
      const bool more = i + 1u < items;
      const uint32_t en = (un << 5) + lane;
      UnitRegs<KIND, TRACK, PREF> nxt;
      if (more) {
        stream_load_actions<KIND>(a, a.act + (size_t)tn * a.act_step_stride, en, en < n, nxt.act);
        if (ahead) stream_load_state<KIND, TRACK, PREF>(a, en, nxt);
      } else {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      }
      const uint32_t e0 = u << 5;
      stream_finish_unit<KIND, TRACK, PREF>(a, t, e0, lane, e0 + lane < n, cur);
      if (more && !ahead) stream_load_state<KIND, TRACK, PREF>(a, en, nxt);
      if (more) cur = nxt;
      t = tn; j = jn; u = un;
    }
  }
}

}  // namespace rmv
