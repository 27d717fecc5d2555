// rmv_stream.cuh — the persistent step STREAM: T consecutive steps of the fused step kernel without launch boundaries.
//
// k_step (rmv_kernels.cuh) is one launch per step.  Back to back, every launch re-synchronises the whole grid: step
// t+1 cannot touch memory before the LAST CTA of step t has drained (griddepcontrol.wait), although env e of step t+1
// depends on nothing but env e of step t.  At N = 65536 that boundary is two thirds of the step (3.0 us per launch for
// 6.6 MB of traffic).
//
// Here the dependency is honoured exactly and nothing else: a WARP owns a fixed set of 32-env units for the whole call
// and walks (step, unit) in order.  Its state still makes the round trip through memory every step — read from the
// tiled state buffer, written back in place; the next step reads what this one wrote (L2 at N = 2^20, HBM at N = 2^23:
// the same 4(2S+A+1)+1 bytes per env-step as k_step) — but no warp ever waits for another one:
//   * no grid-wide barrier between steps (one launch per T steps);
//   * no CTA barrier either: a warp appends its finished envs to its OWN list in shared memory (ballot + prefix
//     popcount, no atomics) and re-draws them itself once per step, one (env, Philox block) pair per lane — the Philox
//     code then runs once per warp and step (for k = 4 units: in 81 % of the warp-steps, with 5 of 32 lanes busy on
//     average) instead of once per unit that holds a finished env;
//   * the grid is sized so that every warp owns the same number of units (k = ceil(units / resident warps), warps =
//     ceil(units / k) rounded to whole CTAs of 4) and all of them are co-resident: at N = 2^20 that is 8192 warps x 4
//     units = 2048 CTAs on 2368 slots.  Unit u of warp w is w + j * warps: the warp's pointers advance by CONSTANT
//     strides (one 64-bit add per unit for the state, 32-bit env index for everything else).
//
// Same arithmetic (EnvStep), same Philox re-draws, same results bit for bit as T launches of k_step
// (tests/test_gpu_parity.py::test_step_stream_bitwise).
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

struct StreamShape {
  uint32_t units_per_warp;   // k (informational)
  uint32_t warps;            // warps that own work, a multiple of 4; warp w owns units w, w + warps, ...
};

constexpr int kStreamWarps = 4;       // warps per CTA at most (the launch may use fewer: a CTA is only a scheduling unit)
constexpr int kStreamList = 64;       // per-warp re-draw list: flushed when more than 32 entries are pending

constexpr int stream_min_ctas(int kind, bool track) {
  // CTAs of 128 threads per SM: 16 = 2048 threads (<= 32 registers), 12 = 1536 (<= 40), 10 = 1280 (<= 48)
  if (kind == QUAD3D_SLUNG) return 10;
  if (kind == QUAD2D_SLUNG) return 12;
  return track ? 12 : 16;
}

// re-draw the `pending` finished envs of this warp's list: pair i = (entry i / NB, Philox block i % NB), one per lane
template <int KIND>
__device__ __forceinline__ void stream_flush(const StepArgs& a, const uint2* list, uint32_t pending, uint32_t lane) {
  constexpr int S = Dims<KIND>::S;
  constexpr int NB = (S + 3) / 4;
  __syncwarp();                                    // the list entries written by other lanes are visible
  const uint32_t pairs = pending * NB;
  for (uint32_t i = lane; i < pairs; i += 32u) {
    const uint32_t slot = i / NB, b = i - slot * NB;
    const uint2 item = list[slot];                 // (env index inside the launch, new reset count)
    const uint32_t er = item.x;
    const uint64_t id = (uint64_t)a.env_id_offset + er;
    uint32_t w4[4];
    philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), item.y, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b, a.ks, w4);
    float* const row = a.state + ((size_t)(er >> 7) * (S * kTile) + (er & (kTile - 1))) + 4 * b * kTile;
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (S % 4 == 0 || 4 * b + l < S) row[l * kTile] = uniform_pm1(w4[l]);
  }
  __syncwarp();                                    // the owner lanes read these rows at the next step
}

template <int KIND, bool TRACK, bool PREF>
__global__ void __launch_bounds__(128, stream_min_ctas(KIND, TRACK))
k_step_stream(const StepArgs a, const StreamShape shape) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  __shared__ uint2 redraw_list[kStreamWarps][kStreamList];

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t units = (n + 31u) >> 5;
  const uint32_t W = shape.warps;
  const uint32_t T = (uint32_t)a.n_steps;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous launch on the stream has completed and flushed
  // units this warp owns: gw, gw + W, ..., K of them (the last one does not exist for some warps).  Both loops below have
  // UNIFORM trip counts (kernel parameters): with a per-warp trip count the compiler treats the loop as divergent and
  // re-materialises the memory descriptors of every load and store from vector registers (2 R2UR per access).
  const uint32_t K = shape.units_per_warp;
  if (gw >= W || gw >= units) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); return; }

  uint2* const list = redraw_list[(threadIdx.x >> 5) & (kStreamWarps - 1)];
  const uint32_t e_first = (gw << 5) + lane;
  const uint32_t de = W << 5;                                            // envs between two units of this warp
  const size_t ds = (size_t)(W >> 2) * (S * kTile);                      // the same distance in the tiled state (W % 4 == 0)
  float* const sp_first = a.state + ((size_t)(e_first >> 7) * (S * kTile) + (e_first & (kTile - 1)));

  for (uint32_t t = 0; t < T; ++t) {
    if (t + 1u == T) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next launch may become resident as we drain
    const float* const act_t = a.act + (size_t)t * a.act_step_stride;
    // uniform per step (never dereferenced when the base is NULL): candidates for the uniform register file
    float* const rew_t = a.reward + (size_t)t * a.out_ld;
    unsigned char* const done_t = reinterpret_cast<unsigned char*>(a.done) + (size_t)t * a.out_ld;
    const uint32_t now = a.step_index + t + 1u;
    float* sp = sp_first;
    uint32_t e = e_first, pending = 0u;
    for (uint32_t j = 0; j < K; ++j, sp += ds, e += de) {
      if (e >= (units << 5)) break;               // this warp's last unit does not exist (only ever at j = K - 1)
      const bool active = e < n;
      float s[S], act[A];
      uint32_t word = 0u, ep_started = 0u;
      float ep_ret = 0.0f;
#pragma unroll
      for (int c = 0; c < S; ++c) s[c] = sp[c * kTile];                 // tiles (and the per-env arrays) are padded: in bounds
      if (PREF) word = a.episode[e];
      if (TRACK) { ep_ret = a.ep_return[e]; ep_started = a.ep_start[e]; }
      if (a.act_soa) {
#pragma unroll
        for (int c = 0; c < A; ++c) act[c] = active ? __ldcs(act_t + (size_t)c * a.act_ld + e) : 0.0f;
      } else if (active) {
        load_contig_stream<A>(act_t + (size_t)(e * (uint32_t)A), act);  // n <= 2^30: e * A fits 32 bits
      } else {
#pragma unroll
        for (int c = 0; c < A; ++c) act[c] = 0.0f;
      }

      float pos_sq, vel_sq, n_rew;
      EnvStep<KIND>::run(a.P, s, act, pos_sq, vel_sq, n_rew);
      const bool done = is_done(a.P, pos_sq, vel_sq) && active;
      float rew = -n_rew;
      uint32_t count = 0u;
      bool redraw = false;
      if (done) {
        // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
        const uint32_t w = PREF ? word : a.episode[e];
        rew = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew);
        count = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
        a.episode[e] = count | kEverDone;
        redraw = a.auto_reset != 0;
        if (TRACK) {
          const float total = ep_ret + rew;
          a.last_return[e] = total;
          a.last_length[e] = now - ep_started;
          atomicAdd(a.cum_return + e, (double)total);
          atomicAdd(a.cum_length + e, (unsigned long long)(now - ep_started));
          atomicAdd(a.fin_count + e, 1u);
          a.ep_start[e] = now;
          ep_ret = -rew;                          // becomes 0 after the "+ rew" below
        }
      }
      if (active) {
        if (a.reward) __stcs(rew_t + e, scaled_reward(a.P, rew));
        if (a.done) __stcs(done_t + e, (unsigned char)(done ? 1 : 0));
        if (TRACK) a.ep_return[e] = ep_ret + rew;
        if (!redraw) {
#pragma unroll
          for (int c = 0; c < S; ++c) sp[c * kTile] = s[c];
        }
      }
      // finished envs go to the warp's list; they are re-drawn once per step (or when the list runs full)
      const uint32_t mask = __ballot_sync(0xffffffffu, redraw);
      if (mask != 0u) {                            // warp-uniform
        if (redraw) list[pending + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = make_uint2(e, count);
        pending += (uint32_t)__popc(mask);
        if (pending > 32u) { stream_flush<KIND>(a, list, pending, lane); pending = 0u; }
      }
    }
    if (pending != 0u) stream_flush<KIND>(a, list, pending, lane);
  }
}

}  // namespace rmv
