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
//   * no CTA barrier per tile either: a warp appends its finished envs to its OWN list in shared memory (ballot + prefix
//     popcount, no atomics); once per STEP the CTA meets at a barrier and re-draws everything its warps collected, one
//     (env, Philox block) pair per thread — the Philox code then runs in one warp per CTA and step (k = 4 units per warp:
//     512 env-steps, 6.6 finished envs, 20 busy lanes) instead of in every warp that saw a finish;
//   * the grid is sized so that every warp owns the same number of units (k = ceil(units / resident warps), warps =
//     ceil(units / k) rounded to whole CTAs of 4) and all of them are co-resident: at N = 2^20 that is 8192 warps x 4
//     units = 2048 CTAs on 2368 slots.  Unit u of warp w is w + j * warps: the warp's pointers advance by CONSTANT
//     strides (one 64-bit add per unit for the state, 32-bit env index for everything else).
//
// Same arithmetic (EnvStep), same Philox re-draws, same results bit for bit as T launches of k_step
// (tests/test_gpu_parity.py::test_step_stream_bitwise).
#pragma once
#include "rmv_kernels.cuh"
#include "rmv_pair.cuh"

namespace rmv {

struct StreamShape {
  uint32_t units_per_warp;   // k
  uint32_t warps;            // warps that own work, a multiple of 4; warp w owns units w, w + warps, ...
  uint32_t limit;            // envs covered by existing units: n rounded up to 32
  uint32_t env_stride;       // 32 * warps: envs between two units of one warp
};

constexpr int kStreamWarps = 4;       // warps per CTA at most (the launch may use fewer: a CTA is only a scheduling unit)
constexpr int kStreamList = 64;       // per-warp re-draw list: flushed when more than 32 entries are pending

constexpr int stream_min_ctas(int kind, bool track, bool apre = false) {
  // CTAs of 128 threads per SM: 16 = 2048 threads (<= 32 registers), 12 = 1536 (<= 40), 10 = 1280 (<= 48), 8 = 1024 (<= 64).
  // The action prefetch keeps A more values live: one step down (it is used where occupancy is not the constraint).
  const int base = (kind == QUAD3D_SLUNG) ? 10 : (kind == QUAD2D_SLUNG ? 12 : (track ? 12 : 16));
  if (!apre) return base;
  return base == 16 ? 12 : (base == 12 ? 10 : 8);
}

// re-draw `pairs / NB` finished envs: pair i = (entry i / NB, Philox block i % NB), pairs idx0, idx0 + stride, ...
// `item_of(slot)` returns (env index inside the launch, new reset count) of list entry `slot`.
template <int KIND, bool ROWS = false, typename ItemOf>
__device__ __forceinline__ void stream_redraw(const StepArgs& a, uint32_t pairs, uint32_t idx0, uint32_t stride, ItemOf item_of) {
  constexpr int S = Dims<KIND>::S;
  constexpr int NB = (S + 3) / 4;
  for (uint32_t i = idx0; i < pairs; i += stride) {
    const uint32_t slot = i / NB, b = i - slot * NB;
    const uint2 item = item_of(slot);
    const uint32_t er = item.x;
    const uint64_t id = (uint64_t)a.env_id_offset + er;
    uint32_t w4[4];
    philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), item.y, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b, a.ks, w4);
    if constexpr (ROWS) {                              // row-state handle: a.state is the [n_pad][S] observation matrix
      float* const row = a.state + (size_t)er * S + 4 * b;
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (S % 4 == 0 || 4 * b + l < S) row[l] = uniform_pm1(w4[l]);
    } else {
      float* const row = a.state + ((size_t)(er >> 7) * (S * kTile) + (er & (kTile - 1))) + 4 * b * kTile;
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (S % 4 == 0 || 4 * b + l < S) row[l * kTile] = uniform_pm1(w4[l]);
    }
  }
}

// APRE (row-per-env actions only): the action of the warp's NEXT work item — its next unit of this step, or its first
// unit of the next step — is fetched while the current unit is computed.  Actions stream from HBM (nothing else in a
// unit does at L2-resident sizes), so at small batches, where a warp has one unit and little company on its SM, that
// load is the longest wait of every step; the state cannot run ahead the same way (with one unit per warp the next item
// reads what this one writes).  Costs A registers: used where occupancy is not the constraint (k <= 2).
//
// ROWS (RMV_STATE_ROWS handles): a.state is the [n_pad][S] observation matrix.  A unit's 32 rows are 32 * S contiguous floats:
// they come in with S coalesced requests through the warp's shared tile and leave the same way (the row-state kernel's
// transposes, rmv_kernels.cuh); a finished env's row is stored like the others and overwritten by the re-draw, which the
// CTA barrier (or, in a mass finish, the warp barrier) orders behind it.
template <int KIND, bool TRACK, bool PREF, bool CTRL = false, bool APRE = false, bool ROWS = false>
__global__ void __launch_bounds__(128, CTRL ? 8 : stream_min_ctas(KIND, TRACK, APRE))
k_step_stream(const StepArgs a, const StreamShape shape) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int NB = (S + 3) / 4;
  constexpr int SP = row_pitch<S>();
  __shared__ uint2 redraw_list[kStreamWarps][kStreamList];
  __shared__ uint32_t redraw_count[kStreamWarps];
  __shared__ __align__(16) float row_tile[ROWS ? kStreamWarps * 32 * SP : 4];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t gw = (blockIdx.x * blockDim.x + tid) >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t limit = shape.limit;                                   // envs covered by existing units
  const uint32_t W = shape.warps, K = shape.units_per_warp;
  const uint32_t T = (uint32_t)a.n_steps;
  if (tid < kStreamWarps) redraw_count[tid] = 0u;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous launch on the stream has completed and flushed
  __syncthreads();

  // Units this warp owns: gw, gw + W, ..., K of them (the last one does not exist for some warps; warps gw >= W own none).
  // Control flow is kept CTA-uniform — loops with kernel-parameter trip counts, no early exit, the per-unit work under one
  // predicate — so that everything that is uniform (step pointers, counters, memory descriptors) lives in the uniform
  // register file; a per-warp trip count or a break makes the compiler re-materialise it per access from vector registers.
  uint2* const list = redraw_list[warp & (kStreamWarps - 1)];
  const uint32_t e_first = (gw < W) ? (gw << 5) + lane : limit;          // warps without work start past the end
  const uint32_t de = shape.env_stride;                                  // envs between two units of this warp
  const size_t ds = ROWS ? (size_t)de * S : (size_t)(de >> 7) * (S * kTile);   // the same distance in the state buffer (W % 4 == 0)
  float* const sp_first = ROWS ? a.state + ((size_t)(e_first - lane) * S + lane)       // element `lane` of the unit's 32 * S floats
                               : a.state + ((size_t)(e_first >> 7) * (S * kTile) + (e_first & (kTile - 1)));
  float* const my_tile = row_tile + (ROWS ? (warp & (kStreamWarps - 1)) * 32 * SP : 0);

  // per-step base pointers: uniform, advanced by a constant per step (the host always passes reward / done buffers:
  // a caller's NULL is parked in scratch with out_ld = 0)
  const float* act_t = a.act;
  float* rew_t = a.reward;
  unsigned char* done_t = reinterpret_cast<unsigned char*>(a.done);
  uint32_t now = a.step_index;
  float act_next[A];
  if constexpr (APRE) {                              // the first item's action
    if (e_first < n) load_contig_stream<A>(act_t + (size_t)(e_first * (uint32_t)A), act_next);
    else {
#pragma unroll
      for (int c = 0; c < A; ++c) act_next[c] = 0.0f;
    }
  }
  for (uint32_t t = 0; t < T; ++t, act_t += a.act_step_stride, rew_t += a.out_ld, done_t += a.out_ld) {
    if (t + 1u == T) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next launch may become resident as we drain
    ++now;
    float* sp = sp_first;
    uint32_t e = e_first, pending = 0u;
    for (uint32_t j = 0; j < K; ++j, sp += ds, e += de) {
      if (e < limit) {                             // warp-uniform: this unit exists
        const bool active = e < n;
        float s[S], act[A];
        uint32_t word = 0u, ep_started = 0u;
        float ep_ret = 0.0f;
        if constexpr (ROWS) {                                           // the rows are padded to whole tiles too: in bounds
          rows_to_tile<S, false>(sp, my_tile, lane, 32u);
          __syncwarp();
          row_from_smem<S>(my_tile + lane * SP, s);
        } else {
#pragma unroll
          for (int c = 0; c < S; ++c) s[c] = sp[c * kTile];             // tiles (and the per-env arrays) are padded: in bounds
        }
        if (PREF) word = a.episode[e];
        if (TRACK) { ep_ret = a.ep_return[e]; ep_started = a.ep_start[e]; }
        if constexpr (CTRL) {                        // closed loop: the reference's controller on the state just loaded
          control_law<KIND>(a.C, s, act);
          if (a.act_out && active) store_action_row<A>(a.act_out + (size_t)t * a.act_step_stride + (size_t)(e * (uint32_t)A), act);
        } else if constexpr (APRE) {
#pragma unroll
          for (int c = 0; c < A; ++c) act[c] = act_next[c];
          // next item: this warp's next unit of this step if it exists, else its first unit of the next step
          const bool same_step = (j + 1u < K) && (e + de < limit);
          const uint32_t en = same_step ? e + de : e_first;
          const float* const an = same_step ? act_t : act_t + a.act_step_stride;
          if ((same_step || t + 1u < T) && en < n) load_contig_stream<A>(an + (size_t)(en * (uint32_t)A), act_next);
        } else if (a.act_soa) {
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
        const bool redraw = done && a.auto_reset != 0;
        float rew = -n_rew;
        uint32_t count = 0u;
        // two thirds of the units hold no finished env: one vote lets those warps BRANCH around the bookkeeping instead of
        // issuing it predicated-off (the compiler's choice for a plain `if (done)`)
        const uint32_t done_mask = __ballot_sync(0xffffffffu, done);
        if (done_mask != 0u) {                       // warp-uniform; the barrier keeps the compiler from predicating the block
          __syncwarp();
          if (done) {
          // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
          const uint32_t w = PREF ? word : a.episode[e];
          rew = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew);
          count = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
          a.episode[e] = count | kEverDone;
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
        }
        if (active) {
          __stcs(rew_t + e, scaled_reward(a.P, rew));
          __stcs(done_t + e, (unsigned char)(done ? 1 : 0));
          if (TRACK) a.ep_return[e] = ep_ret + rew;
          if constexpr (!ROWS) {
            if (!redraw) {
#pragma unroll
              for (int c = 0; c < S; ++c) sp[c * kTile] = s[c];
            }
          }
        }
        if constexpr (ROWS) {                        // all 32 rows of the unit, coalesced (plain stores: the next step reads them)
          row_to_smem<S>(my_tile + lane * SP, s);
          __syncwarp();
          tile_to_rows<S, false, false>(my_tile, sp, lane, 32u);
          __syncwarp();                              // the tile is free for the next unit; the rows are ordered before a warp-level re-draw
        }
        // finished envs go to the warp's list (prefix popcount of the vote: no atomics)
        const uint32_t mask = a.auto_reset ? done_mask : 0u;
        if (mask != 0u) {                            // warp-uniform
          if (redraw) list[pending + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = make_uint2(e, count);
          pending += (uint32_t)__popc(mask);
          if (pending > 32u) {                       // a mass finish: the warp empties its own list right away
            __syncwarp();
            stream_redraw<KIND, ROWS>(a, pending * NB, lane, 32u, [&](uint32_t slot) { return list[slot]; });
            __syncwarp();
            pending = 0u;
          }
        }
      }
    }
    // once per step the CTA re-draws what its warps collected: pair i goes to thread i, so the Philox code runs in one
    // warp per CTA and step (20 of its lanes busy for quadrotor3d at k = 4) instead of once per warp that saw a finish
    if (lane == 0u) redraw_count[warp & (kStreamWarps - 1)] = pending;
    __syncthreads();
    const uint32_t c0 = redraw_count[0], c1 = redraw_count[1], c2 = redraw_count[2], c3 = redraw_count[3];
    const uint32_t total = c0 + c1 + c2 + c3;
    if (total != 0u) {                               // CTA-uniform
      stream_redraw<KIND, ROWS>(a, total * NB, tid, blockDim.x, [&](uint32_t slot) {
        uint32_t w = 0u;
        if (slot >= c0) { slot -= c0; w = 1u; if (slot >= c1) { slot -= c1; w = 2u; if (slot >= c2) { slot -= c2; w = 3u; } } }
        return redraw_list[w][slot];
      });
      __syncthreads();                               // owners read these rows at the next step; the lists are free again
    }
  }
}

// ---- two envs per thread on the packed float32 pipe (quadrotor3d; rmv_pair.cuh) -----------------------------------
// Same stream, unit = 64 envs per warp: lane l owns envs 2l, 2l+1 of the unit — 64-bit row accesses, two 128-bit action
// rows, FFMA2 / FMUL2 / FADD2 arithmetic.  ~115 issued instructions per env-step instead of ~190 at the same number of
// envs in flight per SM (8 CTAs x 4 warps x 64 envs).  Row-per-env actions, no episode tracking (the one-env kernel
// serves those); reward / done buffers 8- / 2-byte aligned with an even out_ld (checked by the host).
constexpr int kPairList = 128;        // per-warp re-draw list of the pair kernel: flushed when more than 64 entries are pending

__global__ void __launch_bounds__(128, 8)
k_step_stream_pair(const StepArgs a, const StreamShape shape) {
  constexpr int KIND = QUAD3D;
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int NB = (S + 3) / 4;
  __shared__ uint2 redraw_list[kStreamWarps][kPairList];
  __shared__ uint32_t redraw_count[kStreamWarps];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t gw = (blockIdx.x * blockDim.x + tid) >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t limit = shape.limit;                                   // envs covered by existing 64-env units
  const uint32_t W = shape.warps, K = shape.units_per_warp;
  const uint32_t T = (uint32_t)a.n_steps;
  if (tid < kStreamWarps) redraw_count[tid] = 0u;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();

  uint2* const list = redraw_list[warp & (kStreamWarps - 1)];
  const uint32_t e_first = (gw < W) ? (gw << 6) + 2u * lane : limit;
  const uint32_t de = shape.env_stride;                                  // 64 * W (W % 4 == 0: whole tiles)
  const size_t ds = (size_t)(de >> 7) * (S * kTile);
  float* const sp_first = a.state + ((size_t)(e_first >> 7) * (S * kTile) + (e_first & (kTile - 1)));

  const float* act_t = a.act;
  float* rew_t = a.reward;
  unsigned char* done_t = reinterpret_cast<unsigned char*>(a.done);
  for (uint32_t t = 0; t < T; ++t, act_t += a.act_step_stride, rew_t += a.out_ld, done_t += a.out_ld) {
    if (t + 1u == T) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float* sp = sp_first;
    uint32_t e = e_first, pending = 0u;
    for (uint32_t j = 0; j < K; ++j, sp += ds, e += de) {
      if (e < limit) {                             // warp-uniform: this unit exists
        const bool act0 = e < n, act1 = e + 1u < n;
        p2 s[S], act[A];
#pragma unroll
        for (int c = 0; c < S; ++c) s[c] = p2(*reinterpret_cast<const float2*>(sp + c * kTile));   // padded tiles: in bounds
        {
          const float4* const ap = reinterpret_cast<const float4*>(act_t + (size_t)(e * (uint32_t)A));
          const float4 r0 = act0 ? __ldcs(ap) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          const float4 r1 = act1 ? __ldcs(ap + 1) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          act[0] = p2(r0.x, r1.x); act[1] = p2(r0.y, r1.y); act[2] = p2(r0.z, r1.z); act[3] = p2(r0.w, r1.w);
        }
        p2 pos_sq, vel_sq, n_rew;
        quad3d_pair_step(a.P, s, act, pos_sq, vel_sq, n_rew);
        const bool done0 = is_done(a.P, pos_sq.v.x, vel_sq.v.x) && act0;
        const bool done1 = is_done(a.P, pos_sq.v.y, vel_sq.v.y) && act1;
        const bool redraw0 = done0 && a.auto_reset != 0, redraw1 = done1 && a.auto_reset != 0;
        float rew0 = -n_rew.v.x, rew1 = -n_rew.v.y;
        uint32_t count0 = 0u, count1 = 0u;
        if (done0 || done1) {                      // rare lane: the ever-done flag decides the terminal reward
          if (done0) {
            const uint32_t w = a.episode[e];
            rew0 = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew.v.x);
            count0 = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
            a.episode[e] = count0 | kEverDone;
          }
          if (done1) {
            const uint32_t w = a.episode[e + 1u];
            rew1 = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew.v.y);
            count1 = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
            a.episode[e + 1u] = count1 | kEverDone;
          }
        }
        if (act1) {                                // both envs exist
          __stcs(reinterpret_cast<float2*>(rew_t + e), make_float2(scaled_reward(a.P, rew0), scaled_reward(a.P, rew1)));
          __stcs(reinterpret_cast<uchar2*>(done_t + e), make_uchar2(done0 ? 1 : 0, done1 ? 1 : 0));
        } else if (act0) {
          __stcs(rew_t + e, scaled_reward(a.P, rew0));
          __stcs(done_t + e, (unsigned char)(done0 ? 1 : 0));
        }
        if (!(redraw0 || redraw1)) {
          if (act0) {                              // (a padding lane writes nothing)
#pragma unroll
            for (int c = 0; c < S; ++c) *reinterpret_cast<float2*>(sp + c * kTile) = s[c].v;
          }
        } else {
          if (!redraw0 && act0) {
#pragma unroll
            for (int c = 0; c < S; ++c) sp[c * kTile] = s[c].v.x;
          }
          if (!redraw1 && act1) {
#pragma unroll
            for (int c = 0; c < S; ++c) sp[c * kTile + 1] = s[c].v.y;
          }
        }
        const uint32_t m0 = __ballot_sync(0xffffffffu, redraw0), m1 = __ballot_sync(0xffffffffu, redraw1);
        if ((m0 | m1) != 0u) {                     // warp-uniform
          const uint32_t lt = (1u << lane) - 1u;
          if (redraw0) list[pending + (uint32_t)__popc(m0 & lt)] = make_uint2(e, count0);
          if (redraw1) list[pending + (uint32_t)__popc(m0) + (uint32_t)__popc(m1 & lt)] = make_uint2(e + 1u, count1);
          pending += (uint32_t)(__popc(m0) + __popc(m1));
          if (pending > 64u) {                     // a mass finish: the warp empties its own list right away
            __syncwarp();
            stream_redraw<KIND>(a, pending * NB, lane, 32u, [&](uint32_t slot) { return list[slot]; });
            __syncwarp();
            pending = 0u;
          }
        }
      }
    }
    if (lane == 0u) redraw_count[warp & (kStreamWarps - 1)] = pending;
    __syncthreads();
    const uint32_t c0 = redraw_count[0], c1 = redraw_count[1], c2 = redraw_count[2], c3 = redraw_count[3];
    const uint32_t total = c0 + c1 + c2 + c3;
    if (total != 0u) {                               // CTA-uniform
      stream_redraw<KIND>(a, total * NB, tid, blockDim.x, [&](uint32_t slot) {
        uint32_t w = 0u;
        if (slot >= c0) { slot -= c0; w = 1u; if (slot >= c1) { slot -= c1; w = 2u; if (slot >= c2) { slot -= c2; w = 3u; } } }
        return redraw_list[w][slot];
      });
      __syncthreads();
    }
  }
}

}  // namespace rmv
