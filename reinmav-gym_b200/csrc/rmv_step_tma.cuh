// rmv_step_tma.cuh — the step kernel with TMA bulk copies and an mbarrier pipeline (sm_100a).
//
// Same arithmetic and same results as k_step (rmv_kernels.cuh); what changes is how bytes move.
// The tiled state layout makes the whole working set of 128 envs ONE contiguous block
// (S x 512 B of state, A x 512 B of row-per-env actions), so a persistent CTA can
//   * fetch tile j+2 and j+1 with `cp.async.bulk.shared.global` (TMA 1-D bulk copy, completion on an
//     mbarrier) while it computes tile j: loads in flight cost neither registers nor stalled warps, and
//     their depth is set by NSTAGE x tile bytes x resident CTAs instead of by warp occupancy —
//     which is what the register-path kernel lacks for the small-state kinds (quadrotor2d moves 28 B
//     per thread: too little memory-level parallelism at 64 warps/SM);
//   * write the tile back with `cp.async.bulk.global.shared` after the cooperative reset pass has
//     patched the finished envs IN SHARED MEMORY (no special-cased stores).
// Grid = SMs x resident CTAs; tiles are walked with stride gridDim.x.
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KIND, bool OBS, bool TRACK, int NSTAGE>
__global__ void __launch_bounds__(kBlock) k_step_tma(const StepArgs a) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int NB = (S + 3) / 4;
  constexpr int SP = (S % 2 == 0) ? S + 1 : S;
  constexpr uint32_t kStateBytes = S * kTile * 4, kActBytes = A * kTile * 4, kStageBytes = kStateBytes + kActBytes;
  extern __shared__ __align__(128) unsigned char smem_raw[];      // [NSTAGE][state | actions] (+ obs transpose tile)
  __shared__ __align__(8) uint64_t full_bar[NSTAGE];
  __shared__ uint32_t redraw_count[kWarps];
  __shared__ uint8_t redraw_lane[kWarps][32];

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t n_tiles = (n + kTile - 1) / kTile;
  const uint32_t G = gridDim.x;
  auto stage_state = [&](uint32_t s) { return reinterpret_cast<float*>(smem_raw + s * kStageBytes); };
  auto stage_act = [&](uint32_t s) { return reinterpret_cast<float*>(smem_raw + s * kStageBytes + kStateBytes); };
  float* const obs_tile = reinterpret_cast<float*>(smem_raw + NSTAGE * kStageBytes);

  // request tile `t` into stage `s` (one thread): state block always (tiles are padded), action block
  // only when the tile is complete (a ragged last tile reads its actions through the register path)
  auto request = [&](uint32_t t, uint32_t s) {
    const bool whole = (t + 1) * kTile <= n;
    mbar_expect_tx(&full_bar[s], kStateBytes + (whole ? kActBytes : 0u));
    bulk_load(stage_state(s), a.state + (size_t)t * (S * kTile), kStateBytes, &full_bar[s]);
    if (whole) bulk_load(stage_act(s), a.act + (size_t)t * (A * kTile), kActBytes, &full_bar[s]);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");     // previous grid's state writes are visible from here on
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      const uint32_t t = blockIdx.x + (uint32_t)s * G;
      if (t < n_tiles) request(t, (uint32_t)s);
    }
  }

  uint32_t k = 0;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += G, ++k) {
    const uint32_t stage = k % NSTAGE, parity = (k / NSTAGE) & 1u;
    float* const st = stage_state(stage);
    const float* const ac = stage_act(stage);
    const uint32_t tile_e0 = t * kTile, e = tile_e0 + tid;
    const bool valid = e < n;
    const bool whole = tile_e0 + kTile <= n;

    mbar_wait(&full_bar[stage], parity);
    float s[S], act[A], ep_ret = 0.0f;
#pragma unroll
    for (int c = 0; c < S; ++c) s[c] = st[c * kTile + tid];
    if (whole) {
      if constexpr (A == 4) {
        const float4 v = *reinterpret_cast<const float4*>(ac + tid * 4);
        act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(ac + tid * 2);
        act[0] = v.x; act[1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < A; ++c) act[c] = valid ? __ldcs(a.act + (size_t)e * A + c) : 0.0f;
    }
    if (TRACK && valid) ep_ret = a.ep_return[e];

    float pos_sq, vel_sq, n_rew;
    EnvStep<KIND>::run(a.P, s, act, pos_sq, vel_sq, n_rew);
    const bool done = is_done(a.P, pos_sq, vel_sq) && valid;
    float r = -n_rew;
    bool redraw = false;
    if (done) {
      // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
      const uint32_t word = a.episode[e];
      r = reward_of(a.P, true, (word & kEverDone) != 0u, n_rew);
      a.episode[e] = ((word & ~kEverDone) + (a.auto_reset ? 1u : 0u)) | kEverDone;
      redraw = a.auto_reset != 0;
      if (TRACK) {
        const float total = ep_ret + r;
        const uint32_t now = a.step_index + 1u;
        const uint32_t len = now - a.ep_start[e];
        a.last_return[e] = total;
        a.last_length[e] = len;
        a.cum_return[e] += (double)total;
        a.cum_length[e] += len;
        a.fin_count[e] += 1u;
        a.ep_start[e] = now;
        ep_ret = -r;
      }
    }
    if (valid) {
      __stcs(a.reward + e, scaled_reward(a.P, r));
      __stcs(reinterpret_cast<unsigned char*>(a.done) + e, (unsigned char)(done ? 1 : 0));
      if (TRACK) a.ep_return[e] = ep_ret + r;
    }
    // new state back into the stage (a finished env's slot is overwritten by the reset pass below)
#pragma unroll
    for (int c = 0; c < S; ++c) st[c * kTile + tid] = s[c];

    const uint32_t my_mask = __ballot_sync(0xffffffffu, redraw);
    if (lane == 0) redraw_count[warp] = __popc(my_mask);
    if (redraw) redraw_lane[warp][__popc(my_mask & ((1u << lane) - 1u))] = (uint8_t)lane;

    if (OBS) {
      float* my = obs_tile + warp * 32 * SP;
      const uint32_t warp_e0 = e - lane;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < S; ++c) my[lane * SP + c] = s[c];
      __syncwarp();
#pragma unroll
      for (int i = lane; i < 32 * S; i += 32) {
        const int rr = i / S, c = i - rr * S;
        const uint32_t eo = warp_e0 + (uint32_t)rr;
        if (eo < n && !((my_mask >> rr) & 1u)) __stcs(a.obs + (size_t)eo * S + c, my[rr * SP + c]);
      }
    }

    // refill the stage the PREVIOUS tile used, once its write-back has finished reading shared memory
    if (tid == 0 && k > 0) {
      bulk_wait_read_all();
      const uint32_t tn = t + (uint32_t)(NSTAGE - 1) * G;
      if (tn < n_tiles) request(tn, (k - 1) % NSTAGE);
    }

    fence_async_smem();                 // generic-proxy writes of this tile -> visible to the bulk store
    __syncthreads();
    uint32_t count[kWarps], total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { count[w] = redraw_count[w]; total += count[w]; }
    if (total != 0u) {                  // uniform
      for (uint32_t i = tid; i < total * NB; i += kBlock) {
        uint32_t slot = i / NB;
        const uint32_t b = i - slot * NB;
        uint32_t rel = 0;
        bool found = false;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          if (!found && slot < count[w]) { rel = (uint32_t)w * 32u + redraw_lane[w][slot]; found = true; }
          if (!found) slot -= count[w];
        }
        const uint32_t er = tile_e0 + rel;
        const uint32_t episode = a.episode[er] & ~kEverDone;     // written by the owner before the barrier
        const uint64_t id = (uint64_t)a.env_id_offset + er;
        uint32_t w4[4];
        philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), episode, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b,
                      (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w4);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const int c = 4 * (int)b + l;
          if (c < S) {
            const float x = uniform_from_bits(w4[l], -1.0f, 2.0f);
            st[c * kTile + rel] = x;
            if (OBS) a.obs[(size_t)er * S + c] = x;
          }
        }
      }
      fence_async_smem();
      __syncthreads();
    }
    if (tid == 0) bulk_store(a.state + (size_t)t * (S * kTile), st, kStateBytes);
  }
  if (tid == 0) bulk_wait_all();        // shared memory must outlive the last write-back
}

}  // namespace rmv
