// rmv_kernels.cuh — the fused step kernel and its small companions (sm_100a).
//
// Data layout in HBM (DESIGN.md "Layout"):
//   state    float32 [n_tiles][S][128]  tiled struct-of-arrays: env e lives in tile e/128, lane e%128;
//                                       one tile = the S components of 128 consecutive envs = S*512
//                                       contiguous bytes.  A warp reads 32*V consecutive envs of one
//                                       component with one coalesced V-wide vector load, the rows of
//                                       a tile sit at COMPILE-TIME offsets from the thread's base
//                                       pointer (no per-row address arithmetic), and a CTA's whole
//                                       footprint is one contiguous block of DRAM pages.
//   episode  uint32  [n_pad]            bit 31 = ever-done flag, bits 0..30 = reset count (Philox index)
//   actions  float32 [n][A] (AOS, what a policy emits) or [A][act_ld] (SOA); streaming loads
//   reward   float32 [n], done uint8 [n]; streaming stores
//
// One thread owns V consecutive envs (V = 1, 2 or 4; V = 1 is the measured optimum).  ~1.5 flop/B: no
// tensor cores, no shared-memory reuse.  Shared memory is used for two things only: transposing
// the optional row-per-env observation copy so its stores coalesce, and the per-warp compacted lists
// that let the rare finished envs (~1.3 % of env-steps) be re-drawn by whichever threads are free
// instead of serialising a Philox evaluation inside every warp that contains one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rmv_math.cuh"

namespace rmv {

constexpr int kTile = 128;            // envs per state tile
constexpr uint32_t kEverDone = 0x80000000u;

struct StepArgs {
  float* state;            // tiled [n_tiles][S][128]
  uint32_t* episode;       // [n_pad]
  const float* act;        // see act_soa
  float* reward;           // [T][out_ld] or nullptr
  uint8_t* done;           // [T][out_ld] or nullptr
  float* obs;              // [n][S] or nullptr (single-step kernel only); k_step_rows: the rows the new state goes to
  const float* rows_in;    // k_step_rows only: [n][S] rows the state comes from (may alias obs)
  // optional episode tracking (RMV_TRACK_EPISODES), all [n_pad]
  float* ep_return;        // running return of the current episode
  uint32_t* ep_start;      // step index at which the current episode began
  float* last_return;      // return / length of the most recently finished episode
  uint32_t* last_length;
  double* cum_return;      // sums over finished episodes
  unsigned long long* cum_length;
  uint32_t* fin_count;
  int64_t n, act_ld, out_ld, act_step_stride, env_id_offset;
  uint64_t seed;
  uint32_t ks[20];         // Philox key schedule of `seed`: ks[2r] = lo + r*W0, ks[2r+1] = hi + r*W1
  uint32_t step_index;
  int n_steps, act_soa, auto_reset;
  DevParams P;
  // fused control() + step() (CTRL instantiations): the action is computed from the state by the reference's geometric
  // controller instead of being read; act_out (nullable, [n][A] rows) receives it for logging / tests
  CtrlParams C;
  float* act_out;
};

// ---- V-wide global access ----------------------------------------------------------------
template <int V, bool STREAM>
__device__ __forceinline__ void load_v(const float* p, float (&o)[V]) {
  if constexpr (V == 1) {
    o[0] = STREAM ? __ldcs(p) : *p;
  } else if constexpr (V == 2) {
    const float2 t = STREAM ? __ldcs(reinterpret_cast<const float2*>(p)) : *reinterpret_cast<const float2*>(p);
    o[0] = t.x; o[1] = t.y;
  } else {
    const float4 t = STREAM ? __ldcs(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  }
}
template <int V, bool STREAM>
__device__ __forceinline__ void store_v(float* p, const float (&o)[V]) {
  if constexpr (V == 1) {
    if (STREAM) __stcs(p, o[0]); else *p = o[0];
  } else if constexpr (V == 2) {
    const float2 t = make_float2(o[0], o[1]);
    if (STREAM) __stcs(reinterpret_cast<float2*>(p), t); else *reinterpret_cast<float2*>(p) = t;
  } else {
    const float4 t = make_float4(o[0], o[1], o[2], o[3]);
    if (STREAM) __stcs(reinterpret_cast<float4*>(p), t); else *reinterpret_cast<float4*>(p) = t;
  }
}
template <int V>
__device__ __forceinline__ void store_done(uint8_t* p, const uint8_t (&d)[V]) {
  if constexpr (V == 1) {
    __stcs(reinterpret_cast<unsigned char*>(p), (unsigned char)d[0]);
  } else if constexpr (V == 2) {
    __stcs(reinterpret_cast<uchar2*>(p), make_uchar2(d[0], d[1]));
  } else {
    __stcs(reinterpret_cast<uchar4*>(p), make_uchar4(d[0], d[1], d[2], d[3]));
  }
}

// row-per-env actions: N contiguous floats starting at p
template <int N>
__device__ __forceinline__ void load_contig_stream(const float* p, float (&o)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(p) + i);
      o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
    }
  } else {
    static_assert(N == 2 || N % 4 == 0, "unexpected action vector length");
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    o[0] = t.x; o[1] = t.y;
  }
}

template <int A>
__device__ __forceinline__ void store_action_row(float* p, const float (&a)[A]) {
  if constexpr (A == 4) *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]);
}

template <int KIND, int V>
__device__ __forceinline__ void load_actions(const StepArgs& a, const float* ap, int64_t e0, bool full,
                                             float (&act)[V][Dims<KIND>::A]) {
  constexpr int A = Dims<KIND>::A;
  if (a.act_soa) {
#pragma unroll
    for (int c = 0; c < A; ++c) {
      float t[V];
      if (full) {
        load_v<V, true>(ap + c * a.act_ld + e0, t);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) t[v] = (e0 + v < a.n) ? __ldcs(ap + c * a.act_ld + e0 + v) : 0.0f;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) act[v][c] = t[v];
    }
  } else {
    float t[V * A];
    if (full) {
      load_contig_stream<V * A>(ap + e0 * A, t);
    } else {
#pragma unroll
      for (int i = 0; i < V * A; ++i) t[i] = (e0 * A + i < a.n * A) ? __ldcs(ap + e0 * A + i) : 0.0f;
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int c = 0; c < A; ++c) act[v][c] = t[v * A + c];
  }
}

template <int V>
__device__ __forceinline__ void store_outputs(const StepArgs& a, int64_t o0, int64_t e0, bool full,
                                              const float (&rew)[V], const uint8_t (&dn)[V]) {
  if (full) {
    if (a.reward) store_v<V, true>(a.reward + o0, rew);
    if (a.done) store_done<V>(a.done + o0, dn);
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (e0 + v < a.n) {
        if (a.reward) __stcs(a.reward + o0 + v, rew[v]);
        if (a.done) __stcs(reinterpret_cast<unsigned char*>(a.done) + o0 + v, (unsigned char)dn[v]);
      }
    }
  }
}

// Book-keeping of an env that just returned done (quadrotor3d.py:112-122): terminal reward from the
// ever-done flag, flag set, reset count advanced when the env is going to be re-drawn.
// Returns the terminal reward; `count_out` = the new reset count (the Philox episode index).
__device__ __forceinline__ float on_done(const StepArgs& a, int64_t e, float n_rew, float ep_ret_before,
                                         uint32_t step_now, uint32_t& count_out) {
  const uint32_t word = a.episode[e];
  const float r = reward_of(a.P, true, (word & kEverDone) != 0u, n_rew);
  uint32_t count = word & ~kEverDone;
  if (a.auto_reset) count += 1u;
  a.episode[e] = count | kEverDone;
  count_out = count;
  if (a.ep_return) {
    const float total = ep_ret_before + r;
    const uint32_t len = step_now - a.ep_start[e];
    a.last_return[e] = total;
    a.last_length[e] = len;
    a.cum_return[e] += (double)total;
    a.cum_length[e] += len;
    a.fin_count[e] += 1u;
    a.ep_start[e] = step_now;
  }
  return r;
}

// ---- [32 rows][S] <-> a warp's shared tile, coalesced on the global side ------------------------------------
// A warp's 32 rows are 32*S contiguous floats in global memory; element i = lane + 32k is accessed by lane `lane`
// in pass k (one 128-byte request per pass).  In shared memory row r starts at r*P words, P = row_pitch<S>():
//   S = 5, 9   P = S      a thread's row access (stride S, odd) and the pass access (consecutive words) are conflict-free
//   S = 10     P = 10     the thread moves its row as five 64-bit words (stride 10: conflict-free per half-warp)
//   S = 16     P = 20     four 128-bit words (stride 20: conflict-free per quarter-warp); the pass index gains 4*(i/16)
// so that in every case the pass address is a COMPILE-TIME offset from one per-lane pointer on both sides — no
// division, no index arithmetic per element.  GUARD: only rows r < rows_here exist (last warp of a ragged batch).
template <int S> __host__ __device__ constexpr int row_pitch() { return S == 16 ? 20 : S; }

template <int S>
__device__ __forceinline__ void row_to_smem(float* row, const float (&s)[S]) {      // row = tile + lane * row_pitch<S>()
  if constexpr (S == 10) {
#pragma unroll
    for (int j = 0; j < 5; ++j) reinterpret_cast<float2*>(row)[j] = make_float2(s[2 * j], s[2 * j + 1]);
  } else if constexpr (S == 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(row)[j] = make_float4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < S; ++c) row[c] = s[c];
  }
}
template <int S>
__device__ __forceinline__ void row_from_smem(const float* row, float (&s)[S]) {
  if constexpr (S == 10) {
#pragma unroll
    for (int j = 0; j < 5; ++j) { const float2 t = reinterpret_cast<const float2*>(row)[j]; s[2 * j] = t.x; s[2 * j + 1] = t.y; }
  } else if constexpr (S == 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = reinterpret_cast<const float4*>(row)[j];
      s[4 * j] = t.x; s[4 * j + 1] = t.y; s[4 * j + 2] = t.z; s[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < S; ++c) s[c] = row[c];
  }
}
// shared-memory word of element i = lane + 32k: a per-lane base plus a compile-time offset
template <int S> __device__ __forceinline__ uint32_t pass_base(uint32_t lane) { return S == 16 ? lane + 4u * (lane >> 4) : lane; }
template <int S> __host__ __device__ constexpr int pass_step() { return S == 16 ? 40 : 32; }

template <int S, bool GUARD, bool STREAM>
__device__ __forceinline__ void tile_to_rows(const float* my, float* dst_lane, uint32_t lane, uint32_t rows_here) {
  const float* const src = my + pass_base<S>(lane);
#pragma unroll
  for (int k = 0; k < S; ++k) {
    if (!GUARD || (lane + 32u * k) < rows_here * S) {
      const float x = src[pass_step<S>() * k];
      if (STREAM) __stcs(dst_lane + 32 * k, x); else dst_lane[32 * k] = x;
    }
  }
}
template <int S, bool GUARD>
__device__ __forceinline__ void rows_to_tile(const float* src_lane, float* my, uint32_t lane, uint32_t rows_here) {
  float x[S];
#pragma unroll
  for (int k = 0; k < S; ++k)                        // all loads first: S independent 128-byte requests per warp in flight
    x[k] = (!GUARD || (lane + 32u * k) < rows_here * S) ? src_lane[32 * k] : 0.0f;
  float* const dst = my + pass_base<S>(lane);
#pragma unroll
  for (int k = 0; k < S; ++k) dst[pass_step<S>() * k] = x[k];
}

// Episode bookkeeping of a finished env without a dependent load: the caller fetched ep_return and ep_start together
// with the state; everything here is a plain store or a fire-and-forget reduction (one writer per env and step, so the
// sums are the same values a load-add-store would produce).
__device__ __forceinline__ void track_finished(const StepArgs& a, uint32_t e, float total, uint32_t started) {
  const uint32_t now = a.step_index + 1u;
  const uint32_t len = now - started;
  a.last_return[e] = total;
  a.last_length[e] = len;
  atomicAdd(a.cum_return + e, (double)total);
  atomicAdd(a.cum_length + e, (unsigned long long)len);
  atomicAdd(a.fin_count + e, 1u);
  a.ep_start[e] = now;
}

// ---- the fused single-step kernel ------------------------------------------------------------
// dynamics + done + reward + reset-on-done (+ optional episode bookkeeping, + optional
// row-per-env observation copy) for n envs: the reference's step() and the caller's reset-on-done
// in ONE launch.  BLOCK threads per CTA (128, 256 or 512), V envs per thread; indices are 32-bit
// (n < 2^31 per launch, enforced by the host), the only 64-bit arithmetic is the one base pointer
// per thread.
//
// Reset-on-done without serialising warps: an env that finished gets its terminal reward and its
// `episode` word updated by its owner thread, which then appends the env to a CTA-wide list in shared
// memory (one shared-memory atomic per finished env; warps without one pay nothing).  After the
// second __syncthreads of the kernel the CTA's finished envs are re-drawn cooperatively, ONE (env,
// Philox block) PAIR PER THREAD, densely packed into the first warp(s), straight into the state tile.
// The inline alternative makes every warp that contains a finished env (a third of all warps under a
// random policy) run the 3 x 10 Philox rounds of a quadrotor3d reset with one lane active; the larger
// the CTA, the fewer warps execute the Philox code at all (one warp per CTA while pairs <= 32).
constexpr int kBlock = 128;           // CTA size of the TMA variant and of V > 1
constexpr int kWarps = kBlock / 32;

constexpr int step_min_ctas(int kind, int v, bool obs, bool track, int block) {
  // resident-thread target per SM: 2048 (<= 32 registers) for the small-state kinds, 1536 for
  // quadrotor2d-slungload, 1280 when the obs copy / tracking / the 16-component state need registers
  if (v == 2) return (obs || track) ? 1 : 1024 / block;     // 64 registers: as many envs in flight per SM as V = 1 at 32
  if (v != 1) return 1;
  const int threads = (kind == QUAD3D_SLUNG) ? 1280 : (kind == QUAD2D_SLUNG ? 1536 : 2048);
  return threads / block > 0 ? threads / block : 1;
}

// One CTA's worth of envs up to (not including) the re-draw pass.  FULL: every env of the CTA exists
// (all CTAs but the last of a ragged batch), so no per-thread bounds checks.
template <int KIND, int V, bool OBS, bool TRACK, int BLOCK, bool PREF, bool FULL, bool CTRL>
__device__ __forceinline__ void step_envs(const StepArgs& a, uint32_t cta_e0, uint32_t* redraw_total,
                                          uint16_t* redraw_list, uint32_t* redraw_episode, float* tile) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int SP = (S % 2 == 0) ? S + 1 : S;     // odd row pitch: conflict-free transposes
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t e0 = cta_e0 + tid * V;
  const bool active = FULL || e0 < n;
  const bool full = FULL || e0 + V <= n;

  // thread's base pointer into the tiled state: rows of the tile are at constant offsets c*kTile
  float* const sp = a.state + ((size_t)(e0 >> 7) * (S * kTile) + (e0 & (kTile - 1)));
  float s[V][S];
  float act[V][A];
  float ep_ret[V];
  uint32_t ep_started[V];                          // TRACK: step index at which the running episode began
  uint32_t word[V];                                // PREF: the env's `episode` word, fetched with the state
  bool redraw[V];
#pragma unroll
  for (int v = 0; v < V; ++v) { redraw[v] = false; word[v] = 0u; }

  if (active) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float t[V];
      load_v<V, false>(sp + c * kTile, t);         // tiles are padded: always in bounds
#pragma unroll
      for (int v = 0; v < V; ++v) s[v][c] = t[v];
    }
    if (PREF) {
#pragma unroll
      for (int v = 0; v < V; ++v) word[v] = a.episode[e0 + v];   // padded to whole tiles: in bounds
    }
    if (TRACK) {
      load_v<V, false>(a.ep_return + e0, ep_ret);
#pragma unroll
      for (int v = 0; v < V; ++v) ep_started[v] = a.ep_start[e0 + v];      // padded to whole tiles: in bounds
    }
    if constexpr (!CTRL) load_actions<KIND, V>(a, a.act, (int64_t)e0, full, act);
  }
  __syncthreads();                                 // *redraw_total = 0 is visible; the loads are in flight
  if constexpr (CTRL) {
    static_assert(!CTRL || V == 1, "the fused controller is built for one env per thread");
    if (active) {
      control_law<KIND>(a.C, s[0], act[0]);        // quadrotor3d.py:126-180 and twins, from the state just loaded
      if (a.act_out) store_action_row<A>(a.act_out + (size_t)e0 * A, act[0]);
    }
  }

  if (active) {
    float rew[V];
    uint8_t dn[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float pos_sq, vel_sq, n_rew;
      EnvStep<KIND>::run(a.P, s[v], act[v], pos_sq, vel_sq, n_rew);
      const bool done = is_done(a.P, pos_sq, vel_sq);
      float r = -n_rew;
      dn[v] = done ? 1 : 0;
      const uint32_t e = e0 + v;
      if (done && (FULL || e < n)) {
        // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
        const uint32_t w = PREF ? word[v] : a.episode[e];
        r = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew);
        const uint32_t count = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
        a.episode[e] = count | kEverDone;
        if (a.auto_reset) {
          redraw[v] = true;
          const uint32_t slot = atomicAdd(redraw_total, 1u);
          redraw_list[slot] = (uint16_t)(tid * V + v);
          redraw_episode[slot] = count;            // the Philox episode index of the re-draw
        }
        if (TRACK) {
          track_finished(a, e, ep_ret[v] + r, ep_started[v]);
          ep_ret[v] = -r;                          // becomes 0 after the "+= r" below
        }
      }
      rew[v] = scaled_reward(a.P, r);
      if (TRACK) ep_ret[v] += r;
    }
    if (full) {
      store_v<V, true>(a.reward + e0, rew);
      store_done<V>(a.done + e0, dn);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (e0 + v < n) {
          __stcs(a.reward + e0 + v, rew[v]);
          __stcs(reinterpret_cast<unsigned char*>(a.done) + e0 + v, (unsigned char)dn[v]);
        }
    }
    // a re-drawn env's new state is written after the barrier; everything else is stored here
    bool any_redraw = false;
#pragma unroll
    for (int v = 0; v < V; ++v) any_redraw |= redraw[v];
    if (!any_redraw) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        float t[V];
#pragma unroll
        for (int v = 0; v < V; ++v) t[v] = s[v][c];
        store_v<V, false>(sp + c * kTile, t);
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (!redraw[v]) {
#pragma unroll
          for (int c = 0; c < S; ++c) sp[c * kTile + v] = s[v][c];
        }
    }
    if (TRACK) store_v<V, false>(a.ep_return + e0, ep_ret);
  }

  if (OBS) {
    // row-per-env observation copy [n][S]: transpose through a per-warp shared tile so that each store instruction
    // writes 128 contiguous bytes instead of 32 scattered words.  Rows of re-drawn envs are written here with their
    // terminal state and overwritten by the re-draw pass after the barrier (same CTA: ordered).
    const uint32_t warp_e0 = e0 - lane * V;
    if constexpr (V == 1) {
      float* my = tile + warp * 32 * row_pitch<S>();
      __syncwarp();
      row_to_smem<S>(my + lane * row_pitch<S>(), s[0]);
      __syncwarp();
      float* const dst_lane = a.obs + (size_t)warp_e0 * S + lane;
      if (FULL) {
        tile_to_rows<S, false, true>(my, dst_lane, lane, 32u);
      } else {
        const uint32_t rows_here = warp_e0 < n ? (n - warp_e0 < 32u ? n - warp_e0 : 32u) : 0u;
        tile_to_rows<S, true, true>(my, dst_lane, lane, rows_here);
      }
    } else {
      float* my = tile + warp * 32 * SP;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        __syncwarp();
#pragma unroll
        for (int c = 0; c < S; ++c) my[lane * SP + c] = s[v][c];
        __syncwarp();
#pragma unroll
        for (int i = lane; i < 32 * S; i += 32) {
          const int r = i / S, c = i - r * S;
          const uint32_t e = warp_e0 + (uint32_t)r * V + v;
          if (FULL || e < n) __stcs(a.obs + (size_t)e * S + c, my[r * SP + c]);
        }
      }
    }
  }
}

template <int KIND, int V, bool OBS, bool TRACK, int BLOCK, bool PREF, bool CTRL = false>
__global__ void __launch_bounds__(BLOCK, CTRL ? 8 : step_min_ctas(KIND, V, OBS, TRACK, BLOCK))
k_step(const StepArgs a) {
  constexpr int S = Dims<KIND>::S;
  constexpr int NB = (S + 3) / 4;                  // Philox blocks per reset
  static_assert(BLOCK * V <= 65536, "redraw_list holds 16-bit indices");
  __shared__ uint32_t redraw_total;                // finished envs of this CTA ...
  __shared__ uint16_t redraw_list[BLOCK * V];      // ... their indices inside the CTA, in arrival order ...
  __shared__ uint32_t redraw_episode[BLOCK * V];   // ... and their new reset counts
  extern __shared__ __align__(16) float tile[];    // [warps][32][pitch], only when OBS

  // Programmatic dependent launch: let the NEXT launch in the stream become resident while this grid
  // drains, and do not touch global memory before the PREVIOUS grid has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tid = threadIdx.x;
  if (tid == 0) redraw_total = 0u;
  const uint32_t cta_e0 = blockIdx.x * (uint32_t)(BLOCK * V);
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (cta_e0 + (uint32_t)(BLOCK * V) <= (uint32_t)a.n)     // uniform
    step_envs<KIND, V, OBS, TRACK, BLOCK, PREF, true, CTRL>(a, cta_e0, &redraw_total, redraw_list, redraw_episode, tile);
  else
    step_envs<KIND, V, OBS, TRACK, BLOCK, PREF, false, CTRL>(a, cta_e0, &redraw_total, redraw_list, redraw_episode, tile);

  // ---- re-draw the finished envs of this CTA: one (env, Philox block) pair per thread ------------
  __syncthreads();
  const uint32_t pairs = redraw_total * NB;
  for (uint32_t i = tid; i < pairs; i += BLOCK) {
    const uint32_t slot = i / NB, b = i - slot * NB;
    const uint32_t e = cta_e0 + redraw_list[slot];
    const uint32_t episode = redraw_episode[slot];
    const uint64_t id = (uint64_t)a.env_id_offset + e;
    uint32_t w4[4];
    philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), episode, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b,
                     a.ks, w4);
    float* const row = a.state + ((size_t)(e >> 7) * (S * kTile) + (e & (kTile - 1))) + 4 * b * kTile;
    float* const orow = OBS ? a.obs + (size_t)e * S + 4 * b : nullptr;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (S % 4 == 0 || 4 * b + l < S) {
        const float x = uniform_pm1(w4[l]);
        row[l * kTile] = x;
        if (OBS) orow[l] = x;
      }
    }
  }
}

// ---- single step on row-major state ("travelling state") ------------------------------------------
// The state comes from rows_in [n][S] and goes to obs [n][S]; the handle's tiled state is not touched.  This is
// the shape of an on-device rollout (gym_reinmav/run.py:63-68 -> ppo2 Runner): step t reads the observation rows
// obs[t] — which ARE the state (obs == state in the reference, quadrotor3d.py:104) — and writes obs[t+1], so a
// step moves 4(2S + A + 1) + 1 bytes per env instead of the 4(3S + A + 1) + 1 of "tiled state in place + a row
// copy", and the rows it reads are the ones the previous step just wrote (still in L2).  rows_in == obs is the
// in-place form.  Same arithmetic, same reset pass, same results as k_step.
//
// A warp owns 32 consecutive rows = 32*S contiguous floats: loaded with S coalesced 128-byte requests into a
// shared tile [32][S|1] (odd pitch), read back one row per thread; stored the same way in reverse.
template <int KIND, bool TRACK, int BLOCK, bool CTRL = false>
__global__ void __launch_bounds__(BLOCK, CTRL ? 1024 / BLOCK : step_min_ctas(KIND, 1, true, TRACK, BLOCK))
k_step_rows(const StepArgs a) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int NB = (S + 3) / 4;
  constexpr int SP = row_pitch<S>();
  __shared__ uint32_t redraw_total;
  __shared__ uint16_t redraw_list[BLOCK];
  __shared__ uint32_t redraw_episode[BLOCK];
  extern __shared__ __align__(16) float tile[];    // [warps][32][SP]

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) redraw_total = 0u;
  const uint32_t n = (uint32_t)a.n;
  const uint32_t cta_e0 = blockIdx.x * (uint32_t)BLOCK;
  const uint32_t e = cta_e0 + tid, warp_e0 = e - lane;
  const bool active = e < n;
  float* const my = tile + warp * 32 * SP;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // rows in: coalesced, transposed through the warp's tile
  const uint32_t rows_here = warp_e0 < n ? (n - warp_e0 < 32u ? n - warp_e0 : 32u) : 0u;
  if (rows_here == 32u) rows_to_tile<S, false>(a.rows_in + (size_t)warp_e0 * S + lane, my, lane, 32u);
  else rows_to_tile<S, true>(a.rows_in + (size_t)warp_e0 * S + lane, my, lane, rows_here);
  float act[1][A];
  float ep_ret = 0.0f;
  uint32_t ep_started = 0u;
  if (active) {
    if constexpr (!CTRL) load_actions<KIND, 1>(a, a.act, (int64_t)e, true, act);
    if (TRACK) { ep_ret = a.ep_return[e]; ep_started = a.ep_start[e]; }
  } else {
#pragma unroll
    for (int c = 0; c < A; ++c) act[0][c] = 0.0f;
  }
  __syncwarp();
  float s[S];
  row_from_smem<S>(my + lane * SP, s);
  __syncthreads();                                 // redraw_total = 0 is visible
  if constexpr (CTRL) {
    if (active) {
      control_law<KIND>(a.C, s, act[0]);
      if (a.act_out) store_action_row<A>(a.act_out + (size_t)e * A, act[0]);
    }
  }

  bool redraw = false;
  if (active) {
    float pos_sq, vel_sq, n_rew;
    EnvStep<KIND>::run(a.P, s, act[0], pos_sq, vel_sq, n_rew);
    const bool done = is_done(a.P, pos_sq, vel_sq);
    float r = -n_rew;
    if (done) {
      // rare lane: the ever-done flag decides the terminal reward (quadrotor3d.py:112-122)
      const uint32_t w = a.episode[e];
      r = reward_of(a.P, true, (w & kEverDone) != 0u, n_rew);
      const uint32_t count = (w & ~kEverDone) + (a.auto_reset ? 1u : 0u);
      a.episode[e] = count | kEverDone;
      if (a.auto_reset) {
        redraw = true;
        const uint32_t slot = atomicAdd(&redraw_total, 1u);
        redraw_list[slot] = (uint16_t)tid;
        redraw_episode[slot] = count;
      }
      if (TRACK) {
        track_finished(a, e, ep_ret + r, ep_started);
        ep_ret = -r;
      }
    }
    __stcs(a.reward + e, scaled_reward(a.P, r));
    __stcs(reinterpret_cast<unsigned char*>(a.done) + e, (unsigned char)(done ? 1 : 0));
    if (TRACK) a.ep_return[e] = ep_ret + r;
  }

  // rows out (plain stores: the next step reads them); a re-drawn env's row is overwritten by the re-draw pass
  __syncwarp();
  row_to_smem<S>(my + lane * SP, s);
  __syncwarp();
  if (rows_here == 32u) tile_to_rows<S, false, false>(my, a.obs + (size_t)warp_e0 * S + lane, lane, 32u);
  else tile_to_rows<S, true, false>(my, a.obs + (size_t)warp_e0 * S + lane, lane, rows_here);

  __syncthreads();
  const uint32_t pairs = redraw_total * NB;
  for (uint32_t i = tid; i < pairs; i += BLOCK) {
    const uint32_t slot = i / NB, b = i - slot * NB;
    const uint32_t er = cta_e0 + redraw_list[slot];
    const uint64_t id = (uint64_t)a.env_id_offset + er;
    uint32_t w4[4];
    philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), redraw_episode[slot], ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | b,
                     a.ks, w4);
    float* const orow = a.obs + (size_t)er * S + 4 * b;
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (S % 4 == 0 || 4 * b + l < S) orow[l] = uniform_pm1(w4[l]);
  }
}

// ---- fused multi-step rollout ---------------------------------------------------------------------
// n_steps consecutive steps with auto-reset; each thread keeps its envs' state in registers in
// between, so state is read and written once per call (actions in, reward/done out every step).
// ROWS (RMV_STATE_ROWS handles, V = 1): a.state is the [n][S] observation matrix; a thread reads and writes its own row
// (S consecutive floats) — twice per call, so the transposes of the per-step kernels are not worth their shared memory here.
template <int KIND, int V, bool ROWS = false>
__global__ void __launch_bounds__(128) k_rollout(const StepArgs a) {
  static_assert(!ROWS || V == 1, "the row-state rollout is one env per thread");
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (e0 >= a.n) return;
  const bool full = e0 + V <= a.n;
  float* const sp = ROWS ? a.state + e0 * S : a.state + ((e0 >> 7) * (int64_t)(S * kTile) + (e0 & (kTile - 1)));
  float s[V][S];
  float ep_ret[V];
  if constexpr (ROWS) {
#pragma unroll
    for (int c = 0; c < S; ++c) s[0][c] = sp[c];
  } else {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float t[V];
      load_v<V, false>(sp + c * kTile, t);
#pragma unroll
      for (int v = 0; v < V; ++v) s[v][c] = t[v];
    }
  }
  if (a.ep_return) load_v<V, false>(a.ep_return + e0, ep_ret);

  for (int step = 0; step < a.n_steps; ++step) {
    float rew[V];
    uint8_t dn[V];
    float act[V][A];
    load_actions<KIND, V>(a, a.act + (int64_t)step * a.act_step_stride, e0, full, act);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float pos_sq, vel_sq, n_rew;
      EnvStep<KIND>::run(a.P, s[v], act[v], pos_sq, vel_sq, n_rew);
      const bool done = is_done(a.P, pos_sq, vel_sq);
      float r = -n_rew;
      dn[v] = done ? 1 : 0;
      const int64_t e = e0 + v;
      if (done && e < a.n) {
        uint32_t count;
        r = on_done(a, e, n_rew, a.ep_return ? ep_ret[v] : 0.0f, a.step_index + (uint32_t)step + 1u, count);
        if (a.ep_return) ep_ret[v] = -r;
        if (a.auto_reset) {
          const uint64_t id = (uint64_t)(a.env_id_offset + e);
#pragma unroll
          for (int b = 0; b < (S + 3) / 4; ++b) {
            uint32_t w4[4];
            philox4x32_10_ks((uint32_t)id, (uint32_t)(id >> 32), count, ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | (uint32_t)b,
                             a.ks, w4);
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (4 * b + l < S) s[v][4 * b + l] = uniform_pm1(w4[l]);
          }
        }
      }
      rew[v] = scaled_reward(a.P, r);
      if (a.ep_return) ep_ret[v] += r;
    }
    store_outputs<V>(a, (int64_t)step * a.out_ld + e0, e0, full, rew, dn);
  }

  if constexpr (ROWS) {
#pragma unroll
    for (int c = 0; c < S; ++c) sp[c] = s[0][c];
  } else {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float t[V];
#pragma unroll
      for (int v = 0; v < V; ++v) t[v] = s[v][c];
      store_v<V, false>(sp + c * kTile, t);
    }
  }
  if (a.ep_return) store_v<V, false>(a.ep_return + e0, ep_ret);
}

// ---- reset(): all envs or a masked subset --------------------------------------------------
// ROWS: the handle keeps its state as [n][S] rows (RMV_STATE_ROWS) instead of tiles.
template <int KIND, bool ROWS>
__global__ void k_reset(float* state, uint32_t* episode, int64_t n, const uint8_t* mask, uint64_t seed,
                        int64_t env_id_offset, float* obs, float* ep_return, uint32_t* ep_start,
                        uint32_t step_index) {
  constexpr int S = Dims<KIND>::S;
  constexpr int kPitch = ROWS ? 1 : kTile;         // distance between two components of one env
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* const row = ROWS ? state + e * S : state + (e / kTile) * (int64_t)(S * kTile) + (e % kTile);
  float s[S];
  if (mask == nullptr || mask[e] != 0) {
    const uint32_t word = episode[e];
    const uint32_t count = (word & ~kEverDone) + 1u;
    episode[e] = (word & kEverDone) | count;     // reset() never clears steps_beyond_done (quadrotor3d.py:182-185)
    reset_draws<KIND>(seed, (uint64_t)(env_id_offset + e), count, s);
#pragma unroll
    for (int c = 0; c < S; ++c) row[c * kPitch] = s[c];
    if (ep_return) { ep_return[e] = 0.0f; ep_start[e] = step_index; }
  } else if (obs) {
#pragma unroll
    for (int c = 0; c < S; ++c) s[c] = row[c * kPitch];
  }
  if (obs && obs != state) {
#pragma unroll
    for (int c = 0; c < S; ++c) obs[e * S + c] = s[c];
  }
}

// ---- [n][S] rows <-> tiled state (set_state / get_state; not on the hot path) ------------------
__global__ void k_rows_to_tiles(const float* rows, float* tiles, int64_t n, int S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* const row = tiles + (e / kTile) * (int64_t)(S * kTile) + (e % kTile);
  for (int c = 0; c < S; ++c) row[c * kTile] = rows[e * S + c];
}
__global__ void k_tiles_to_rows(const float* tiles, float* rows, int64_t n, int S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* const row = tiles + (e / kTile) * (int64_t)(S * kTile) + (e % kTile);
  for (int c = 0; c < S; ++c) rows[e * S + c] = row[c * kTile];
}

// ---- random-policy actions ~ U[lo, hi) ---------------------------------------------------------
__global__ void k_fill_uniform(float* act, int64_t n, int A, int soa, int64_t act_ld, float lo,
                               float width, uint64_t seed, int64_t env_id_offset, uint32_t step_index) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint64_t id = (uint64_t)(env_id_offset + e);
  for (int b = 0; b < (A + 3) / 4; ++b) {
    uint32_t w[4];
    philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), step_index, (STREAM_ACTION << 8) | (uint32_t)b,
                  (uint32_t)seed, (uint32_t)(seed >> 32), w);
    for (int l = 0; l < 4 && 4 * b + l < A; ++l) {
      const int c = 4 * b + l;
      const float x = uniform_from_bits(w[l], lo, width);
      if (soa) act[c * act_ld + e] = x; else act[e * A + c] = x;
    }
  }
}

// ---- gather of per-env episode info for a list of env indices ------------------------------------
__global__ void k_gather_episodes(const float* last_return, const uint32_t* last_length, const int64_t* idx, int64_t k,
                                  float* out_return, uint32_t* out_length) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  out_return[i] = last_return[idx[i]];
  out_length[i] = last_length[idx[i]];
}

// ---- episode statistics: sums over envs ---------------------------------------------------------
__global__ void k_episode_stats(const uint32_t* fin_count, const double* cum_return,
                                const unsigned long long* cum_length, int64_t n, double* out3) {
  double c = 0.0, r = 0.0, l = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    c += (double)fin_count[e]; r += cum_return[e]; l += (double)cum_length[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
    l += __shfl_xor_sync(0xffffffffu, l, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out3, c); atomicAdd(out3 + 1, r); atomicAdd(out3 + 2, l); }
}

}  // namespace rmv
