// rmv_kernels.cuh — the fused step kernel and its small companions (sm_100a).
//
// Data layout in HBM (DESIGN.md "Layout"):
//   state    float32 [S][ld]   component-major; env e of this shard is column e
//   episode  uint32  [ld]      bit 31 = ever-done flag, bits 0..30 = reset count (Philox index)
//   actions  float32 [n][A] (AOS) or [A][act_ld] (SOA), read with streaming loads
//   reward   float32 [n], done uint8 [n], written with streaming stores
//
// One thread owns V consecutive envs (V = 1, 2 or 4): every state row is then a V-wide vector
// load/store that a warp covers with 32*V*4 contiguous bytes.  The path is HBM-bound (~1 flop/B):
// no tensor cores, no shared-memory reuse; shared memory is used only to transpose the optional
// row-per-env observation copy so that its stores coalesce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rmv_math.cuh"

namespace rmv {

constexpr int kMaxBlock = 256;
constexpr uint32_t kEverDone = 0x80000000u;

struct StepArgs {
  float* state;            // [S][ld]
  uint32_t* episode;       // [ld]
  const float* act;        // see act_soa
  float* reward;           // [T][out_ld] or nullptr
  uint8_t* done;           // [T][out_ld] or nullptr
  float* obs;              // [n][S] or nullptr (only with n_steps == 1)
  // optional episode tracking (RMV_TRACK_EPISODES), all [ld]
  float* ep_return;        // running return of the current episode
  uint32_t* ep_start;      // step index at which the current episode began
  float* last_return;      // return / length of the most recently finished episode
  uint32_t* last_length;
  double* cum_return;      // sums over finished episodes
  unsigned long long* cum_length;
  uint32_t* fin_count;
  int64_t ld, n, act_ld, out_ld, act_step_stride, env_id_offset;
  uint64_t seed;
  uint32_t step_index;
  int n_steps, act_soa, auto_reset;
  DevParams P;
};

// ---- V-wide global access ----------------------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<1> { using T = float; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

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

// row-per-env actions: V*A contiguous floats starting at p
template <int N>
__device__ __forceinline__ void load_contig_stream(const float* p, float (&o)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(p) + i);
      o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
    }
  } else {
    static_assert(N == 2, "unexpected action vector length");
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    o[0] = t.x; o[1] = t.y;
  }
}

// ---- the fused step kernel -----------------------------------------------------------------
// dynamics + done + reward + reset-on-done (+ optional episode bookkeeping, + optional
// row-per-env observation copy) for n envs, n_steps consecutive steps with the state held in
// registers in between (n_steps == 1 for the ordinary step()).
template <int KIND, int V>
__global__ void __launch_bounds__(kMaxBlock) k_step(const StepArgs a) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int SP = (S % 2 == 0) ? S + 1 : S;   // odd row pitch: conflict-free transposes
  extern __shared__ float tile[];                // [warps][32][SP], only when a.obs != nullptr

  const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  const bool active = e0 < a.n;
  const bool full = e0 + V <= a.n;

  float s[V][S];
  float ep_ret[V];
  if (active) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float t[V];
      load_v<V, false>(a.state + c * a.ld + e0, t);   // rows are padded to ld: always in bounds
#pragma unroll
      for (int v = 0; v < V; ++v) s[v][c] = t[v];
    }
    if (a.ep_return) load_v<V, false>(a.ep_return + e0, ep_ret);
  }

  for (int step = 0; step < a.n_steps; ++step) {
    float rew[V];
    uint8_t dn[V];
    if (active) {
      float act[V][A];
      const float* ap = a.act + (int64_t)step * a.act_step_stride;
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

#pragma unroll
      for (int v = 0; v < V; ++v) {
        float n_pos, n_vel, n_rew;
        EnvStep<KIND>::run(a.P, s[v], act[v], n_pos, n_vel, n_rew);
        const bool done = is_done(a.P, n_pos, n_vel);
        float r = -n_rew;
        dn[v] = done ? 1 : 0;
        const int64_t e = e0 + v;
        if (done && e < a.n) {
          // rare lane (~1.3 % of env-steps under a random policy): flag word, terminal reward, re-draw
          const uint32_t word = a.episode[e];
          r = reward_of(a.P, true, (word & kEverDone) != 0u, n_rew);
          uint32_t count = word & ~kEverDone;
          if (a.auto_reset) {
            count += 1u;
            reset_draws<KIND>(a.seed, (uint64_t)(a.env_id_offset + e), count, s[v]);
          }
          a.episode[e] = count | kEverDone;
          if (a.ep_return) {
            const float total = ep_ret[v] + r;
            const uint32_t now = a.step_index + (uint32_t)step + 1u;
            const uint32_t len = now - a.ep_start[e];
            a.last_return[e] = total;
            a.last_length[e] = len;
            a.cum_return[e] += (double)total;
            a.cum_length[e] += len;
            a.fin_count[e] += 1u;
            a.ep_start[e] = now;
            ep_ret[v] = -r;          // becomes 0 after the "+= r" below
          }
        }
        rew[v] = r;
        if (a.ep_return) ep_ret[v] += r;
      }

      const int64_t o0 = (int64_t)step * a.out_ld + e0;
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
  }

  if (active) {
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float t[V];
#pragma unroll
      for (int v = 0; v < V; ++v) t[v] = s[v][c];
      store_v<V, false>(a.state + c * a.ld + e0, t);
    }
    if (a.ep_return) store_v<V, false>(a.ep_return + e0, ep_ret);
  }

  if (a.obs) {
    // row-per-env observation copy [n][S]: transpose through a per-warp shared tile so that
    // each store instruction writes runs of S contiguous floats instead of 32 scattered words.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* my = tile + warp * 32 * SP;
    const int64_t warp_e0 = e0 - (int64_t)lane * V;   // first env of this warp
#pragma unroll
    for (int v = 0; v < V; ++v) {
      __syncwarp();
#pragma unroll
      for (int c = 0; c < S; ++c) my[lane * SP + c] = s[v][c];
      __syncwarp();
#pragma unroll
      for (int i = lane; i < 32 * S; i += 32) {
        const int r = i / S, c = i - r * S;
        const int64_t e = warp_e0 + (int64_t)r * V + v;
        if (e < a.n) __stcs(a.obs + e * S + c, my[r * SP + c]);
      }
    }
  }
}

// ---- reset(): all envs or a masked subset --------------------------------------------------
template <int KIND>
__global__ void k_reset(float* state, uint32_t* episode, int64_t ld, int64_t n, const uint8_t* mask,
                        uint64_t seed, int64_t env_id_offset, float* obs, float* ep_return,
                        uint32_t* ep_start, uint32_t step_index) {
  constexpr int S = Dims<KIND>::S;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s[S];
  if (mask == nullptr || mask[e] != 0) {
    const uint32_t word = episode[e];
    const uint32_t count = (word & ~kEverDone) + 1u;
    episode[e] = (word & kEverDone) | count;     // reset() never clears steps_beyond_done (quadrotor3d.py:182-185)
    reset_draws<KIND>(seed, (uint64_t)(env_id_offset + e), count, s);
#pragma unroll
    for (int c = 0; c < S; ++c) state[c * ld + e] = s[c];
    if (ep_return) { ep_return[e] = 0.0f; ep_start[e] = step_index; }
  } else if (obs) {
#pragma unroll
    for (int c = 0; c < S; ++c) s[c] = state[c * ld + e];
  }
  if (obs) {
#pragma unroll
    for (int c = 0; c < S; ++c) obs[e * S + c] = s[c];
  }
}

// ---- [n][S] rows <-> [S][ld] columns (set_state / get_state; not on the hot path) ------------
__global__ void k_rows_to_cols(const float* rows, float* cols, int64_t n, int64_t ld, int S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int c = 0; c < S; ++c) cols[c * ld + e] = rows[e * S + c];
}
__global__ void k_cols_to_rows(const float* cols, float* rows, int64_t n, int64_t ld, int S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int c = 0; c < S; ++c) rows[e * S + c] = cols[c * ld + e];
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
