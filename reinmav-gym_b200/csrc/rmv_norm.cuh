// rmv_norm.cuh — VecNormalize on the device: running observation / return normalisation of
// baselines' `VecNormalize` wrapper, the one the reference's launcher puts around a vec-env (gym_reinmav/run.py:91-92;
// there only for env_type == 'mujoco', so for the native envs it is an opt-in).
//
// Third-party algorithm (openai/baselines@90d6677, README.md:20,27 — not vendored, not in this image), restated from
// its published source baselines/common/vec_env/vec_normalize.py + baselines/common/running_mean_std.py:
//     ret  = ret * gamma + rew
//     ob_rms.update(obs)                 batch mean / var over the N envs merged into the running (mean, var, count)
//     obs' = clip((obs - ob_rms.mean) / sqrt(ob_rms.var + eps), -clipob, clipob)        with the UPDATED statistics
//     ret_rms.update(ret)
//     rew' = clip(rew / sqrt(ret_rms.var + eps), -cliprew, cliprew)                     (no mean subtraction)
//     ret[done] = 0
// with update_mean_var_count_from_moments (Chan et al. parallel variance):
//     delta = bm - m ; tot = c + bc ; m' = m + delta * bc / tot ; M2 = v*c + bv*bc + delta^2 * c * bc / tot ; v' = M2 / tot
// running statistics in float64 (as NumPy does there), initial mean 0, var 1, count 1e-4.
//
// Two kernels per step, no host round trip.  Both walk the observation matrix as ONE flat array of n * S floats, four at
// a time (float4, fully coalesced), with a grid stride that is a multiple of S: the component (flat index mod S) of each
// of a thread's four lanes is then the same in every iteration, so a thread keeps four running (sum, sum of squares)
// pairs in float64 registers / four (mean, 1/std) pairs, and never computes a modulus in the loop.
//   k_norm_moments  batch sums of every observation component (per-thread partials -> shared memory -> ONE atomicAdd per
//                   CTA and component) and, in a second loop over the envs, the return update and its sums
//   k_norm_apply    every CTA merges the batch moments into the running statistics (S + 1 components, in shared memory),
//                   then normalises its share; CTA 0 also stores the new statistics.  Statistics and batch sums are
//                   double-buffered by step parity, so no kernel ever reads a buffer another one is writing.
// The first version read one row per thread (stride-S addresses, 10 cache-line requests per load) and issued two FP64
// atomics per warp and component on 2 (S + 1) addresses: 222 us per step at N = 2^20, S = 10 against 15 us for the env
// step itself; this version 32 us (k_norm_moments 12.9 + k_norm_apply 20.1 under ncu, profiles/README.md r2c).
// One wave of CTAs is the fastest grid: 2 / 4 / 8 waves measured 49.6 / 59.9 / 82.4 us per VecNormalize step against 47.4
// (every CTA repeats the reduction epilogue and the statistics prologue).
// Algorithmic bytes: 4(2S + 2) + 1 + 16 per env-step (obs in/out, reward in/out, done, the float64 return).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmv {

constexpr int kNormMaxDim = 16;                      // largest observation (quadrotor3d-slungload)
constexpr int kNormStats = 2 * (kNormMaxDim + 1) + 2;   // mean[S+1], var[S+1], count_ob, count_ret   (component S = the return)
constexpr int kNormSums = 2 * (kNormMaxDim + 1);        // sum[S+1], sumsq[S+1]
// The batch sums are accumulated with atomics.  One copy of them would put every CTA's atomics on two cache lines of ONE
// L2 slice, where they serialise (measured: the slice busy for 11 000 cycles after the SMs had finished); CTA b adds into
// copy b mod kNormCopies, each copy on lines of its own, and k_norm_apply adds the copies up.
constexpr int kNormCopies = 32, kNormCopyStride = 64;   // doubles
constexpr int kNormSumsAll = kNormCopies * kNormCopyStride;

struct NormArgs {
  const float* obs;        // [n][S] raw observation rows (nullptr: observations are not normalised)
  const float* reward;     // [n] raw rewards (nullptr in reset())
  const uint8_t* done;     // [n]
  float* obs_out;          // [n][S]
  float* reward_out;       // [n]
  double* ret;             // [n] discounted return per env (float64, as baselines' self.ret)
  double* stats_cur;       // kNormStats: statistics before this step
  double* stats_next;      // kNormStats: statistics after this step (written by CTA 0 of k_norm_apply)
  double* sums_cur;        // kNormSumsAll: batch sums of this step, kNormCopies copies (filled by k_norm_moments)
  double* sums_next;       // kNormSumsAll: zeroed by CTA 0 of k_norm_apply for the next step
  int64_t n;
  int S;
  int norm_obs, norm_ret, update;
  float clip_obs, clip_rew;
  double gamma, epsilon;   // Python floats in baselines
};

constexpr int kNormBlock = 256;
constexpr int kNormGroups = 8;           // 2 * kNormMaxDim * kNormGroups <= kNormBlock

constexpr int kNormBatch = 4;            // chunks a thread loads before it touches any of them

// chunk i of the flat matrix (zeros past the end: they add nothing to the sums and are never stored)
template <int VEC>
__device__ __forceinline__ void norm_load(const float* __restrict__ base, int64_t i, int64_t chunks, float (&v)[VEC]) {
  if (VEC == 4) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < chunks) q = __ldg(reinterpret_cast<const float4*>(base) + i);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    v[0] = (i < chunks) ? __ldg(base + i) : 0.f;
  }
}

// the component of flat element j, for the four lanes of a thread's first chunk
template <int VEC>
__device__ __forceinline__ void norm_lane_components(int64_t chunk0, int S, int (&comp)[VEC]) {
#pragma unroll
  for (int k = 0; k < VEC; ++k) comp[k] = (int)((chunk0 * VEC + k) % S);
}

// sums over the batch; also advances the discounted returns.  gridDim.x is a multiple of S (norm_launch).
template <int VEC>
__global__ void __launch_bounds__(kNormBlock, 6) k_norm_moments(const NormArgs a) {
  const int S = a.S, tid = threadIdx.x;
  __shared__ double part[2][kNormBlock * VEC];
  __shared__ double part2[kNormBlock];
  double* const sums = a.sums_cur + (blockIdx.x % kNormCopies) * kNormCopyStride;
  // programmatic dependent launch, as the step kernels do: this grid may become resident while the env step drains, and
  // lets k_norm_apply do the same; nothing is read before the wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.norm_obs && a.obs) {
    const int64_t total = a.n * S, chunks = total / VEC, stride = (int64_t)gridDim.x * kNormBlock;
    const int64_t chunk0 = (int64_t)blockIdx.x * kNormBlock + tid;
    double sum[VEC], sq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { sum[k] = 0.0; sq[k] = 0.0; }
    for (int64_t i = chunk0; i < chunks; i += kNormBatch * stride) {        // kNormBatch loads in flight, then the arithmetic
      float v[kNormBatch][VEC];
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) norm_load<VEC>(a.obs, i + u * stride, chunks, v[u]);
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) { const double x = (double)v[u][k]; sum[k] += x; sq[k] = ::fma(x, x, sq[k]); }
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) { part[0][tid * VEC + k] = sum[k]; part[1][tid * VEC + k] = sq[k]; }
    __syncthreads();
    // thread (which, c, g) adds every kNormGroups-th partial of component c; thread (which, c, 0) adds the groups: one atomic
    const int which = tid / (S * kNormGroups), c = (tid / kNormGroups) % S, g = tid % kNormGroups;
    const bool reducer = tid < 2 * S * kNormGroups;
    double acc = 0.0;
    if (reducer) {
      const int first = (int)(((int64_t)blockIdx.x * kNormBlock * VEC) % S);      // component of the CTA's first lane
      for (int i = (c - first + S) % S + g * S; i < kNormBlock * VEC; i += S * kNormGroups) acc += part[which][i];
    }
    __syncthreads();
    if (reducer) part2[tid] = acc;
    __syncthreads();
    if (reducer && g == 0) {
#pragma unroll
      for (int j = 1; j < kNormGroups; ++j) acc += part2[tid + j];
      atomicAdd(sums + which * (kNormMaxDim + 1) + c, acc);
    }
    if (blockIdx.x == 0 && tid < (int)(total - chunks * VEC)) {                   // the last total mod VEC elements
      const int64_t j = chunks * VEC + tid;
      const double x = (double)a.obs[j];
      atomicAdd(sums + (int)(j % S), x);
      atomicAdd(sums + (kNormMaxDim + 1) + (int)(j % S), x * x);
    }
    __syncthreads();
  }
  if (a.reward) {
    double s = 0.0, q = 0.0;
    const int64_t estride = (int64_t)gridDim.x * kNormBlock;
    for (int64_t e = (int64_t)blockIdx.x * kNormBlock + tid; e < a.n; e += kNormBatch * estride) {
      double old[kNormBatch];
      float rew[kNormBatch];
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
        const int64_t eu = e + u * estride;
        old[u] = eu < a.n ? a.ret[eu] : 0.0;
        rew[u] = eu < a.n ? __ldg(a.reward + eu) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
        const int64_t eu = e + u * estride;
        if (eu < a.n) {
          const double r = old[u] * a.gamma + (double)rew[u];          // self.ret = self.ret * self.gamma + rews
          a.ret[eu] = r;
          s += r; q = ::fma(r, r, q);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((tid & 31) == 0) { part[0][tid >> 5] = s; part[1][tid >> 5] = q; }
    __syncthreads();
    if (tid < 2) {
      double acc = 0.0;
      for (int w = 0; w < kNormBlock / 32; ++w) acc += part[tid][w];
      atomicAdd(sums + tid * (kNormMaxDim + 1) + S, acc);
    }
  }
}

// RunningMeanStd.update_from_moments for one component
__device__ __forceinline__ void merge_moments(double mean, double var, double count, double bsum, double bsq, double bcount,
                                              double& mean_out, double& var_out, double& count_out) {
  const double bmean = bsum / bcount;
  double bvar = bsq / bcount - bmean * bmean;              // np.var (population variance) from the one-pass sums
  if (bvar < 0.0) bvar = 0.0;
  const double delta = bmean - mean, tot = count + bcount;
  mean_out = mean + delta * bcount / tot;
  const double m2 = var * count + bvar * bcount + delta * delta * count * bcount / tot;
  var_out = m2 / tot;
  count_out = tot;
}

template <int VEC>
__global__ void __launch_bounds__(kNormBlock, 5) k_norm_apply(const NormArgs a) {
  const int S = a.S;
  __shared__ double s_mean[kNormMaxDim + 1], s_inv[kNormMaxDim + 1];
  __shared__ double s_sums[4][kNormCopyStride];
  const int c = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");     // k_norm_moments has completed: the batch sums are final
  {                                                    // add up the copies of the batch sums: 4 groups of 8 copies, then 4 values
    const int slot = c & (kNormCopyStride - 1), grp = c >> 6;
    double acc = 0.0;
    if (slot < kNormSums) {
#pragma unroll
      for (int j = 0; j < kNormCopies / 4; ++j) acc += a.sums_cur[(grp * (kNormCopies / 4) + j) * kNormCopyStride + slot];
    }
    s_sums[grp][slot] = acc;
  }
  __syncthreads();
  if (c <= S) {                                       // component c < S: observation c; component S: the return
    const bool is_ret = (c == S);
    const bool on = is_ret ? (a.norm_ret && a.reward != nullptr) : (a.norm_obs && a.obs != nullptr);
    double mean = a.stats_cur[c], var = a.stats_cur[(kNormMaxDim + 1) + c];
    double count = a.stats_cur[2 * (kNormMaxDim + 1) + (is_ret ? 1 : 0)];
    const int q = (kNormMaxDim + 1) + c;
    const double bsum = (s_sums[0][c] + s_sums[1][c]) + (s_sums[2][c] + s_sums[3][c]);
    const double bsq = (s_sums[0][q] + s_sums[1][q]) + (s_sums[2][q] + s_sums[3][q]);
    if (on && a.update) merge_moments(mean, var, count, bsum, bsq, (double)a.n, mean, var, count);
    s_mean[c] = mean; s_inv[c] = 1.0 / ::sqrt(var + a.epsilon);
    if (blockIdx.x == 0) {                             // the statistics every later step starts from
      a.stats_next[c] = mean; a.stats_next[(kNormMaxDim + 1) + c] = var;
      if (c == 0) a.stats_next[2 * (kNormMaxDim + 1)] = count;          // component 0 carries the observation count
      if (is_ret) a.stats_next[2 * (kNormMaxDim + 1) + 1] = count;
    }
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < kNormSumsAll; i += kNormBlock) a.sums_next[i] = 0.0;
  }
  __syncthreads();
  if (a.obs && a.obs_out) {
    // (obs - mean) / sqrt(var + eps) evaluated in float64 as (obs - mean) * (1 / sqrt(var + eps)), stored as float32 and
    // clipped there: the bounds are float32 values and the cast is monotonic, so clip and cast commute
    const int64_t total = a.n * S, chunks = total / VEC, stride = (int64_t)gridDim.x * kNormBlock;
    const int64_t chunk0 = (int64_t)blockIdx.x * kNormBlock + threadIdx.x;
    int comp[VEC];
    norm_lane_components<VEC>(chunk0, S, comp);
    double m[VEC], iv[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = s_mean[comp[k]]; iv[k] = s_inv[comp[k]]; }
    // not normalising = mean 0, scale 1, no bound: the same arithmetic returns the input bits
    const float lim = a.norm_obs ? a.clip_obs : __int_as_float(0x7f800000);
    if (!a.norm_obs) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) { m[k] = 0.0; iv[k] = 1.0; }
    }
    for (int64_t i = chunk0; i < chunks; i += kNormBatch * stride) {
      float v[kNormBatch][VEC];
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) norm_load<VEC>(a.obs, i + u * stride, chunks, v[u]);
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
        const int64_t iu = i + u * stride;
        if (iu < chunks) {
          float y[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) y[k] = fminf(fmaxf((float)(((double)v[u][k] - m[k]) * iv[k]), -lim), lim);
          if (VEC == 4) reinterpret_cast<float4*>(a.obs_out)[iu] = make_float4(y[0], y[1], y[2], y[3]);
          else a.obs_out[iu] = y[0];
        }
      }
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < (int)(total - chunks * VEC)) {
      const int64_t j = chunks * VEC + threadIdx.x;
      const int cj = (int)(j % S);
      float y = a.obs[j];
      if (a.norm_obs) y = fminf(fmaxf((float)(((double)y - s_mean[cj]) * s_inv[cj]), -lim), lim);
      a.obs_out[j] = y;
    }
  }
  if (a.reward && a.reward_out) {
    const double iv = s_inv[S];
    const float lim = a.clip_rew;
    const int64_t estride = (int64_t)gridDim.x * kNormBlock;
    const bool clear = a.done != nullptr;
    for (int64_t e = (int64_t)blockIdx.x * kNormBlock + threadIdx.x; e < a.n; e += kNormBatch * estride) {
      float rew[kNormBatch];
      uint8_t fin[kNormBatch];
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
        const int64_t eu = e + u * estride;
        rew[u] = eu < a.n ? __ldg(a.reward + eu) : 0.f;
        fin[u] = (clear && eu < a.n) ? __ldg(a.done + eu) : (uint8_t)0;
      }
#pragma unroll
      for (int u = 0; u < kNormBatch; ++u) {
        const int64_t eu = e + u * estride;
        if (eu < a.n) {
          float r = rew[u];
          if (a.norm_ret) r = fminf(fmaxf((float)((double)r * iv), -lim), lim);   // no mean subtraction
          a.reward_out[eu] = r;
          if (fin[u]) a.ret[eu] = 0.0;                      // self.ret[news] = 0.
        }
      }
    }
  }
}

}  // namespace rmv
