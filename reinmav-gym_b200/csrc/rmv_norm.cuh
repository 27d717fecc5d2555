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
// Two kernels per step, no host round trip:
//   k_norm_moments  one pass over obs rows [n][S] and rewards: ret update, per-component sum and sum of squares in
//                   float64 (warp shuffles -> one atomicAdd per warp and component)
//   k_norm_apply    every CTA merges the batch moments into the running statistics (S + 1 components, in shared memory),
//                   then normalises its envs; CTA 0 also stores the new statistics.  Statistics and batch sums are
//                   double-buffered by step parity, so no kernel ever reads a buffer another one is writing.
// HBM-bound: 4(2S + 2) + 1 + 16 bytes per env-step (obs in/out, reward in/out, done, the float64 return).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmv {

constexpr int kNormMaxDim = 16;                      // largest observation (quadrotor3d-slungload)
constexpr int kNormStats = 2 * (kNormMaxDim + 1) + 2;   // mean[S+1], var[S+1], count_ob, count_ret   (component S = the return)
constexpr int kNormSums = 2 * (kNormMaxDim + 1);        // sum[S+1], sumsq[S+1]

struct NormArgs {
  const float* obs;        // [n][S] raw observation rows (nullptr: observations are not normalised)
  const float* reward;     // [n] raw rewards (nullptr in reset())
  const uint8_t* done;     // [n]
  float* obs_out;          // [n][S]
  float* reward_out;       // [n]
  double* ret;             // [n] discounted return per env (float64, as baselines' self.ret)
  double* stats_cur;       // kNormStats: statistics before this step
  double* stats_next;      // kNormStats: statistics after this step (written by CTA 0 of k_norm_apply)
  double* sums_cur;        // kNormSums: batch sums of this step (filled by k_norm_moments)
  double* sums_next;       // kNormSums: zeroed by CTA 0 of k_norm_apply for the next step
  int64_t n;
  int S;
  int norm_obs, norm_ret, update;
  float clip_obs, clip_rew;
  double gamma, epsilon;   // Python floats in baselines
};

// sums over the batch; also advances the discounted returns
__global__ void __launch_bounds__(256) k_norm_moments(const NormArgs a) {
  const int S = a.S;
  double sum[kNormMaxDim + 1], sq[kNormMaxDim + 1];
#pragma unroll
  for (int c = 0; c <= kNormMaxDim; ++c) { sum[c] = 0.0; sq[c] = 0.0; }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x) {
    if (a.norm_obs && a.obs) {
#pragma unroll
      for (int c = 0; c < kNormMaxDim; ++c)
        if (c < S) { const double x = (double)a.obs[e * S + c]; sum[c] += x; sq[c] += x * x; }
    }
    if (a.reward) {
      const double r = a.ret[e] * a.gamma + (double)a.reward[e];     // self.ret = self.ret * self.gamma + rews
      a.ret[e] = r;
      sum[kNormMaxDim] += r; sq[kNormMaxDim] += r * r;
    }
  }
#pragma unroll
  for (int c = 0; c <= kNormMaxDim; ++c) {
    if (c < S || c == kNormMaxDim) {
      double s = sum[c], q = sq[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if ((threadIdx.x & 31) == 0) {
        const int slot = (c == kNormMaxDim) ? S : c;
        atomicAdd(a.sums_cur + slot, s);
        atomicAdd(a.sums_cur + (kNormMaxDim + 1) + slot, q);
      }
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

__global__ void __launch_bounds__(256) k_norm_apply(const NormArgs a) {
  const int S = a.S;
  __shared__ double s_mean[kNormMaxDim + 1], s_var[kNormMaxDim + 1];
  const int c = threadIdx.x;
  if (c <= S) {                                       // component c < S: observation c; component S: the return
    const bool is_ret = (c == S);
    const bool on = is_ret ? (a.norm_ret && a.reward != nullptr) : (a.norm_obs && a.obs != nullptr);
    double mean = a.stats_cur[c], var = a.stats_cur[(kNormMaxDim + 1) + c];
    double count = a.stats_cur[2 * (kNormMaxDim + 1) + (is_ret ? 1 : 0)];
    if (on && a.update) merge_moments(mean, var, count, a.sums_cur[c], a.sums_cur[(kNormMaxDim + 1) + c], (double)a.n, mean, var, count);
    s_mean[c] = mean; s_var[c] = var;
    if (blockIdx.x == 0) {                             // the statistics every later step starts from
      a.stats_next[c] = mean; a.stats_next[(kNormMaxDim + 1) + c] = var;
      if (c == 0) a.stats_next[2 * (kNormMaxDim + 1)] = count;          // component 0 carries the observation count
      if (is_ret) a.stats_next[2 * (kNormMaxDim + 1) + 1] = count;
      a.sums_next[c] = 0.0; a.sums_next[(kNormMaxDim + 1) + c] = 0.0;
    }
  }
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  if (a.obs && a.obs_out) {
    for (int k = 0; k < S; ++k) {
      const float x = a.obs[e * S + k];
      float y = x;
      if (a.norm_obs) {
        // (obs - mean) / sqrt(var + eps) in float64, as NumPy's promotion does there, then clipped
        const double z = ((double)x - s_mean[k]) / ::sqrt(s_var[k] + a.epsilon);
        y = (float)fmin(fmax(z, -(double)a.clip_obs), (double)a.clip_obs);
      }
      a.obs_out[e * S + k] = y;
    }
  }
  if (a.reward && a.reward_out) {
    float r = a.reward[e];
    if (a.norm_ret) {
      const double z = (double)r / ::sqrt(s_var[S] + a.epsilon);
      r = (float)fmin(fmax(z, -(double)a.clip_rew), (double)a.clip_rew);
    }
    a.reward_out[e] = r;
    if (a.done && a.done[e]) a.ret[e] = 0.0;            // self.ret[news] = 0.
  }
}

}  // namespace rmv
