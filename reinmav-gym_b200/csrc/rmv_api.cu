// rmv_api.cu — implementation of the C-ABI declared in include/rmv.h.
// Host side only decides launch shapes and owns buffers; all arithmetic is in rmv_kernels.cuh.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <limits>
#include <new>
#include <string>

#include "rmv.h"
#include "rmv_kernels.cuh"
#include "rmv_control.cuh"
#include "rmv_step_tma.cuh"
#include "rmv_stream.cuh"
#include "rmv_norm.cuh"
#include "rmv_reinmav.cuh"

using namespace rmv;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define RMV_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t err__ = (expr);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return fail(err__ == cudaErrorMemoryAllocation ? RMV_ENOMEM : RMV_ECUDA, "%s: %s (%s:%d)", \
                  #expr, cudaGetErrorString(err__), __FILE__, __LINE__);                    \
  } while (0)

#define RMV_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) return fail(RMV_EINVAL, __VA_ARGS__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};
#define RMV_ON_DEVICE(dev)  \
  DeviceGuard guard__(dev); \
  if (!guard__.ok) return fail(RMV_ECUDA, "cannot select CUDA device %d: %s", (dev), cudaGetErrorString(cudaGetLastError()))

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct rmv_env {
  int kind = 0, device = 0, S = 0, A = 0;
  int64_t n = 0, ld = 0, env_id_offset = 0;
  uint64_t seed = 0;
  uint32_t flags = 0, step_index = 0;
  int last_variant = -1;   // RMV_VARIANT_* of the most recent step launch (rmv_last_variant)
  bool rows = false;       // RMV_STATE_ROWS: state is [n_pad][S] rows, stepped in place by the row-state kernel
  double sim_time = 0.0;   // reinmav-v0: self.t (reinmav_env.py:73,119)
  rmv_params params{};
  DevParams dp{};
  float* state = nullptr;
  uint32_t* episode = nullptr;
  // tracking
  float* ep_return = nullptr;
  uint32_t* ep_start = nullptr;
  float* last_return = nullptr;
  uint32_t* last_length = nullptr;
  double* cum_return = nullptr;
  unsigned long long* cum_length = nullptr;
  uint32_t* fin_count = nullptr;
  double* stats_dev = nullptr;
  int64_t* gather_idx = nullptr;   // scratch of rmv_gather_episodes (lazy)
  float* gather_ret = nullptr;
  uint32_t* gather_len = nullptr;
  int64_t gather_cap = 0;
  // staging for the *_host entry points (lazy)
  float* h_act = nullptr;
  float* h_rew = nullptr;
  uint8_t* h_done = nullptr;
  float* h_obs = nullptr;
  // pipeline of the *_host entry points (lazy)
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pipe_kernel[3] = {nullptr, nullptr, nullptr};   // after the last step kernel queued on pipe[i]
  cudaStream_t pipe_small = nullptr;                          // reward / done leave on their own stream, once per call
  cudaEvent_t pipe_small_done = nullptr;
  cudaEvent_t pipe_start = nullptr;
  int host_chunks = 0;      // 0 = auto
  // tuning
  int vec = 0;     // 0 = auto
  int block = 0;   // 0 = auto (default_block)
  int prefetch = -1;       // -1 auto, 0 / 1: episode-word prefetch of the one-env-per-thread step kernel
  int l2_persist = -1;     // -1 auto (= off), 0 / 1: keep the state resident in L2 between steps (access-policy window)
  int l2_ready = 0;        // window active
  cudaAccessPolicyWindow l2_window{};
  int use_tma = 0;         // 1: TMA-pipelined persistent step kernel where it applies
  int stream_warps_per_sm = 0;   // 0 = as many as fit; else cap on resident warps per SM (sweeps)
  int stream_block = 0;    // 0 = auto; else 32, 64 or 128 threads per CTA of the stream kernel
  int stream_apre = -1;    // -1 auto, 0 / 1: fetch the warp's next action while the current unit is computed
  // rmv_step_many(fused = 0): the K single-step launches of a repeated call are captured once in a CUDA graph and replayed
  struct StepGraph {
    int n_steps = 0, layout = 0;
    const float* act = nullptr; float* rew = nullptr; uint8_t* done = nullptr;
    int64_t act_ld = 0, out_ld = 0;
    uint64_t epoch = 0, stamp = 0;
    cudaGraphExec_t exec = nullptr;
  };
  StepGraph graphs[4];
  StepGraph seen[4];               // recently seen call shapes: a call is captured the second time it is seen
  uint64_t config_epoch = 1;       // bumped by every setter that changes what a launch looks like
  uint64_t graph_clock = 0;
  int graphs_off = 0;              // capture failed once: plain launches from then on
  cudaStream_t capture_stream = nullptr;
  int tma_ctas_per_sm = 0; // 0 = as many as fit
};

static int g_use_graph = (getenv("RMV_NO_GRAPH") == nullptr) ? 1 : 0;   // developer switch
static int g_host_batch_small = getenv("RMV_HOST_BATCH_SMALL") ? atoi(getenv("RMV_HOST_BATCH_SMALL")) : 1;   // developer switch (A/B)
static int g_use_stream = (getenv("RMV_NO_STREAM") == nullptr) ? 1 : 0; // developer switch: rmv_step_many(mode 0) falls back to launches
static int g_force_stream = (getenv("RMV_FORCE_STREAM") != nullptr) ? 1 : 0; // developer switch: the stream at every size (sweeps)
static int g_ctas_cap = getenv("RMV_CTAS_PER_SM") ? atoi(getenv("RMV_CTAS_PER_SM")) : 0;
static int enable_l2_window(rmv_env* e);
static void disable_l2_window(rmv_env* e);

// developer switch for A/B measurements (tools/sweep.py); the default is the product configuration
static int g_use_pdl = (getenv("RMV_NO_PDL") == nullptr) ? 1 : 0;

// per-device caches (function attributes, occupancy, the persisting-L2 carve-out) are indexed by device ordinal
constexpr int kMaxDevices = 64;
static inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

static inline bool aligned(const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static inline unsigned blocks_for(int64_t n, int64_t block) { return (unsigned)((n + block - 1) / block); }

static const int kS[RMV_NUM_KINDS] = {5, 9, 10, 16, 13};
static const int kA[RMV_NUM_KINDS] = {2, 2, 4, 4, 0};

static DevParams to_dev(const rmv_params& p) {
  DevParams d;
  d.mass = p.mass; d.load_mass = p.load_mass; d.dt = p.dt; d.neg_gravity = -p.gravity;
  d.tether_len = p.tether_len; d.pos_done = p.pos_done; d.vel_done = p.vel_done;
  d.thrust_scale = p.thrust_scale; d.first_done_reward = p.first_done_reward;
  d.later_done_reward = p.later_done_reward; d.thrust_clamp = p.thrust_clamp;
  d.reward_scale = p.reward_scale;
  finish_params(d);
  return d;
}

extern "C" RMV_API int rmv_abi_version(void) { return RMV_ABI_VERSION; }
extern "C" RMV_API const char* rmv_last_error(void) { return g_err.c_str(); }

extern "C" RMV_API int rmv_kind_dims(int kind, int* state_dim, int* action_dim) {
  RMV_REQUIRE(kind >= 0 && kind < RMV_NUM_KINDS, "unknown env kind %d", kind);
  if (state_dim) *state_dim = kS[kind];
  if (action_dim) *action_dim = kA[kind];
  return RMV_OK;
}

extern "C" RMV_API int rmv_default_params(int kind, rmv_params* out) {
  RMV_REQUIRE(kind >= 0 && kind < RMV_NUM_KINDS, "unknown env kind %d", kind);
  RMV_REQUIRE(out != nullptr, "out is NULL");
  rmv_params p;
  p.mass = 1.0f; p.load_mass = 0.1f; p.dt = 0.01f; p.gravity = 9.8f; p.tether_len = 0.0f;
  p.pos_done = 3.0f; p.vel_done = 10.0f; p.thrust_scale = 1.0f; p.thrust_clamp = 0;
  p.first_done_reward = 1.0f; p.later_done_reward = 0.0f; p.reward_scale = 1.0f;
  if (kind == RMV_QUAD2D) { p.thrust_scale = 10.0f; p.thrust_clamp = 1; }
  if (kind == RMV_QUAD2D_SLUNG) { p.tether_len = 0.5f; p.pos_done = 2.0f; }
  if (kind == RMV_QUAD3D_SLUNG) { p.tether_len = 1.5f; }
  if (kind == RMV_REINMAV) { p.mass = 0.18f; p.gravity = 9.81f; p.load_mass = 0.0f; }   // informational (reinmav_env.py:55-57)
  *out = p;
  return RMV_OK;
}

extern "C" RMV_API int64_t rmv_padded_ld(int64_t n_envs) { return n_envs <= 0 ? 0 : ((n_envs + 127) / 128) * 128; }

extern "C" RMV_API int rmv_action_bounds(int kind, float* low, float* high) {
  RMV_REQUIRE(kind >= 0 && kind < RMV_NUM_KINDS, "unknown env kind %d", kind);
  if (low) *low = (kind == RMV_QUAD3D || kind == RMV_REINMAV) ? 0.0f : -10.0f;
  if (high) *high = (kind == RMV_REINMAV) ? 0.0f : 10.0f;
  return RMV_OK;
}

template <typename T>
static cudaError_t alloc_zero(T** p, size_t count) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemset(*p, 0, count * sizeof(T));
}

static void free_all(rmv_env* e) {
  cudaFree(e->state); cudaFree(e->episode); cudaFree(e->ep_return); cudaFree(e->ep_start);
  cudaFree(e->last_return); cudaFree(e->last_length); cudaFree(e->cum_return);
  cudaFree(e->cum_length); cudaFree(e->fin_count); cudaFree(e->stats_dev);
  cudaFree(e->gather_idx); cudaFree(e->gather_ret); cudaFree(e->gather_len);
  cudaFree(e->h_act); cudaFree(e->h_rew); cudaFree(e->h_done); cudaFree(e->h_obs);
  for (int i = 0; i < 3; ++i) {
    if (e->pipe[i]) cudaStreamDestroy(e->pipe[i]);
    if (e->pipe_done[i]) cudaEventDestroy(e->pipe_done[i]);
    if (e->pipe_kernel[i]) cudaEventDestroy(e->pipe_kernel[i]);
  }
  if (e->pipe_small) cudaStreamDestroy(e->pipe_small);
  if (e->pipe_small_done) cudaEventDestroy(e->pipe_small_done);
  if (e->pipe_start) cudaEventDestroy(e->pipe_start);
  for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->capture_stream) cudaStreamDestroy(e->capture_stream);
}

extern "C" RMV_API int rmv_create(rmv_env** out, int kind, int64_t n_envs, int64_t env_id_offset, int device,
                          uint64_t seed, const rmv_params* params, uint32_t flags) {
  RMV_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  RMV_REQUIRE(kind >= 0 && kind < RMV_NUM_KINDS, "unknown env kind %d", kind);
  RMV_REQUIRE(n_envs > 0, "n_envs must be positive (got %lld)", (long long)n_envs);
  RMV_REQUIRE(env_id_offset >= 0, "env_id_offset must be >= 0");
  RMV_REQUIRE(n_envs <= (int64_t)1 << 30, "n_envs must be <= 2^30 per handle (32-bit env indices inside a launch)");
  RMV_REQUIRE(!((flags & RMV_STATE_ROWS) && kind == RMV_REINMAV), "RMV_STATE_ROWS is not available for reinmav-v0");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count <= 0)
    return fail(RMV_ECUDA, "no CUDA device available (%s); librmv has no CPU path", cudaGetErrorString(ce));
  RMV_REQUIRE(device >= 0 && device < count, "device %d out of range (have %d)", device, count);
  RMV_ON_DEVICE(device);
  rmv_env* e = new (std::nothrow) rmv_env();
  if (!e) return fail(RMV_ENOMEM, "host allocation failed");
  e->kind = kind; e->device = device; e->S = kS[kind]; e->A = kA[kind];
  e->n = n_envs; e->ld = rmv_padded_ld(n_envs); e->env_id_offset = env_id_offset;
  e->seed = seed; e->flags = flags;
  e->rows = (flags & RMV_STATE_ROWS) != 0;
  if (params) e->params = *params; else rmv_default_params(kind, &e->params);
  e->dp = to_dev(e->params);
  cudaError_t err = alloc_zero(&e->state, (size_t)e->S * e->ld);
  if (err == cudaSuccess) err = alloc_zero(&e->episode, (size_t)e->ld);
  if (err == cudaSuccess && (flags & RMV_TRACK_EPISODES)) {
    err = alloc_zero(&e->ep_return, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->ep_start, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->last_return, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->last_length, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->cum_return, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->cum_length, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->fin_count, (size_t)e->ld);
    if (err == cudaSuccess) err = alloc_zero(&e->stats_dev, 4);
  }
  if (err == cudaSuccess && kind == RMV_REINMAV) {
    k_reinmav_init<<<blocks_for(e->n, 256), 256>>>(e->state, e->n);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaDeviceSynchronize();   // memsets done before any user stream touches the buffers
  if (err != cudaSuccess) {
    free_all(e);
    delete e;
    return fail(err == cudaErrorMemoryAllocation ? RMV_ENOMEM : RMV_ECUDA, "rmv_create: %s", cudaGetErrorString(err));
  }
  *out = e;
  return RMV_OK;
}

extern "C" RMV_API int rmv_destroy(rmv_env* env) {
  if (!env) return RMV_OK;
  {
    RMV_ON_DEVICE(env->device);
    cudaDeviceSynchronize();
    disable_l2_window(env);
    free_all(env);
  }
  delete env;
  return RMV_OK;
}

extern "C" RMV_API int rmv_dims(const rmv_env* env, int* state_dim, int* action_dim, int64_t* n_envs, int64_t* ld) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  if (state_dim) *state_dim = env->S;
  if (action_dim) *action_dim = env->A;
  if (n_envs) *n_envs = env->n;
  if (ld) *ld = env->ld;
  return RMV_OK;
}

extern "C" RMV_API int rmv_get_params(const rmv_env* env, rmv_params* out) {
  RMV_REQUIRE(env != nullptr && out != nullptr, "NULL argument");
  *out = env->params;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_seed(rmv_env* env, uint64_t seed) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  env->seed = seed;
  env->config_epoch++;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_host_chunks(rmv_env* env, int chunks) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(chunks >= 0 && chunks <= 1024, "chunks must be in [0, 1024] (0 = auto)");
  env->host_chunks = chunks;
  return RMV_OK;
}

extern "C" RMV_API int rmv_last_variant(const rmv_env* env, int* variant) {
  RMV_REQUIRE(env != nullptr && variant != nullptr, "NULL argument");
  *variant = env->last_variant;
  return RMV_OK;
}

extern "C" RMV_API int rmv_get_time(const rmv_env* env, double* t) {
  RMV_REQUIRE(env != nullptr && t != nullptr, "NULL argument");
  *t = env->sim_time;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_time(rmv_env* env, double t) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  env->sim_time = t;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_kernel(rmv_env* env, int use_tma, int ctas_per_sm) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 32, "ctas_per_sm must be in [0, 32]");
  env->use_tma = use_tma ? 1 : 0;
  env->tma_ctas_per_sm = ctas_per_sm;
  env->config_epoch++;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_stream(rmv_env* env, int warps_per_sm, int block, int action_prefetch) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(action_prefetch >= -1 && action_prefetch <= 1, "action_prefetch must be -1 (auto), 0 or 1");
  env->stream_apre = action_prefetch;
  RMV_REQUIRE(warps_per_sm >= 0 && warps_per_sm <= 64, "warps_per_sm must be in [0, 64] (0 = as many as fit)");
  RMV_REQUIRE(block == 0 || block == 32 || block == 64 || block == 128, "block must be 0 (auto), 32, 64 or 128");
  env->stream_warps_per_sm = warps_per_sm; env->stream_block = block;
  env->config_epoch++;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_l2_persist(rmv_env* env, int mode) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(mode >= -1 && mode <= 1, "mode must be -1 (auto = off), 0 or 1");
  RMV_ON_DEVICE(env->device);
  env->l2_persist = mode;
  env->config_epoch++;
  if (mode == 1) return env->l2_ready ? RMV_OK : enable_l2_window(env);
  disable_l2_window(env);
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_prefetch(rmv_env* env, int mode) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(mode >= -1 && mode <= 1, "mode must be -1 (auto), 0 or 1");
  env->prefetch = mode;
  env->config_epoch++;
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_tuning(rmv_env* env, int vec, int block) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(vec == 0 || vec == 1 || vec == 2 || vec == 4, "vec must be 0 (auto), 1, 2 or 4");
  RMV_REQUIRE(block == 0 || block == 128 || block == 256 || block == 512, "block must be 0 (auto), 128, 256 or 512 threads per CTA");
  env->vec = vec; env->block = block;
  env->config_epoch++;
  return RMV_OK;
}

extern "C" RMV_API int rmv_state_ptr(rmv_env* env, float** state_dev, uint32_t** episode_dev) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  if (state_dev) *state_dev = env->state;
  if (episode_dev) *episode_dev = env->episode;
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// state in / out
// ------------------------------------------------------------------------------------------
static int ensure_staging(rmv_env* e) {
  if (!e->h_act) RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->h_act), (size_t)e->n * (e->A > 0 ? e->A : 1) * sizeof(float)));
  if (!e->h_rew) RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->h_rew), (size_t)e->n * sizeof(float)));
  if (!e->h_done) RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->h_done), (size_t)e->n));
  if (!e->h_obs) RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->h_obs), (size_t)e->n * e->S * sizeof(float)));
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_state(rmv_env* env, const float* aos, int is_device, void* stream) {
  RMV_REQUIRE(env != nullptr && aos != nullptr, "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* src = aos;
  if (!is_device && env->rows) {
    RMV_CUDA(cudaMemcpyAsync(env->state, aos, (size_t)env->n * env->S * sizeof(float), cudaMemcpyHostToDevice, st));
    RMV_CUDA(cudaStreamSynchronize(st));
    return RMV_OK;
  }
  if (!is_device) {
    int rc = ensure_staging(env);
    if (rc) return rc;
    RMV_CUDA(cudaMemcpyAsync(env->h_obs, aos, (size_t)env->n * env->S * sizeof(float), cudaMemcpyHostToDevice, st));
    src = env->h_obs;
  }
  if (env->rows) {
    if (src != env->state)
      RMV_CUDA(cudaMemcpyAsync(env->state, src, (size_t)env->n * env->S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    k_rows_to_tiles<<<blocks_for(env->n, 256), 256, 0, st>>>(src, env->state, env->n, env->S);
    RMV_CUDA(cudaGetLastError());
  }
  if (!is_device) RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

extern "C" RMV_API int rmv_get_state(rmv_env* env, float* aos, int is_device, void* stream) {
  RMV_REQUIRE(env != nullptr && aos != nullptr, "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dst = aos;
  if (!is_device && !env->rows) {
    int rc = ensure_staging(env);
    if (rc) return rc;
    dst = env->h_obs;
  }
  if (env->rows) {
    if (aos != env->state)
      RMV_CUDA(cudaMemcpyAsync(aos, env->state, (size_t)env->n * env->S * sizeof(float),
                               is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!is_device) RMV_CUDA(cudaStreamSynchronize(st));
    return RMV_OK;
  }
  k_tiles_to_rows<<<blocks_for(env->n, 256), 256, 0, st>>>(env->state, dst, env->n, env->S);
  RMV_CUDA(cudaGetLastError());
  if (!is_device) {
    RMV_CUDA(cudaMemcpyAsync(aos, dst, (size_t)env->n * env->S * sizeof(float), cudaMemcpyDeviceToHost, st));
    RMV_CUDA(cudaStreamSynchronize(st));
  }
  return RMV_OK;
}

extern "C" RMV_API int rmv_set_episode(rmv_env* env, const uint32_t* episode, int is_device, void* stream) {
  RMV_REQUIRE(env != nullptr && episode != nullptr, "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RMV_CUDA(cudaMemcpyAsync(env->episode, episode, (size_t)env->n * sizeof(uint32_t),
                           is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  if (!is_device) RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

extern "C" RMV_API int rmv_get_episode(rmv_env* env, uint32_t* episode, int is_device, void* stream) {
  RMV_REQUIRE(env != nullptr && episode != nullptr, "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RMV_CUDA(cudaMemcpyAsync(episode, env->episode, (size_t)env->n * sizeof(uint32_t),
                           is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!is_device) RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// reset
// ------------------------------------------------------------------------------------------
template <int KIND>
static void launch_reset(rmv_env* e, const uint8_t* mask, float* obs, cudaStream_t st) {
  if (e->rows)
    k_reset<KIND, true><<<blocks_for(e->n, 256), 256, 0, st>>>(e->state, e->episode, e->n, mask, e->seed, e->env_id_offset, obs,
                                                              e->ep_return, e->ep_start, e->step_index);
  else
    k_reset<KIND, false><<<blocks_for(e->n, 256), 256, 0, st>>>(e->state, e->episode, e->n, mask, e->seed, e->env_id_offset, obs,
                                                               e->ep_return, e->ep_start, e->step_index);
}

extern "C" RMV_API int rmv_reset(rmv_env* env, const uint8_t* mask_dev, float* obs_aos_dev, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (env->kind == RMV_REINMAV) {
    // ReinmavEnv.reset() returns the CURRENT state and resets nothing (reinmav_env.py:348-351)
    if (obs_aos_dev) {
      k_tiles_to_rows<<<blocks_for(env->n, 256), 256, 0, st>>>(env->state, obs_aos_dev, env->n, env->S);
      RMV_CUDA(cudaGetLastError());
    }
    return RMV_OK;
  }
  switch (env->kind) {
    case RMV_QUAD2D: launch_reset<QUAD2D>(env, mask_dev, obs_aos_dev, st); break;
    case RMV_QUAD2D_SLUNG: launch_reset<QUAD2D_SLUNG>(env, mask_dev, obs_aos_dev, st); break;
    case RMV_QUAD3D: launch_reset<QUAD3D>(env, mask_dev, obs_aos_dev, st); break;
    default: launch_reset<QUAD3D_SLUNG>(env, mask_dev, obs_aos_dev, st); break;
  }
  RMV_CUDA(cudaGetLastError());
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------
// L2 residency of the state between steps (opt-in, rmv_set_l2_persist).  The state is the one buffer every step
// re-reads (and re-writes in place); actions, rewards and done flags stream through once.  Left to LRU, the streams
// push the state out of the 126 MB L2 once state + streams of one step exceed it (quadrotor3d-slungload at N = 2^20:
// 67 MB state + 89 MB streams), and every byte then moves through HBM.  An access-policy window on the state
// (persisting on hit, streaming on miss) pins it.  Measured on B200 (profiles/r1b_sweep_l2_persist.txt): +10-13 % in
// exactly that band (state fits the carve-out, state + streams do not fit L2); no effect when everything fits anyway;
// 2.3x SLOWER with a hit ratio < 1 (state larger than the carve-out: the persisting subset thrashes), which is why that
// case is refused; slower with observation rows (the carve-out is taken from the streams' share).  The carve-out is a
// device-wide limit that also shrinks the L2 available to everything else on the GPU (the policy network) — hence opt-in.
// The carve-out (cudaLimitPersistingL2CacheSize) is ONE device-wide number shared by every handle on the device:
// it is kept at the largest request of the handles that currently hold a window and released by the last one.
struct L2Carve { int users = 0; size_t bytes = 0; };
static L2Carve g_l2_carve[kMaxDevices];

static int enable_l2_window(rmv_env* e) {
  int max_persist = 0, max_window = 0;
  RMV_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device));
  RMV_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, e->device));
  const size_t state_bytes = (size_t)e->S * e->ld * sizeof(float);
  if (max_persist <= 0 || max_window <= 0)
    return fail(RMV_EINVAL, "this device has no persisting L2 carve-out");
  // a window slightly larger than the carve-out is fine (the few lines that miss just stream); a partial window thrashes
  const size_t carve = state_bytes < (size_t)max_persist ? state_bytes : (size_t)max_persist;
  if (state_bytes > (size_t)max_window || (double)carve < 0.9 * (double)state_bytes)
    return fail(RMV_EINVAL, "state (%zu bytes) exceeds the persisting L2 carve-out (%d bytes) / window (%d bytes): a partial "
                            "window thrashes (measured 2.3x slower), refused", state_bytes, max_persist, max_window);
  L2Carve& dev_carve = g_l2_carve[e->device < kMaxDevices ? e->device : 0];
  if (carve > dev_carve.bytes) {                   // grow only: another handle's window stays covered
    RMV_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    dev_carve.bytes = carve;
  }
  dev_carve.users += 1;
  e->l2_window = cudaAccessPolicyWindow{};
  e->l2_window.base_ptr = e->state;
  e->l2_window.num_bytes = state_bytes;
  e->l2_window.hitRatio = (float)((double)carve / (double)state_bytes);
  e->l2_window.hitProp = cudaAccessPropertyPersisting;
  e->l2_window.missProp = cudaAccessPropertyStreaming;
  e->l2_ready = 1;
  return RMV_OK;
}
static void disable_l2_window(rmv_env* e) {
  if (!e->l2_ready) return;
  e->l2_ready = 0;
  e->l2_window = cudaAccessPolicyWindow{};
  L2Carve& dev_carve = g_l2_carve[e->device < kMaxDevices ? e->device : 0];
  if (dev_carve.users > 0) dev_carve.users -= 1;
  if (dev_carve.users == 0) {                                  // last window on this device: un-pin and give the L2 back
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    dev_carve.bytes = 0;
  }
  cudaGetLastError();
}
// the window is attached to launches without observation rows only (measured: rows + carve-out is a net loss)
static const cudaAccessPolicyWindow* l2_window_of(const rmv_env* e, const StepArgs& args) {
  return (e->l2_ready && args.obs == nullptr) ? &e->l2_window : nullptr;
}

// Launch shape of the one-env-per-thread step kernel when rmv_set_tuning left it on auto.  Measured on B200
// (profiles/r1b_sweep_block_prefetch.txt): what matters is whether one step's working set stays in the 126 MB
// L2 between launches ("small": L2-bandwidth / issue bound) or streams from HBM ("big": latency bound for the
// small-state kinds, where fetching the `episode` word together with the state takes a dependent memory round
// trip out of every CTA that contains a finished env).
static int g_prefetch_override = getenv("RMV_PREFETCH") ? atoi(getenv("RMV_PREFETCH")) : -1;   // developer switches
static int g_block_override = getenv("RMV_BLOCK") ? atoi(getenv("RMV_BLOCK")) : 0;            // (tools/sweep.py, parity runs)
static bool working_set_is_big(int kind, int64_t n) {
  static const int64_t bytes[4] = {53, 85, 101, 149};
  return n * bytes[kind] > ((int64_t)96 << 20);
}
static int default_block(int kind, int64_t n) {
  if (g_block_override == 128 || g_block_override == 256 || g_block_override == 512) return g_block_override;
  return (kind == QUAD2D && !working_set_is_big(kind, n)) ? 256 : 128;
}
static bool default_prefetch(int kind, int64_t n) {
  if (g_prefetch_override >= 0) return g_prefetch_override != 0;
  if (kind == QUAD2D) return true;
  if (kind == QUAD3D) return false;
  return working_set_is_big(kind, n);
}

// Where the persistent step stream beats one launch per step (profiles/r2_sweep_stream_vs_launches.txt).  The stream
// wins wherever launch boundaries matter — everything that is L2-resident or a little larger — and loses once a step
// streams several hundred MB through HBM: its warps drift apart (that is the point of it), so DRAM sees thousands of
// independent 5 KB accesses instead of one grid sweeping the array in address order (0.68 vs 1.0 of the copy bandwidth
// at N = 2^23).  Above the threshold RMV_MANY_STREAM issues launches.
static bool stream_pays(int kind, int64_t n) {
  static const int64_t bytes[4] = {53, 85, 101, 149};
  static const int64_t limit_mb[4] = {256, 32, 256, 256};      // quadrotor2d-slungload gains up to N = 2^18 only
  return kind >= 0 && kind < 4 && n * bytes[kind] <= (limit_mb[kind] << 20);
}

// Row-state handles: the unit's rows go through shared memory both ways, which costs the stream more than it costs the
// row-state kernel once the launch boundary stops mattering — measured (profiles/r2c_sweep_rows_stream.txt): 1.66 against
// 3.14 us per quadrotor3d step at N = 65536, 3.79 / 5.60 at 2^18, 13.1 / 13.6 at 2^20 (15.5 / 15.1 with episode tracking);
// quadrotor3d-slungload 6.75 / 7.98 at 2^18 but 27.5 / 24.3 at 2^20; quadrotor2d 9.49 / 10.45 at 2^20 (55 MB per step).
static bool stream_pays_for(const rmv_env* e) {
  static const int64_t bytes[4] = {53, 85, 101, 149};
  if (!stream_pays(e->kind, e->n)) return false;
  return !e->rows || e->n * bytes[e->kind] <= ((int64_t)64 << 20);
}

extern "C" RMV_API int rmv_stream_pays(int kind, int64_t n_envs) {
  return (g_use_stream && stream_pays(kind, n_envs)) ? 1 : 0;
}

extern "C" RMV_API int rmv_launch_shape(int kind, int64_t n_envs, int* block, int* prefetch) {
  RMV_REQUIRE(kind >= 0 && kind < RMV_REINMAV, "kind %d has no one-env-per-thread step kernel", kind);
  RMV_REQUIRE(n_envs > 0, "n_envs must be positive");
  if (block) *block = default_block(kind, n_envs);
  if (prefetch) *prefetch = default_prefetch(kind, n_envs) ? 1 : 0;
  return RMV_OK;
}

template <int KIND, int V, bool OBS, bool TRACK, int BLOCK, bool PREF, bool CTRL = false>
static cudaError_t launch_step_kvf(const StepArgs& args, const cudaAccessPolicyWindow* win, cudaStream_t st) {
  constexpr int S = Dims<KIND>::S;
  constexpr int SP = V == 1 ? row_pitch<S>() : ((S % 2 == 0) ? S + 1 : S);
  size_t smem = OBS ? (size_t)BLOCK * SP * sizeof(float) : 0;           // [warps][32][pitch] transpose tiles
  if (g_ctas_cap > 0) {               // developer experiment: cap resident CTAs per SM by padding dynamic shared memory
    const size_t per = (size_t)(227 * 1024) / (size_t)g_ctas_cap - 2048;
    if (per > smem) smem = per & ~(size_t)127;
  }
  if (smem > 48 * 1024) {
    static bool raised[kMaxDevices] = {};      // per instantiation AND per device (function attributes are per context)
    const int dev = current_device();
    if (!raised[dev]) {
      cudaError_t e = cudaFuncSetAttribute(k_step<KIND, V, OBS, TRACK, BLOCK, PREF, CTRL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      raised[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks_for(args.n, (int64_t)BLOCK * V));
  cfg.blockDim = dim3(BLOCK);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this launch's ramp with the previous grid's tail
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (win) {
    attr[na].id = cudaLaunchAttributeAccessPolicyWindow;               // state stays L2-resident between steps
    attr[na].val.accessPolicyWindow = *win;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, k_step<KIND, V, OBS, TRACK, BLOCK, PREF, CTRL>, args);
}

// fused control() + step(): one env per thread, 128-thread CTAs
template <int KIND>
static cudaError_t launch_control_step_k(const StepArgs& args, cudaStream_t st) {
  const bool track = args.ep_return != nullptr;
  if (args.obs) return track ? launch_step_kvf<KIND, 1, true, true, 128, false, true>(args, nullptr, st)
                             : launch_step_kvf<KIND, 1, true, false, 128, false, true>(args, nullptr, st);
  return track ? launch_step_kvf<KIND, 1, false, true, 128, false, true>(args, nullptr, st)
               : launch_step_kvf<KIND, 1, false, false, 128, false, true>(args, nullptr, st);
}

template <int KIND, bool TRACK, int BLOCK, bool CTRL = false>
static cudaError_t launch_step_rows_kb(const StepArgs& args, cudaStream_t st) {
  constexpr int S = Dims<KIND>::S;
  constexpr int SP = row_pitch<S>();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks_for(args.n, BLOCK));
  cfg.blockDim = dim3(BLOCK);
  cfg.dynamicSmemBytes = (size_t)BLOCK * SP * sizeof(float);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_step_rows<KIND, TRACK, BLOCK, CTRL>, args);
}
template <int KIND>
static cudaError_t launch_step_rows_k(const rmv_env* e, const StepArgs& args, cudaStream_t st, bool ctrl = false) {
  const bool track = args.ep_return != nullptr;
  if (ctrl) return track ? launch_step_rows_kb<KIND, true, 128, true>(args, st) : launch_step_rows_kb<KIND, false, 128, true>(args, st);
  const int block = e->block ? e->block : (KIND == QUAD2D ? 256 : 128);
  if (block == 256) return track ? launch_step_rows_kb<KIND, true, 256>(args, st) : launch_step_rows_kb<KIND, false, 256>(args, st);
  return track ? launch_step_rows_kb<KIND, true, 128>(args, st) : launch_step_rows_kb<KIND, false, 128>(args, st);
}

// TMA-pipelined persistent variant (V = 1, row-per-env actions, 16-byte aligned action buffer)
constexpr int kTmaStages = 3;
template <int KIND, bool OBS, bool TRACK>
static cudaError_t launch_step_tma(const rmv_env* e, const StepArgs& args, cudaStream_t st) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  constexpr int SP = (S % 2 == 0) ? S + 1 : S;
  const size_t smem = (size_t)kTmaStages * (S + A) * kTile * 4 + (OBS ? (size_t)kWarps * 32 * SP * sizeof(float) : 0);
  static int resident_of[kMaxDevices] = {}, sms_of[kMaxDevices] = {};      // per instantiation and per device
  const int dev = current_device();
  if (resident_of[dev] == 0) {
    cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_step_tma<KIND, OBS, TRACK, kTmaStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_of[dev], k_step_tma<KIND, OBS, TRACK, kTmaStages>, kBlock, smem);
    if (resident_of[dev] < 1) resident_of[dev] = 1;
    if (sms_of[dev] < 1) sms_of[dev] = 1;
  }
  const int resident = resident_of[dev], sms = sms_of[dev];
  const unsigned tiles = blocks_for(args.n, kTile);
  int per_sm = e->tma_ctas_per_sm > 0 && e->tma_ctas_per_sm < resident ? e->tma_ctas_per_sm : resident;
  const unsigned persistent = (unsigned)sms * (unsigned)per_sm;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles < persistent ? tiles : persistent);
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_step_tma<KIND, OBS, TRACK, kTmaStages>, args);
}

template <int KIND>
static cudaError_t launch_step_tma_k(const rmv_env* e, const StepArgs& args, cudaStream_t st) {
  const bool track = args.ep_return != nullptr;
  if (args.obs) return track ? launch_step_tma<KIND, true, true>(e, args, st) : launch_step_tma<KIND, true, false>(e, args, st);
  return track ? launch_step_tma<KIND, false, true>(e, args, st) : launch_step_tma<KIND, false, false>(e, args, st);
}

template <int KIND, int V, int BLOCK, bool PREF>
static cudaError_t launch_step_kvb(const StepArgs& args, const cudaAccessPolicyWindow* win, cudaStream_t st) {
  const bool track = args.ep_return != nullptr;
  if (args.obs) return track ? launch_step_kvf<KIND, V, true, true, BLOCK, PREF>(args, win, st) : launch_step_kvf<KIND, V, true, false, BLOCK, PREF>(args, win, st);
  return track ? launch_step_kvf<KIND, V, false, true, BLOCK, PREF>(args, win, st) : launch_step_kvf<KIND, V, false, false, BLOCK, PREF>(args, win, st);
}

template <int KIND, int V>
static cudaError_t launch_step_kv(rmv_env* e, const StepArgs& args, cudaStream_t st) {
  if (args.n_steps > 1) {
    const int64_t groups = (args.n + V - 1) / V;
    if constexpr (V == 1) {
      if (e->rows) { k_rollout<KIND, 1, true><<<blocks_for(groups, 128), 128, 0, st>>>(args); return cudaGetLastError(); }
    }
    k_rollout<KIND, V><<<blocks_for(groups, 128), 128, 0, st>>>(args);
    return cudaGetLastError();
  }
  const cudaAccessPolicyWindow* win = l2_window_of(e, args);
  if constexpr (V == 1) {            // CTA size is a tuning knob of the one-env-per-thread kernel only
    const int block = e->block ? e->block : default_block(KIND, args.n);
    const bool pref = e->prefetch >= 0 ? e->prefetch != 0 : default_prefetch(KIND, args.n);
    if (pref) {
      if (block == 512) return launch_step_kvb<KIND, 1, 512, true>(args, win, st);
      if (block == 256) return launch_step_kvb<KIND, 1, 256, true>(args, win, st);
      return launch_step_kvb<KIND, 1, 128, true>(args, win, st);
    }
    if (block == 512) return launch_step_kvb<KIND, 1, 512, false>(args, win, st);
    if (block == 256) return launch_step_kvb<KIND, 1, 256, false>(args, win, st);
  }
  return launch_step_kvb<KIND, V, 128, false>(args, win, st);
}

// ---- the persistent step stream (rmv_stream.cuh): one launch for n_steps steps ----------------------------
// Shape: every warp owns the same number k of 32-env units and all warps are co-resident.
static int g_stream_warps_override = getenv("RMV_STREAM_WARPS") ? atoi(getenv("RMV_STREAM_WARPS")) : 0;
static int g_stream_block_override = getenv("RMV_STREAM_BLOCK") ? atoi(getenv("RMV_STREAM_BLOCK")) : 0;

static void stream_shape_for(int64_t n, int sms, int resident_warps_per_sm, int cap_warps_per_sm, int forced_block,
                             StreamShape* shape, unsigned* grid, unsigned* block, uint32_t unit_envs = 32u) {
  const uint32_t units = (uint32_t)((n + unit_envs - 1) / unit_envs);
  int per_sm = resident_warps_per_sm;
  if (cap_warps_per_sm > 0 && cap_warps_per_sm < per_sm) per_sm = cap_warps_per_sm;
  const uint32_t slots = (uint32_t)sms * (uint32_t)per_sm;
  const uint32_t k = (units + slots - 1) / slots;
  const uint32_t warps = ((units + k - 1) / k + 3u) & ~3u;      // whole tiles of 4 units: constant pointer strides per warp
  // small batches: smaller CTAs spread the few warps evenly over the SMs (a CTA is only a scheduling unit here)
  unsigned b = 128;
  if (warps <= (uint32_t)sms * 24u) b = 32; else if (warps <= (uint32_t)sms * 48u) b = 64;
  if (forced_block == 32 || forced_block == 64 || forced_block == 128) b = (unsigned)forced_block;
  shape->units_per_warp = k; shape->warps = warps;
  shape->limit = units * unit_envs; shape->env_stride = warps * unit_envs;
  *block = b;
  *grid = (warps * 32u + b - 1u) / b;
}

extern "C" RMV_API int rmv_stream_shape(int64_t n_envs, int sms, int resident_warps_per_sm, int* units_per_warp, int* warps,
                                        int* grid, int* block) {
  RMV_REQUIRE(n_envs > 0 && sms > 0 && resident_warps_per_sm > 0, "arguments must be positive");
  StreamShape sh; unsigned g = 0, b = 0;
  stream_shape_for(n_envs, sms, resident_warps_per_sm, 0, 0, &sh, &g, &b);
  if (units_per_warp) *units_per_warp = (int)sh.units_per_warp;
  if (warps) *warps = (int)sh.warps;
  if (grid) *grid = (int)g;
  if (block) *block = (int)b;
  return RMV_OK;
}

static int g_stream_apre = getenv("RMV_STREAM_APRE") ? atoi(getenv("RMV_STREAM_APRE")) : -1;   // developer switch: -1 = by shape

template <int KIND, bool TRACK, bool PREF, bool CTRL = false, bool APRE = false, bool ROWS = false>
static cudaError_t launch_stream_t(rmv_env* e, const StepArgs& args, cudaStream_t st) {
  static int resident_of[kMaxDevices] = {}, sms_of[kMaxDevices] = {};      // per instantiation and per device
  const int dev = current_device();
  if (resident_of[dev] == 0) {
    cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_of[dev], k_step_stream<KIND, TRACK, PREF, CTRL, APRE, ROWS>, 128, 0);
    if (resident_of[dev] < 1) resident_of[dev] = 1;
    if (sms_of[dev] < 1) sms_of[dev] = 1;
  }
  StreamShape shape; unsigned grid = 0, block = 0;
  const int cap = e->stream_warps_per_sm > 0 ? e->stream_warps_per_sm : g_stream_warps_override;
  const int fb = e->stream_block > 0 ? e->stream_block : g_stream_block_override;
  stream_shape_for(args.n, sms_of[dev], resident_of[dev] * 4, cap, fb, &shape, &grid, &block);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (const cudaAccessPolicyWindow* win = l2_window_of(e, args)) {
    attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[na].val.accessPolicyWindow = *win;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, k_step_stream<KIND, TRACK, PREF, CTRL, APRE, ROWS>, args, shape);
}

// quadrotor3d, two envs per thread on the packed float32 pipe (k_step_stream_pair)
static int g_stream_pair = getenv("RMV_STREAM_PAIR") ? atoi(getenv("RMV_STREAM_PAIR")) : -1;   // developer switch: -1 = the policy below
// measured (profiles/r2_sweep_stream_pair.txt): 10.5 vs 10.1 us per step at N = 2^20, 3.15 vs 3.25 at N = 2^18, 1.67 vs 1.32
// at N = 65536 — fewer issue slots (52 % busy instead of 72 %) but half the warps to hide latency with: opt-in (vec = 2)
static bool stream_pair_default(int64_t n) { (void)n; return false; }
static cudaError_t launch_stream_pair(rmv_env* e, const StepArgs& args, cudaStream_t st) {
  static int resident_of[kMaxDevices] = {}, sms_of[kMaxDevices] = {};
  const int dev = current_device();
  if (resident_of[dev] == 0) {
    cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_of[dev], k_step_stream_pair, 128, 0);
    if (resident_of[dev] < 1) resident_of[dev] = 1;
    if (sms_of[dev] < 1) sms_of[dev] = 1;
  }
  StreamShape shape; unsigned grid = 0, block = 0;
  const int cap = e->stream_warps_per_sm > 0 ? e->stream_warps_per_sm : g_stream_warps_override;
  const int fb = e->stream_block > 0 ? e->stream_block : g_stream_block_override;
  stream_shape_for(args.n, sms_of[dev], resident_of[dev] * 4, cap, fb, &shape, &grid, &block, 64u);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_step_stream_pair, args, shape);
}

template <int KIND>
static cudaError_t launch_stream_k(rmv_env* e, const StepArgs& args, cudaStream_t st, bool ctrl = false) {
  const bool track = args.ep_return != nullptr;
  if (e->rows) {                                   // row-state handle: the stream with the row transposes (no episode-word prefetch)
    if (ctrl) return track ? launch_stream_t<KIND, true, false, true, false, true>(e, args, st)
                           : launch_stream_t<KIND, false, false, true, false, true>(e, args, st);
    bool apre = !args.act_soa && args.n <= (int64_t)148 * 32 * 32;        // as below: small batches fetch the next action row early
    if (g_stream_apre >= 0) apre = !args.act_soa && g_stream_apre != 0;
    if (e->stream_apre >= 0) apre = !args.act_soa && e->stream_apre != 0;
    if (apre) return track ? launch_stream_t<KIND, true, false, false, true, true>(e, args, st)
                           : launch_stream_t<KIND, false, false, false, true, true>(e, args, st);
    return track ? launch_stream_t<KIND, true, false, false, false, true>(e, args, st)
                 : launch_stream_t<KIND, false, false, false, false, true>(e, args, st);
  }
  if constexpr (KIND == QUAD3D) {
    const bool fits = !track && !ctrl && !args.act_soa && aligned(args.act, 16) && aligned(args.reward, 8) && aligned(args.done, 2) &&
                      (args.out_ld % 2) == 0 && (args.act_step_stride % 4) == 0;
    const bool want = e->vec == 2 || (e->vec == 0 && (g_stream_pair >= 0 ? g_stream_pair != 0 : stream_pair_default(args.n)));
    if (fits && want) { e->last_variant = RMV_VARIANT_STREAM_PAIR; return launch_stream_pair(e, args, st); }
  }
  if (ctrl) return track ? launch_stream_t<KIND, true, false, true>(e, args, st) : launch_stream_t<KIND, false, false, true>(e, args, st);
  const bool pref = e->prefetch >= 0 ? e->prefetch != 0 : default_prefetch(KIND, args.n);
  // action prefetch (APRE): row-per-env actions, small batches (one unit per warp, at most 32 warps per SM: the action row is
  // the one load of a unit that comes from HBM, and nothing hides it there) — profiles/r2_sweep_stream_apre.txt
  bool apre = !args.act_soa && args.n <= (int64_t)148 * 32 * 32;
  if (g_stream_apre >= 0) apre = !args.act_soa && g_stream_apre != 0;
  if (e->stream_apre >= 0) apre = !args.act_soa && e->stream_apre != 0;
  if (apre) {
    if (track) return pref ? launch_stream_t<KIND, true, true, false, true>(e, args, st) : launch_stream_t<KIND, true, false, false, true>(e, args, st);
    return pref ? launch_stream_t<KIND, false, true, false, true>(e, args, st) : launch_stream_t<KIND, false, false, false, true>(e, args, st);
  }
  if (track) return pref ? launch_stream_t<KIND, true, true>(e, args, st) : launch_stream_t<KIND, true, false>(e, args, st);
  return pref ? launch_stream_t<KIND, false, true>(e, args, st) : launch_stream_t<KIND, false, false>(e, args, st);
}

template <int KIND>
static cudaError_t launch_step_k(rmv_env* e, const StepArgs& args, int vec, cudaStream_t st) {
  if (e->use_tma && vec == 1 && args.n_steps == 1 && !args.act_soa && aligned(args.act, 16) && !e->rows)
    return launch_step_tma_k<KIND>(e, args, st);
  switch (vec) {
    case 4: return launch_step_kv<KIND, 4>(e, args, st);
    case 2: return launch_step_kv<KIND, 2>(e, args, st);
    default: return launch_step_kv<KIND, 1>(e, args, st);
  }
}

// inverse of the reference's inertia matrix (reinmav_env.py:60-63), by cofactors in float64
static void reinmav_inverse_inertia(double out[9]) {
  const double a = 0.00025, c = 2.55e-06, e = 0.000232, g = 2.55e-06, i = 0.0003738;   // [[a,0,c],[0,e,0],[g,0,i]]
  const double det = e * (a * i - c * g);
  out[0] = e * i / det; out[1] = 0.0; out[2] = -c * e / det;
  out[3] = 0.0; out[4] = (a * i - c * g) / det; out[5] = 0.0;
  out[6] = -e * g / det; out[7] = 0.0; out[8] = a * e / det;
}

// Launches the step (or rollout) kernel for the envs [c0, c0 + cn) of the handle (c0 a multiple of the
// 128-env tile; the whole handle when cn == n).  act / rew / done / obs point at env c0's element.
static int launch_step(rmv_env* e, const float* act, int act_layout, int64_t act_ld, int64_t act_step_stride,
                       float* rew, uint8_t* done, int64_t out_ld, float* obs, int auto_reset, int n_steps,
                       cudaStream_t st, int64_t c0 = 0, int64_t cn = -1, bool advance_clock = true, bool stream = false,
                       const CtrlParams* ctrl = nullptr, float* act_out = nullptr) {
  if (cn < 0) cn = e->n;
  if (e->kind == RMV_REINMAV) {
    // no action; n_steps consecutive steps, each np.arange(t, t + dt, 1/5000) substeps (reinmav_env.py:90-98)
    double t = e->sim_time;
    for (int k = 0; k < n_steps; ++k) {
      ReinmavArgs ra;
      ra.state = e->state + (c0 / kTile) * (int64_t)(e->S * kTile);
      ra.reward = rew ? rew + (int64_t)k * out_ld : nullptr;
      ra.done = done ? done + (int64_t)k * out_ld : nullptr;
      ra.obs = obs; ra.n = cn;
      volatile double stop = t + 1.0 / 100, next = t + 1.0 / 5000;
      ra.n_sub = (int)ceil((stop - t) / (1.0 / 5000));     // NumPy's arange length ...
      if (ra.n_sub > kReinmavMaxSub) return fail(RMV_EINVAL, "reinmav substep count %d exceeds %d", ra.n_sub, kReinmavMaxSub);
      ra.delta = next - t;                                  // ... and fill increment
      ra.t0 = t;
      reinmav_inverse_inertia(ra.inv_inertia);
      k_reinmav_step<<<blocks_for(cn, 128), 128, 0, st>>>(ra);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) return fail(RMV_ECUDA, "reinmav step launch failed: %s", cudaGetErrorString(err));
      t = t + 1.0 / 100;
    }
    if (advance_clock) { e->sim_time = t; e->step_index += (uint32_t)n_steps; }
    return RMV_OK;
  }
  if (ctrl) {                                      // fused control() + step(): no action input
    act = nullptr; act_layout = RMV_ACT_AOS; act_ld = 0; act_step_stride = cn * e->A;
    RMV_REQUIRE(aligned(act_out, (size_t)e->A * 4u), "act_out_dev must be aligned to one action row (%d bytes)", e->A * 4);
  } else {
    RMV_REQUIRE(act != nullptr, "act_dev is NULL");
  }
  RMV_REQUIRE(act_layout == RMV_ACT_AOS || act_layout == RMV_ACT_SOA, "bad act_layout %d", act_layout);
  if (act_layout == RMV_ACT_SOA) RMV_REQUIRE(act_ld >= cn, "act_ld (%lld) < n_envs", (long long)act_ld);
  RMV_REQUIRE(aligned(act, 4) && aligned(rew, 4) && aligned(obs, 4), "float buffers must be 4-byte aligned");
  // a row-per-env action is fetched with ONE A*4-byte vector load (float4 / float2): every row must be aligned to it
  if (act_layout == RMV_ACT_AOS && (!aligned(act, (size_t)e->A * 4u) || ((act_step_stride * 4) % ((int64_t)e->A * 4)) != 0))
    return fail(RMV_EALIGN, "row-per-env actions must be aligned to one row (%d bytes); got %p, step stride %lld floats",
                e->A * 4, (const void*)act, (long long)act_step_stride);
  // widest vector width every buffer supports (auto = 1: measured fastest, profiles/r1_sweep_v1_tiled.txt)
  int vec = e->vec ? e->vec : 1;
  auto ok_for = [&](int v) {
    if (v == 1) return true;
    const size_t fb = 4u * v;
    if (!aligned(rew, fb) || !aligned(done, v)) return false;
    if (n_steps > 1 && (out_ld % v) != 0) return false;
    if (act_layout == RMV_ACT_SOA) {
      if (!aligned(act, fb) || (act_ld % v) != 0 || (act_step_stride % v) != 0) return false;
    } else {
      const size_t ab = (size_t)v * e->A * 4u >= 16u ? 16u : 8u;
      if (!aligned(act, ab) || ((act_step_stride * 4) % (int64_t)ab) != 0) return false;
    }
    return true;
  };
  while (vec > 1 && !ok_for(vec)) vec >>= 1;
  if (e->vec && vec != e->vec)
    return fail(RMV_EALIGN, "buffers are not aligned for the forced vector width %d", e->vec);

  StepArgs a;
  auto at = [c0](auto* p) { return p ? p + c0 : p; };
  a.state = e->state + (c0 / kTile) * (int64_t)(e->S * kTile); a.episode = e->episode + c0;
  a.act = act; a.reward = rew; a.done = done; a.obs = obs;
  a.ep_return = at(e->ep_return); a.ep_start = at(e->ep_start); a.last_return = at(e->last_return);
  a.last_length = at(e->last_length); a.cum_return = at(e->cum_return); a.cum_length = at(e->cum_length);
  a.fin_count = at(e->fin_count);
  a.n = cn; a.act_ld = act_ld; a.out_ld = out_ld; a.act_step_stride = act_step_stride;
  a.env_id_offset = e->env_id_offset + c0; a.seed = e->seed; a.step_index = e->step_index;
  philox_key_schedule(e->seed, a.ks);
  a.n_steps = n_steps; a.act_soa = (act_layout == RMV_ACT_SOA); a.auto_reset = auto_reset ? 1 : 0;
  a.P = e->dp;
  a.C = ctrl ? *ctrl : CtrlParams{};
  a.act_out = act_out;
  a.rows_in = nullptr;
  const bool fused_ctrl = ctrl != nullptr;
  cudaError_t err;
  if (e->rows && n_steps > 1) vec = 1;             // the multi-step kernels on row-state handles are one env per thread
  if (e->rows && !stream && n_steps == 1) {
    // RMV_STATE_ROWS: the state buffer is the [n][S] observation matrix; the row-state kernel updates it in place
    RMV_REQUIRE(rew != nullptr && done != nullptr, "reward_dev and done_dev must not be NULL");
    float* const rows = e->state + c0 * e->S;
    a.rows_in = rows; a.obs = rows;
    e->last_variant = RMV_VARIANT_ROWS;
    switch (e->kind) {
      case RMV_QUAD2D: err = launch_step_rows_k<QUAD2D>(e, a, st, fused_ctrl); break;
      case RMV_QUAD2D_SLUNG: err = launch_step_rows_k<QUAD2D_SLUNG>(e, a, st, fused_ctrl); break;
      case RMV_QUAD3D: err = launch_step_rows_k<QUAD3D>(e, a, st, fused_ctrl); break;
      default: err = launch_step_rows_k<QUAD3D_SLUNG>(e, a, st, fused_ctrl); break;
    }
    if (err == cudaSuccess && obs != nullptr && obs != rows)
      err = cudaMemcpyAsync(obs, rows, (size_t)cn * e->S * sizeof(float), cudaMemcpyDeviceToDevice, st);
  } else if (stream) {                             // (a row-state handle too: a.state = e->state + c0 * S either way)
    e->last_variant = RMV_VARIANT_STREAM;          // (launch_stream_k<QUAD3D> may overwrite it with RMV_VARIANT_STREAM_PAIR)
    switch (e->kind) {
      case RMV_QUAD2D: err = launch_stream_k<QUAD2D>(e, a, st, fused_ctrl); break;
      case RMV_QUAD2D_SLUNG: err = launch_stream_k<QUAD2D_SLUNG>(e, a, st, fused_ctrl); break;
      case RMV_QUAD3D: err = launch_stream_k<QUAD3D>(e, a, st, fused_ctrl); break;
      default: err = launch_stream_k<QUAD3D_SLUNG>(e, a, st, fused_ctrl); break;
    }
  } else if (fused_ctrl) {
    if (n_steps != 1) return fail(RMV_EINVAL, "internal: the fused control+step kernel is single-step");
    e->last_variant = RMV_VARIANT_STEP;
    switch (e->kind) {
      case RMV_QUAD2D: err = launch_control_step_k<QUAD2D>(a, st); break;
      case RMV_QUAD2D_SLUNG: err = launch_control_step_k<QUAD2D_SLUNG>(a, st); break;
      case RMV_QUAD3D: err = launch_control_step_k<QUAD3D>(a, st); break;
      default: err = launch_control_step_k<QUAD3D_SLUNG>(a, st); break;
    }
  } else {
    e->last_variant = n_steps > 1 ? RMV_VARIANT_ROLLOUT : ((e->use_tma && vec == 1 && !a.act_soa && aligned(a.act, 16)) ? RMV_VARIANT_TMA : RMV_VARIANT_STEP);
    switch (e->kind) {
      case RMV_QUAD2D: err = launch_step_k<QUAD2D>(e, a, vec, st); break;
      case RMV_QUAD2D_SLUNG: err = launch_step_k<QUAD2D_SLUNG>(e, a, vec, st); break;
      case RMV_QUAD3D: err = launch_step_k<QUAD3D>(e, a, vec, st); break;
      default: err = launch_step_k<QUAD3D_SLUNG>(e, a, vec, st); break;
    }
  }
  if (err != cudaSuccess) return fail(RMV_ECUDA, "step kernel launch failed: %s", cudaGetErrorString(err));
  if (advance_clock) e->step_index += (uint32_t)n_steps;
  return RMV_OK;
}

extern "C" RMV_API int rmv_step(rmv_env* env, const float* act_dev, int act_layout, int64_t act_ld, float* reward_dev,
                        uint8_t* done_dev, float* obs_aos_dev, int auto_reset, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->kind == RMV_REINMAV || (reward_dev != nullptr && done_dev != nullptr), "reward_dev and done_dev must not be NULL");
  RMV_ON_DEVICE(env->device);
  return launch_step(env, act_dev, act_layout, act_ld, 0, reward_dev, done_dev, env->n, obs_aos_dev, auto_reset, 1,
                     static_cast<cudaStream_t>(stream));
}

extern "C" RMV_API int rmv_step_rows(rmv_env* env, const float* rows_in_dev, const float* act_dev, int act_layout, int64_t act_ld,
                                     float* reward_dev, uint8_t* done_dev, float* rows_out_dev, int auto_reset, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->kind != RMV_REINMAV, "reinmav-v0 has no row-state step");
  RMV_REQUIRE(rows_in_dev != nullptr && rows_out_dev != nullptr && act_dev != nullptr && reward_dev != nullptr && done_dev != nullptr,
              "NULL argument");
  RMV_REQUIRE(act_layout == RMV_ACT_AOS || act_layout == RMV_ACT_SOA, "bad act_layout %d", act_layout);
  if (act_layout == RMV_ACT_SOA) RMV_REQUIRE(act_ld >= env->n, "act_ld (%lld) < n_envs", (long long)act_ld);
  RMV_REQUIRE(aligned(rows_in_dev, 4) && aligned(rows_out_dev, 4) && aligned(reward_dev, 4), "float buffers must be 4-byte aligned");
  RMV_REQUIRE(act_layout == RMV_ACT_SOA || aligned(act_dev, env->A * 4), "row-per-env actions must be aligned to one row (%d bytes)", env->A * 4);
  RMV_ON_DEVICE(env->device);
  rmv_env* e = env;
  StepArgs a{};
  a.state = e->state; a.episode = e->episode;
  a.act = act_dev; a.reward = reward_dev; a.done = done_dev; a.obs = rows_out_dev; a.rows_in = rows_in_dev;
  a.ep_return = e->ep_return; a.ep_start = e->ep_start; a.last_return = e->last_return; a.last_length = e->last_length;
  a.cum_return = e->cum_return; a.cum_length = e->cum_length; a.fin_count = e->fin_count;
  a.n = e->n; a.act_ld = act_ld; a.out_ld = e->n; a.act_step_stride = 0;
  a.env_id_offset = e->env_id_offset; a.seed = e->seed; a.step_index = e->step_index;
  philox_key_schedule(e->seed, a.ks);
  a.n_steps = 1; a.act_soa = (act_layout == RMV_ACT_SOA); a.auto_reset = auto_reset ? 1 : 0;
  a.P = e->dp;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t err;
  switch (e->kind) {
    case RMV_QUAD2D: err = launch_step_rows_k<QUAD2D>(e, a, st); break;
    case RMV_QUAD2D_SLUNG: err = launch_step_rows_k<QUAD2D_SLUNG>(e, a, st); break;
    case RMV_QUAD3D: err = launch_step_rows_k<QUAD3D>(e, a, st); break;
    default: err = launch_step_rows_k<QUAD3D_SLUNG>(e, a, st); break;
  }
  if (err != cudaSuccess) return fail(RMV_ECUDA, "row-state step kernel launch failed: %s", cudaGetErrorString(err));
  e->step_index += 1u;
  return RMV_OK;
}

extern "C" RMV_API int rmv_step_many(rmv_env* env, int n_steps, const float* act_dev, int act_layout, int64_t act_ld,
                             float* reward_dev, uint8_t* done_dev, int64_t out_ld, int fused, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(n_steps > 0, "n_steps must be positive");
  if (reward_dev || done_dev) RMV_REQUIRE(out_ld >= env->n, "out_ld (%lld) < n_envs", (long long)out_ld);
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (env->kind == RMV_REINMAV)
    return launch_step(env, nullptr, RMV_ACT_AOS, 0, 0, reward_dev, done_dev, out_ld, nullptr, 0, n_steps, st);
  const int64_t stride = (act_layout == RMV_ACT_SOA) ? (int64_t)env->A * act_ld : env->n * env->A;
  RMV_REQUIRE(fused >= 0 && fused <= 2, "mode must be RMV_MANY_STREAM (0), RMV_MANY_ROLLOUT (1) or RMV_MANY_LAUNCHES (2)");
  if (fused == RMV_MANY_ROLLOUT)
    return launch_step(env, act_dev, act_layout, act_ld, stride, reward_dev, done_dev, out_ld, nullptr, 1, n_steps, st);
  if (!reward_dev || !done_dev) {                 // the step kernels always write both: park a missing one in scratch
    int rc = ensure_staging(env);
    if (rc) return rc;
  }
  float* const rew0 = reward_dev ? reward_dev : env->h_rew;
  uint8_t* const done0 = done_dev ? done_dev : env->h_done;
  const int64_t out_stride = reward_dev ? out_ld : 0, done_stride = done_dev ? out_ld : 0;
  if (fused == RMV_MANY_STREAM && g_use_stream && n_steps >= 2 && (g_force_stream || stream_pays_for(env))) {
    // one stride for both outputs: if only one of them is parked, the other one's scratch must be T rows as well
    if ((reward_dev == nullptr) != (done_dev == nullptr))
      return fail(RMV_EINVAL, "rmv_step_many(RMV_MANY_STREAM): pass both reward_dev and done_dev, or neither");
    return launch_step(env, act_dev, act_layout, act_ld, stride, rew0, done0, out_stride, nullptr, 1, n_steps, st, 0, -1, true, true);
  }
  auto launch_all = [&](cudaStream_t s, bool advance) -> int {
    for (int t = 0; t < n_steps; ++t) {
      int rc = launch_step(env, act_dev + (int64_t)t * stride, act_layout, act_ld, 0, rew0 + (int64_t)t * out_stride,
                           done0 + (int64_t)t * done_stride, out_ld, nullptr, 1, 1, s, 0, -1, advance);
      if (rc) return rc;
    }
    return RMV_OK;
  };

  // CUDA-graph replay of a repeated call (launch-bound at small batches: 3-12 % at N <= 2^18, 2.5 % at N = 2^20,
  // profiles/r1b_probe_cuda_graph.txt).  Not with episode tracking (the step counter is a launch argument), not while
  // the caller's stream is itself being captured (the plain launches then become part of the caller's graph).
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool graph_ok = g_use_graph && !env->graphs_off && n_steps >= 4 && env->ep_return == nullptr &&
                        cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
  if (!graph_ok) { cudaGetLastError(); return launch_all(st, true); }

  auto same = [&](const rmv_env::StepGraph& g) {
    return g.n_steps == n_steps && g.layout == act_layout && g.act == act_dev && g.rew == rew0 && g.done == done0 &&
           g.act_ld == act_ld && g.out_ld == out_ld && g.epoch == env->config_epoch;
  };
  for (auto& g : env->graphs) {
    if (g.exec && same(g)) {
      g.stamp = ++env->graph_clock;
      RMV_CUDA(cudaGraphLaunch(g.exec, st));
      env->step_index += (uint32_t)n_steps;
      return RMV_OK;
    }
  }
  rmv_env::StepGraph* known = nullptr;
  for (auto& g : env->seen) if (same(g)) known = &g;
  if (!known) {                                    // first sighting: remember it (LRU of 4 shapes), launch plainly
    rmv_env::StepGraph* slot = &env->seen[0];
    for (auto& g : env->seen) if (g.stamp < slot->stamp) slot = &g;
    slot->n_steps = n_steps; slot->layout = act_layout; slot->act = act_dev; slot->rew = rew0; slot->done = done0;
    slot->act_ld = act_ld; slot->out_ld = out_ld; slot->epoch = env->config_epoch; slot->stamp = ++env->graph_clock;
    return launch_all(st, true);
  }
  // second sighting: capture on an internal stream (the caller's may be the legacy default stream)
  if (!env->capture_stream && cudaStreamCreateWithFlags(&env->capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError(); env->graphs_off = 1; return launch_all(st, true);
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = cudaStreamBeginCapture(env->capture_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    const int rc = launch_all(env->capture_stream, false);
    const cudaError_t ce = cudaStreamEndCapture(env->capture_stream, &graph);
    ok = rc == RMV_OK && ce == cudaSuccess && graph != nullptr;
  }
  if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError(); env->graphs_off = 1;
    return launch_all(st, true);
  }
  rmv_env::StepGraph* slot = nullptr;              // an empty slot, else the least recently used one
  for (auto& g : env->graphs) {
    if (!g.exec) { slot = &g; break; }
    if (!slot || g.stamp < slot->stamp) slot = &g;
  }
  if (slot->exec) cudaGraphExecDestroy(slot->exec);
  *slot = *known;
  slot->exec = exec;
  slot->stamp = ++env->graph_clock;
  RMV_CUDA(cudaGraphLaunch(exec, st));
  env->step_index += (uint32_t)n_steps;
  return RMV_OK;
}

// Host-buffer step.  Large batches are cut into chunks of whole state tiles that flow through three
// internal streams: chunk i's host->device action copy, its step kernel and its device->host result copies
// are ordered inside one stream, different chunks overlap — so the two PCIe directions and the kernel run
// concurrently and the call costs ~max(H2D, D2H) instead of their sum.
static int ensure_pipe(rmv_env* e) {
  for (int i = 0; i < 3; ++i) {
    if (!e->pipe[i]) RMV_CUDA(cudaStreamCreateWithFlags(&e->pipe[i], cudaStreamNonBlocking));
    if (!e->pipe_done[i]) RMV_CUDA(cudaEventCreateWithFlags(&e->pipe_done[i], cudaEventDisableTiming));
    if (!e->pipe_kernel[i]) RMV_CUDA(cudaEventCreateWithFlags(&e->pipe_kernel[i], cudaEventDisableTiming));
  }
  if (!e->pipe_small) RMV_CUDA(cudaStreamCreateWithFlags(&e->pipe_small, cudaStreamNonBlocking));
  if (!e->pipe_small_done) RMV_CUDA(cudaEventCreateWithFlags(&e->pipe_small_done, cudaEventDisableTiming));
  if (!e->pipe_start) RMV_CUDA(cudaEventCreateWithFlags(&e->pipe_start, cudaEventDisableTiming));
  return RMV_OK;
}

extern "C" RMV_API int rmv_step_host(rmv_env* env, const float* act_host, float* reward_host, uint8_t* done_host,
                             float* obs_host, int auto_reset, void* stream) {
  RMV_REQUIRE(env != nullptr && (act_host != nullptr || env->A == 0), "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_staging(env);
  if (rc) return rc;
  const int64_t n = env->n;
  const int S = env->S, A = env->A;
  // Chunking: 8 equal chunks for large batches.  (Chunks that grow geometrically — n/32, n/32, n/16, ... n/2, to start the
  // result copies earlier — measured no better: 1.02 against 1.00 ms per step at N = 2^20, profiles/README.md r2.)
  int chunks = env->host_chunks > 0 ? env->host_chunks : (n >= (1 << 16) ? 8 : 1);
  if (env->A == 0) chunks = 1;
  int64_t per = ((n + chunks - 1) / chunks + 1023) / 1024 * 1024;      // whole tiles, aligned for every vector width
  if (per >= n) chunks = 1;
  // observation rows leave from the staging copy the kernel writes, or — on a row-state handle — from the state itself
  float* const obs_dev = env->rows ? nullptr : (obs_host ? env->h_obs : nullptr);
  const float* const obs_src = env->rows ? env->state : env->h_obs;
  if (chunks == 1) {
    if (A > 0) RMV_CUDA(cudaMemcpyAsync(env->h_act, act_host, (size_t)n * A * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = launch_step(env, env->h_act, RMV_ACT_AOS, 0, 0, env->h_rew, env->h_done, n, obs_dev, auto_reset, 1, st);
    if (rc) return rc;
    if (reward_host) RMV_CUDA(cudaMemcpyAsync(reward_host, env->h_rew, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (done_host) RMV_CUDA(cudaMemcpyAsync(done_host, env->h_done, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (obs_host) RMV_CUDA(cudaMemcpyAsync(obs_host, obs_src, (size_t)n * S * sizeof(float), cudaMemcpyDeviceToHost, st));
    RMV_CUDA(cudaStreamSynchronize(st));
    return RMV_OK;
  }
  rc = ensure_pipe(env);
  if (rc) return rc;
  RMV_CUDA(cudaEventRecord(env->pipe_start, st));          // everything queued on the caller's stream comes first
  for (int i = 0; i < 3; ++i) RMV_CUDA(cudaStreamWaitEvent(env->pipe[i], env->pipe_start, 0));
  RMV_CUDA(cudaStreamWaitEvent(env->pipe_small, env->pipe_start, 0));
  // The observation rows are the bulk (4S of the 4S + 5 bytes per env going back): they leave chunk by chunk behind their
  // kernel.  Rewards and done flags leave ONCE per call on a stream of their own, as soon as the last kernel has run
  // (the kernels are done long before the observation copies are): 2 small DMA transfers instead of 2 per chunk, and
  // their set-up overlaps the bulk copies instead of sitting between them.
  const bool batch_small = g_host_batch_small != 0 && obs_host != nullptr;
  int k = 0;
  for (int64_t c0 = 0; c0 < n; c0 += per, ++k) {
    const int64_t cn = (n - c0 < per) ? n - c0 : per;
    cudaStream_t ps = env->pipe[k % 3];
    RMV_CUDA(cudaMemcpyAsync(env->h_act + c0 * A, act_host + c0 * A, (size_t)cn * A * sizeof(float), cudaMemcpyHostToDevice, ps));
    rc = launch_step(env, env->h_act + c0 * A, RMV_ACT_AOS, 0, 0, env->h_rew + c0, env->h_done + c0, cn,
                     obs_dev ? obs_dev + c0 * S : nullptr, auto_reset, 1, ps, c0, cn, false);
    if (rc) return rc;
    if (batch_small) RMV_CUDA(cudaEventRecord(env->pipe_kernel[k % 3], ps));
    if (obs_host) RMV_CUDA(cudaMemcpyAsync(obs_host + c0 * S, obs_src + c0 * S, (size_t)cn * S * sizeof(float), cudaMemcpyDeviceToHost, ps));
    if (!batch_small) {
      if (reward_host) RMV_CUDA(cudaMemcpyAsync(reward_host + c0, env->h_rew + c0, (size_t)cn * sizeof(float), cudaMemcpyDeviceToHost, ps));
      if (done_host) RMV_CUDA(cudaMemcpyAsync(done_host + c0, env->h_done + c0, (size_t)cn, cudaMemcpyDeviceToHost, ps));
    }
  }
  env->step_index += 1u;
  if (batch_small) {
    for (int i = 0; i < 3 && i < k; ++i) RMV_CUDA(cudaStreamWaitEvent(env->pipe_small, env->pipe_kernel[i], 0));
    if (reward_host) RMV_CUDA(cudaMemcpyAsync(reward_host, env->h_rew, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, env->pipe_small));
    if (done_host) RMV_CUDA(cudaMemcpyAsync(done_host, env->h_done, (size_t)n, cudaMemcpyDeviceToHost, env->pipe_small));
    RMV_CUDA(cudaEventRecord(env->pipe_small_done, env->pipe_small));
    RMV_CUDA(cudaStreamWaitEvent(st, env->pipe_small_done, 0));
  }
  for (int i = 0; i < 3; ++i) {
    RMV_CUDA(cudaEventRecord(env->pipe_done[i], env->pipe[i]));
    RMV_CUDA(cudaStreamWaitEvent(st, env->pipe_done[i], 0));   // later work on the caller's stream sees the new state
  }
  for (int i = 0; i < 3; ++i) RMV_CUDA(cudaStreamSynchronize(env->pipe[i]));
  if (batch_small) RMV_CUDA(cudaStreamSynchronize(env->pipe_small));
  return RMV_OK;
}

extern "C" RMV_API int rmv_reset_host(rmv_env* env, float* obs_host, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_staging(env);
  if (rc) return rc;
  rc = rmv_reset(env, nullptr, (obs_host && !env->rows) ? env->h_obs : nullptr, stream);
  if (rc) return rc;
  if (obs_host) RMV_CUDA(cudaMemcpyAsync(obs_host, env->rows ? env->state : env->h_obs, (size_t)env->n * env->S * sizeof(float), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// controllers
// ------------------------------------------------------------------------------------------
// self.ref_pos / self.ref_vel defaults of the reference classes (quadrotor3d.py:55-56 and twins) + the constants control() reads
static CtrlParams ctrl_params_of(const rmv_env* env, const float* ref_pos, const float* ref_vel) {
  static const float kRef[RMV_NUM_KINDS][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 2.0f}, {0, 0, 1.0f}, {0, 0, 0}};
  const int D = (env->kind == RMV_QUAD2D || env->kind == RMV_QUAD2D_SLUNG) ? 2 : 3;
  CtrlParams C;
  for (int i = 0; i < 3; ++i) {
    C.ref_pos[i] = (i < D) ? (ref_pos ? ref_pos[i] : kRef[env->kind][i]) : 0.0f;
    C.ref_vel[i] = (i < D && ref_vel) ? ref_vel[i] : 0.0f;
  }
  C.mass = env->params.mass; C.gravity = env->params.gravity;
  return C;
}

extern "C" RMV_API int rmv_control_step(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_out_dev,
                                        float* reward_dev, uint8_t* done_dev, float* obs_aos_dev, int auto_reset, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->kind != RMV_REINMAV, "reinmav-v0 runs its controller inside step() already");
  RMV_REQUIRE(reward_dev != nullptr && done_dev != nullptr, "reward_dev and done_dev must not be NULL");
  RMV_ON_DEVICE(env->device);
  const CtrlParams C = ctrl_params_of(env, ref_pos, ref_vel);
  return launch_step(env, nullptr, RMV_ACT_AOS, 0, 0, reward_dev, done_dev, env->n, obs_aos_dev, auto_reset, 1,
                     static_cast<cudaStream_t>(stream), 0, -1, true, false, &C, act_out_dev);
}

extern "C" RMV_API int rmv_control_step_many(rmv_env* env, int n_steps, const float* ref_pos, const float* ref_vel,
                                             float* act_out_dev, float* reward_dev, uint8_t* done_dev, int64_t out_ld, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->kind != RMV_REINMAV, "reinmav-v0 runs its controller inside step() already");
  RMV_REQUIRE(n_steps > 0, "n_steps must be positive");
  RMV_REQUIRE((reward_dev == nullptr) == (done_dev == nullptr), "pass both reward_dev and done_dev, or neither");
  if (reward_dev) RMV_REQUIRE(out_ld >= env->n, "out_ld (%lld) < n_envs", (long long)out_ld);
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const CtrlParams C = ctrl_params_of(env, ref_pos, ref_vel);
  if (!reward_dev) { int rc = ensure_staging(env); if (rc) return rc; }
  float* const rew0 = reward_dev ? reward_dev : env->h_rew;
  uint8_t* const done0 = done_dev ? done_dev : env->h_done;
  const int64_t stride = reward_dev ? out_ld : 0;
  if (g_use_stream && n_steps >= 2 && (g_force_stream || stream_pays_for(env)))   // ONE launch: the closed loop runs inside the step stream
    return launch_step(env, nullptr, RMV_ACT_AOS, 0, 0, rew0, done0, stride, nullptr, 1, n_steps, st, 0, -1, true, true, &C, act_out_dev);
  for (int t = 0; t < n_steps; ++t) {
    int rc = launch_step(env, nullptr, RMV_ACT_AOS, 0, 0, rew0 + (int64_t)t * stride, done0 + (int64_t)t * stride, env->n, nullptr, 1, 1,
                         st, 0, -1, true, false, &C, act_out_dev ? act_out_dev + (int64_t)t * env->n * env->A : nullptr);
    if (rc) return rc;
  }
  return RMV_OK;
}

extern "C" RMV_API int rmv_control(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_dev,
                                   int act_layout, int64_t act_ld, void* stream) {
  RMV_REQUIRE(env != nullptr && act_dev != nullptr, "NULL argument");
  RMV_REQUIRE(env->kind != RMV_REINMAV, "reinmav-v0 has no action input: its controller runs inside step()");
  RMV_REQUIRE(act_layout == RMV_ACT_AOS || act_layout == RMV_ACT_SOA, "bad act_layout %d", act_layout);
  if (act_layout == RMV_ACT_SOA) RMV_REQUIRE(act_ld >= env->n, "act_ld (%lld) < n_envs", (long long)act_ld);
  RMV_ON_DEVICE(env->device);
  ControlArgs a;
  a.state = env->state; a.act = act_dev; a.n = env->n; a.act_ld = act_ld; a.act_soa = (act_layout == RMV_ACT_SOA);
  a.rows = env->rows ? 1 : 0;
  a.C = ctrl_params_of(env, ref_pos, ref_vel);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = blocks_for(env->n, 128);
  switch (env->kind) {
    case RMV_QUAD2D: k_control<QUAD2D><<<grid, 128, 0, st>>>(a); break;
    case RMV_QUAD2D_SLUNG: k_control<QUAD2D_SLUNG><<<grid, 128, 0, st>>>(a); break;
    case RMV_QUAD3D: k_control<QUAD3D><<<grid, 128, 0, st>>>(a); break;
    default: k_control<QUAD3D_SLUNG><<<grid, 128, 0, st>>>(a); break;
  }
  RMV_CUDA(cudaGetLastError());
  return RMV_OK;
}

extern "C" RMV_API int rmv_control_host(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_host,
                                        void* stream) {
  RMV_REQUIRE(env != nullptr && act_host != nullptr, "NULL argument");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_staging(env);
  if (rc) return rc;
  rc = rmv_control(env, ref_pos, ref_vel, env->h_act, RMV_ACT_AOS, 0, stream);
  if (rc) return rc;
  RMV_CUDA(cudaMemcpyAsync(act_host, env->h_act, (size_t)env->n * env->A * sizeof(float), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic actions, statistics, pinned memory
// ------------------------------------------------------------------------------------------
extern "C" RMV_API int rmv_fill_uniform_actions(float* act_dev, int64_t n_envs, int action_dim, int act_layout,
                                        int64_t act_ld, float low, float high, uint64_t seed,
                                        int64_t env_id_offset, uint32_t step_index, int device, void* stream) {
  RMV_REQUIRE(act_dev != nullptr && n_envs > 0, "bad buffer");
  RMV_REQUIRE(action_dim > 0 && action_dim <= 16, "bad action_dim");
  RMV_REQUIRE(act_layout == RMV_ACT_AOS || act_layout == RMV_ACT_SOA, "bad act_layout %d", act_layout);
  if (act_layout == RMV_ACT_SOA) RMV_REQUIRE(act_ld >= n_envs, "act_ld < n_envs");
  RMV_ON_DEVICE(device);
  const float width = high - low;
  k_fill_uniform<<<blocks_for(n_envs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      act_dev, n_envs, action_dim, act_layout == RMV_ACT_SOA, act_ld, low, width, seed, env_id_offset, step_index);
  RMV_CUDA(cudaGetLastError());
  return RMV_OK;
}

extern "C" RMV_API int rmv_episode_stats(rmv_env* env, double out[4], void* stream) {
  RMV_REQUIRE(env != nullptr && out != nullptr, "NULL argument");
  RMV_REQUIRE(env->flags & RMV_TRACK_EPISODES, "handle was created without RMV_TRACK_EPISODES");
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RMV_CUDA(cudaMemsetAsync(env->stats_dev, 0, 4 * sizeof(double), st));
  k_episode_stats<<<296, 256, 0, st>>>(env->fin_count, env->cum_return, env->cum_length, env->n, env->stats_dev);
  RMV_CUDA(cudaGetLastError());
  RMV_CUDA(cudaMemcpyAsync(out, env->stats_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaStreamSynchronize(st));
  out[3] = (double)env->step_index * (double)env->n;
  return RMV_OK;
}

extern "C" RMV_API int rmv_gather_episodes(rmv_env* env, const int64_t* idx_host, int64_t k, float* return_host,
                                           uint32_t* length_host, void* stream) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->flags & RMV_TRACK_EPISODES, "handle was created without RMV_TRACK_EPISODES");
  RMV_REQUIRE(k >= 0 && k <= env->n, "k out of range");
  if (k == 0) return RMV_OK;
  RMV_REQUIRE(idx_host != nullptr && return_host != nullptr && length_host != nullptr, "NULL argument");
  for (int64_t i = 0; i < k; ++i) RMV_REQUIRE(idx_host[i] >= 0 && idx_host[i] < env->n, "env index %lld out of range", (long long)idx_host[i]);
  RMV_ON_DEVICE(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (env->gather_cap < k) {
    cudaFree(env->gather_idx); cudaFree(env->gather_ret); cudaFree(env->gather_len);
    env->gather_idx = nullptr; env->gather_ret = nullptr; env->gather_len = nullptr; env->gather_cap = 0;
    int64_t cap = 1024;
    while (cap < k) cap *= 2;
    RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&env->gather_idx), (size_t)cap * sizeof(int64_t)));
    RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&env->gather_ret), (size_t)cap * sizeof(float)));
    RMV_CUDA(cudaMalloc(reinterpret_cast<void**>(&env->gather_len), (size_t)cap * sizeof(uint32_t)));
    env->gather_cap = cap;
  }
  RMV_CUDA(cudaMemcpyAsync(env->gather_idx, idx_host, (size_t)k * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  k_gather_episodes<<<blocks_for(k, 256), 256, 0, st>>>(env->last_return, env->last_length, env->gather_idx, k,
                                                        env->gather_ret, env->gather_len);
  RMV_CUDA(cudaGetLastError());
  RMV_CUDA(cudaMemcpyAsync(return_host, env->gather_ret, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaMemcpyAsync(length_host, env->gather_len, (size_t)k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaStreamSynchronize(st));
  return RMV_OK;
}

extern "C" RMV_API int rmv_episode_ptr(rmv_env* env, float** last_return_dev, uint32_t** last_length_dev) {
  RMV_REQUIRE(env != nullptr, "env is NULL");
  RMV_REQUIRE(env->flags & RMV_TRACK_EPISODES, "handle was created without RMV_TRACK_EPISODES");
  if (last_return_dev) *last_return_dev = env->last_return;
  if (last_length_dev) *last_length_dev = env->last_length;
  return RMV_OK;
}

// ------------------------------------------------------------------------------------------
// VecNormalize on the device (rmv_norm.cuh)
// ------------------------------------------------------------------------------------------
struct rmv_norm {
  int device = 0, S = 0;
  int64_t n = 0;
  int norm_obs = 1, norm_ret = 1;
  float clip_obs = 10.0f, clip_rew = 10.0f;
  double gamma = 0.99, epsilon = 1e-8;
  double* ret = nullptr;       // [n]
  double* stats = nullptr;     // [2][kNormStats]
  double* sums = nullptr;      // [2][kNormSumsAll]
  uint64_t step = 0;           // parity selects the current halves
};

static void norm_host_defaults(double* st) {
  for (int c = 0; c <= kNormMaxDim; ++c) { st[c] = 0.0; st[(kNormMaxDim + 1) + c] = 1.0; }   // RunningMeanStd: mean 0, var 1
  st[2 * (kNormMaxDim + 1)] = 1e-4; st[2 * (kNormMaxDim + 1) + 1] = 1e-4;                      // count = epsilon = 1e-4
}

extern "C" RMV_API int rmv_norm_create(rmv_norm** out, int64_t n_envs, int obs_dim, int device, int norm_obs, int norm_ret,
                                       float clip_obs, float clip_rew, double gamma, double epsilon) {
  RMV_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  RMV_REQUIRE(n_envs > 0 && obs_dim > 0 && obs_dim <= kNormMaxDim, "n_envs must be positive and obs_dim in [1, %d]", kNormMaxDim);
  RMV_ON_DEVICE(device);
  rmv_norm* nm = new (std::nothrow) rmv_norm();
  if (!nm) return fail(RMV_ENOMEM, "host allocation failed");
  nm->device = device; nm->S = obs_dim; nm->n = n_envs; nm->norm_obs = norm_obs ? 1 : 0; nm->norm_ret = norm_ret ? 1 : 0;
  nm->clip_obs = clip_obs; nm->clip_rew = clip_rew; nm->gamma = gamma; nm->epsilon = epsilon;
  cudaError_t err = alloc_zero(&nm->ret, (size_t)n_envs);
  if (err == cudaSuccess) err = alloc_zero(&nm->stats, (size_t)2 * kNormStats);
  if (err == cudaSuccess) err = alloc_zero(&nm->sums, (size_t)2 * kNormSumsAll);
  if (err == cudaSuccess) {
    double st[kNormStats];
    norm_host_defaults(st);
    err = cudaMemcpy(nm->stats, st, sizeof(st), cudaMemcpyHostToDevice);
  }
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    cudaFree(nm->ret); cudaFree(nm->stats); cudaFree(nm->sums);
    delete nm;
    return fail(err == cudaErrorMemoryAllocation ? RMV_ENOMEM : RMV_ECUDA, "rmv_norm_create: %s", cudaGetErrorString(err));
  }
  *out = nm;
  return RMV_OK;
}

extern "C" RMV_API int rmv_norm_destroy(rmv_norm* nm) {
  if (!nm) return RMV_OK;
  {
    RMV_ON_DEVICE(nm->device);
    cudaDeviceSynchronize();
    cudaFree(nm->ret); cudaFree(nm->stats); cudaFree(nm->sums);
  }
  delete nm;
  return RMV_OK;
}

static int norm_launch(rmv_norm* nm, const float* obs, const float* rew, const uint8_t* done, float* obs_out, float* rew_out,
                       int update, cudaStream_t st) {
  NormArgs a;
  const int cur = (int)(nm->step & 1u);
  a.obs = obs; a.reward = rew; a.done = done; a.obs_out = obs_out; a.reward_out = rew_out; a.ret = nm->ret;
  a.stats_cur = nm->stats + cur * kNormStats; a.stats_next = nm->stats + (cur ^ 1) * kNormStats;
  a.sums_cur = nm->sums + cur * kNormSumsAll; a.sums_next = nm->sums + (cur ^ 1) * kNormSumsAll;
  a.n = nm->n; a.S = nm->S; a.norm_obs = nm->norm_obs; a.norm_ret = nm->norm_ret; a.update = update ? 1 : 0;
  a.clip_obs = nm->clip_obs; a.clip_rew = nm->clip_rew; a.gamma = nm->gamma; a.epsilon = nm->epsilon;
  // flat float4 walk when both matrices allow it; ONE wave of CTAs, a multiple of S of them so that a thread's lanes keep
  // their components
  const bool vec4 = aligned(obs, 16) && aligned(obs_out, 16);
  static int wave_of[kMaxDevices][2] = {};
  int& wave = wave_of[nm->device][vec4 ? 1 : 0];
  if (wave == 0) {
    int sms = 1, r0 = 1, r1 = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, nm->device);
    if (vec4) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r0, k_norm_moments<4>, kNormBlock, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r1, k_norm_apply<4>, kNormBlock, 0);
    } else {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r0, k_norm_moments<1>, kNormBlock, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r1, k_norm_apply<1>, kNormBlock, 0);
    }
    wave = std::max(1, sms) * std::max(1, std::min(r0, r1));
  }
  const int64_t work = vec4 ? (nm->n * nm->S + 3) / 4 : nm->n * nm->S;
  int64_t want = (std::max<int64_t>(work, nm->n) + kNormBlock - 1) / kNormBlock;
  unsigned grid;
  if (want >= wave) grid = (unsigned)std::max<int64_t>(nm->S, (int64_t)wave / nm->S * nm->S);
  else grid = (unsigned)((want + nm->S - 1) / nm->S * nm->S);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kNormBlock); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  if (vec4) { RMV_CUDA(cudaLaunchKernelEx(&cfg, k_norm_moments<4>, a)); RMV_CUDA(cudaLaunchKernelEx(&cfg, k_norm_apply<4>, a)); }
  else { RMV_CUDA(cudaLaunchKernelEx(&cfg, k_norm_moments<1>, a)); RMV_CUDA(cudaLaunchKernelEx(&cfg, k_norm_apply<1>, a)); }
  RMV_CUDA(cudaGetLastError());
  nm->step += 1;
  return RMV_OK;
}

extern "C" RMV_API int rmv_norm_step(rmv_norm* nm, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev,
                                     float* obs_out_dev, float* reward_out_dev, int update, void* stream) {
  RMV_REQUIRE(nm != nullptr && reward_dev != nullptr && done_dev != nullptr && reward_out_dev != nullptr, "NULL argument");
  RMV_REQUIRE((obs_dev == nullptr) == (obs_out_dev == nullptr), "pass both obs_dev and obs_out_dev, or neither");
  RMV_ON_DEVICE(nm->device);
  return norm_launch(nm, obs_dev, reward_dev, done_dev, obs_out_dev, reward_out_dev, update, static_cast<cudaStream_t>(stream));
}

extern "C" RMV_API int rmv_norm_reset(rmv_norm* nm, const float* obs_dev, float* obs_out_dev, int update, void* stream) {
  RMV_REQUIRE(nm != nullptr && obs_dev != nullptr && obs_out_dev != nullptr, "NULL argument");
  RMV_ON_DEVICE(nm->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RMV_CUDA(cudaMemsetAsync(nm->ret, 0, (size_t)nm->n * sizeof(double), st));       // self.ret = np.zeros(self.num_envs)
  return norm_launch(nm, obs_dev, nullptr, nullptr, obs_out_dev, nullptr, update, st);
}

extern "C" RMV_API int rmv_norm_get(rmv_norm* nm, double* ob_mean, double* ob_var, double* ob_count, double* ret_mean,
                                    double* ret_var, double* ret_count, void* stream) {
  RMV_REQUIRE(nm != nullptr, "nm is NULL");
  RMV_ON_DEVICE(nm->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double h[kNormStats];
  RMV_CUDA(cudaMemcpyAsync(h, nm->stats + (nm->step & 1u) * kNormStats, sizeof(h), cudaMemcpyDeviceToHost, st));
  RMV_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < nm->S; ++c) {
    if (ob_mean) ob_mean[c] = h[c];
    if (ob_var) ob_var[c] = h[(kNormMaxDim + 1) + c];
  }
  if (ob_count) *ob_count = h[2 * (kNormMaxDim + 1)];
  if (ret_mean) *ret_mean = h[nm->S];
  if (ret_var) *ret_var = h[(kNormMaxDim + 1) + nm->S];
  if (ret_count) *ret_count = h[2 * (kNormMaxDim + 1) + 1];
  return RMV_OK;
}

extern "C" RMV_API int rmv_norm_set(rmv_norm* nm, const double* ob_mean, const double* ob_var, double ob_count, double ret_mean,
                                    double ret_var, double ret_count, void* stream) {
  RMV_REQUIRE(nm != nullptr && ob_mean != nullptr && ob_var != nullptr, "NULL argument");
  RMV_ON_DEVICE(nm->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double h[kNormStats];
  norm_host_defaults(h);
  for (int c = 0; c < nm->S; ++c) { h[c] = ob_mean[c]; h[(kNormMaxDim + 1) + c] = ob_var[c]; }
  h[nm->S] = ret_mean; h[(kNormMaxDim + 1) + nm->S] = ret_var;
  h[2 * (kNormMaxDim + 1)] = ob_count; h[2 * (kNormMaxDim + 1) + 1] = ret_count;
  RMV_CUDA(cudaStreamSynchronize(st));
  RMV_CUDA(cudaMemcpy(nm->stats + (nm->step & 1u) * kNormStats, h, sizeof(h), cudaMemcpyHostToDevice));
  return RMV_OK;
}

extern "C" RMV_API int rmv_host_alloc(void** out, size_t bytes) {
  RMV_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  RMV_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return RMV_OK;
}

extern "C" RMV_API int rmv_host_free(void* p) {
  if (p) RMV_CUDA(cudaFreeHost(p));
  return RMV_OK;
}
