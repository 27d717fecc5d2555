/* rmv.h — C-ABI of librmv.so: the batched native-quadrotor step()/reset() path on B200.
 *
 * This is the drop-in boundary for ONE path of ethz-asl/reinmav-gym: the native envs'
 * step()/reset() (dynamics + reward + done + reset-on-done).  The reference has no FFI of
 * its own (it is pure Python), so each entry point cites the Python method it replaces;
 * paths are relative to /root/reference/gym_reinmav/.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RMV_E* code on failure; it never
 *     throws or aborts.  rmv_last_error() returns a thread-local message for the last failure.
 *   - rmv_env is an opaque handle that owns the struct-of-arrays state of n envs on ONE device.
 *     A handle is not thread-safe; distinct handles are independent (one per GPU).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All device work
 *     is enqueued on it and nothing synchronises the host, except the *_host entry points and
 *     calls with host-side destination/source buffers, which say so.
 *   - pointers named *_dev are device pointers on the handle's device; the caller keeps them
 *     alive until the stream has run past the call.
 *   - state layout in HBM: float32 [n_pad/128][S][128] — tiled struct-of-arrays: env e is lane
 *     e%128 of tile e/128, a tile holds the S components of 128 consecutive envs contiguously
 *     (component-major inside the tile); n_pad = rmv_padded_ld(n).
 *     Component order is the reference's tuple order:
 *       quadrotor2d            [x y theta vx vy]                          envs/native/quadrotor2d.py:93
 *       quadrotor2d-slungload  [x y theta vx vy lx ly lvx lvy]            envs/native/quadrotor2d_slungload.py:134
 *       quadrotor3d            [px py pz qw qx qy qz vx vy vz]            envs/native/quadrotor3d.py:104
 *       quadrotor3d-slungload  [.. quadrotor3d .. lx ly lz lvx lvy lvz]   envs/native/quadrotor3d_slungload.py:147
 *   - there is no CPU implementation behind this ABI: without a CUDA device every call that
 *     needs one fails with RMV_ECUDA.
 */
#ifndef RMV_H_
#define RMV_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RMV_API __attribute__((visibility("default")))
#else
#define RMV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RMV_ABI_VERSION 2

/* env kinds = the four registered native ids, gym_reinmav/__init__.py:8-26 */
enum {
  RMV_QUAD2D = 0,        /* quadrotor2d-v0            envs/native/quadrotor2d.py            */
  RMV_QUAD2D_SLUNG = 1,  /* quadrotor2d-slungload-v0  envs/native/quadrotor2d_slungload.py  */
  RMV_QUAD3D = 2,        /* quadrotor3d-v0            envs/native/quadrotor3d.py            */
  RMV_QUAD3D_SLUNG = 3,  /* quadrotor3d-slungload-v0  envs/native/quadrotor3d_slungload.py  */
  RMV_REINMAV = 4,       /* reinmav-v0                envs/native/reinmav_env.py: closed-loop
                            13-state rigid body [x y z vx vy vz qw qx qy qz p q r], NO action (action_dim 0,
                            act_dev may be NULL), reward 90 and done 1 every step, reset() keeps the state
                            (reinmav_env.py:99-126,348-351); rmv_params is not used by this kind */
  RMV_NUM_KINDS = 5
};

enum {
  RMV_OK = 0,
  RMV_EINVAL = -1,   /* bad argument                                   */
  RMV_ECUDA = -2,    /* CUDA runtime / driver error, or no device      */
  RMV_ENOMEM = -3,   /* host or device allocation failed               */
  RMV_EALIGN = -4    /* pointer / leading dimension not aligned enough */
};

/* action buffer layouts */
enum {
  RMV_ACT_AOS = 0,   /* [n][A]      — what a policy emits (row per env)                   */
  RMV_ACT_SOA = 1    /* [A][act_ld] — component-major, act_ld >= n                         */
};

/* rmv_create flags */
enum {
  RMV_TRACK_EPISODES = 1u, /* keep per-env episode return/length (baselines Monitor's info['episode']);
                              costs 8 extra bytes per env-step */
  RMV_STATE_ROWS = 2u      /* the handle keeps its state as float32 [n_pad][S] ROWS instead of tiles: the state buffer
                              IS the observation matrix a policy reads (obs == state in the reference,
                              quadrotor3d.py:104), rmv_step updates it in place through the row-state kernel and needs no
                              observation copy — 4(2S+A+1)+1 bytes per env-step where tiles + obs rows move
                              4(3S+A+1)+1.  rmv_state_ptr / rmv_step's obs_aos_dev == NULL then mean "read the rows
                              at rmv_state_ptr".  Same results bit for bit.  Not for RMV_REINMAV. */
};

/* Model constants.  The reference hard-codes them in each __init__ (quadrotor3d.py:45-56,
 * quadrotor3d_slungload.py:45-59, quadrotor2d.py:44-54, quadrotor2d_slungload.py:44-58);
 * rmv_default_params() returns exactly those. */
typedef struct rmv_params {
  float mass;               /* 1.0                                                            */
  float load_mass;          /* 0.1 (slung-load kinds)                                         */
  float dt;                 /* 0.01                                                           */
  float gravity;            /* 9.8, acts along -y (2D) / -z (3D)                               */
  float tether_len;         /* 1.5 (3D slung) | 0.5 (2D slung)                                 */
  float pos_done;           /* 3.0 | 2.0 (2D slung): done when the position norm exceeds it     */
  float vel_done;           /* 10.0 | 2.0 (2D, repair B of quadrotor2d.py:95-98)                */
  float thrust_scale;       /* 10.0 for quadrotor2d (quadrotor2d.py:75), else 1.0              */
  int32_t thrust_clamp;     /* 1 for quadrotor2d: thrust = max(thrust, 0) (quadrotor2d.py:76-77) */
  float first_done_reward;  /* 1.0  first done ever returned by an env (quadrotor3d.py:114-117)  */
  float later_done_reward;  /* 0.0  every later done (quadrotor3d.py:118-122)                   */
  float reward_scale;       /* 1.0  the launcher's --reward_scale (run.py:89 -> baselines RewardScaler, applied
                               OUTSIDE Monitor): the reward a step returns is reward * reward_scale, one float32
                               product; episode returns (info['episode']['r']) stay unscaled            */
} rmv_params;

typedef struct rmv_env rmv_env;

/* ---- queries that need no device -------------------------------------------------------- */
RMV_API int rmv_abi_version(void);
RMV_API const char* rmv_last_error(void);
/* state / action dimensions of a kind: 5/2, 9/2, 10/4, 16/4 (observation_space / action_space
 * shapes, e.g. quadrotor3d.py:70-71) */
RMV_API int rmv_kind_dims(int kind, int* state_dim, int* action_dim);
RMV_API int rmv_default_params(int kind, rmv_params* out);
/* padded env count used for n envs (n rounded up to a multiple of the 128-env state tile) */
RMV_API int64_t rmv_padded_ld(int64_t n_envs);
/* U(action_space) bounds of a kind (quadrotor2d.py:62, quadrotor3d.py:70, ...) */
RMV_API int rmv_action_bounds(int kind, float* low, float* high);

/* ---- lifetime --------------------------------------------------------------------------- */
/* Replaces constructing n env objects (Quadrotor3D.__init__, quadrotor3d.py:44-74) + seed()
 * (:77-79).  env_id_offset is the global index of env 0 of this shard: reset draws are keyed by
 * (seed, global env id, episode) so results do not depend on how envs are sharded over GPUs.
 * params == NULL -> rmv_default_params(kind).  State is zero until rmv_reset(). */
RMV_API int rmv_create(rmv_env** out, int kind, int64_t n_envs, int64_t env_id_offset, int device,
               uint64_t seed, const rmv_params* params, uint32_t flags);
RMV_API int rmv_destroy(rmv_env* env);
RMV_API int rmv_dims(const rmv_env* env, int* state_dim, int* action_dim, int64_t* n_envs, int64_t* ld);
RMV_API int rmv_get_params(const rmv_env* env, rmv_params* out);
/* seed(): re-key the reset stream (Quadrotor3D.seed, quadrotor3d.py:77-79).  State is untouched. */
RMV_API int rmv_set_seed(rmv_env* env, uint64_t seed);

/* reinmav-v0 only: the simulation clock `self.t` (reinmav_env.py:73,119), advanced by dt = 0.01 per step;
 * the trajectory generator and the substep grid depend on it. */
RMV_API int rmv_get_time(const rmv_env* env, double* t);
RMV_API int rmv_set_time(rmv_env* env, double t);

/* Launch tuning (benchmarks / sweeps): vec = envs per thread (0 = auto, 1, 2, 4; for the step stream of quadrotor3d
 * vec = 2 selects the two-envs-per-thread kernel on the packed float32 pipe, vec = 1 the one-env kernel),
 * block = threads per CTA of the one-env-per-thread step kernel (0 = auto, 128, 256, 512; vec > 1 always uses 128). */
RMV_API int rmv_set_tuning(rmv_env* env, int vec, int block);
/* the launch shape "auto" resolves to for a batch of n_envs of `kind` (device-free query): threads per CTA and whether the
 * episode word is fetched with the state — by kind and by whether one step's working set stays in L2 */
RMV_API int rmv_launch_shape(int kind, int64_t n_envs, int* block, int* prefetch);
/* step-kernel variant: use_tma != 0 selects the persistent TMA-bulk-copy / mbarrier-pipelined kernel
 * (V = 1, row-per-env actions) where it applies, 0 the register-path kernel; ctas_per_sm caps the
 * persistent grid (0 = as many as fit).  Results are identical. */
RMV_API int rmv_set_kernel(rmv_env* env, int use_tma, int ctas_per_sm);
/* one-env-per-thread step kernel: fetch every env's `episode` word together with its state (1) instead of only
 * on the lanes that finish (0); -1 = auto (by kind and working-set size).  Trades 4 B per env-step of read
 * traffic for one dependent memory round trip fewer in every CTA that holds a finished env.  Results are identical. */
RMV_API int rmv_set_prefetch(rmv_env* env, int mode);
/* Opt-in: keep the handle's state resident in L2 between steps with an access-policy window on the state buffer
 * (persisting on hit, streaming on miss).  mode 1 sets the DEVICE-WIDE cudaLimitPersistingL2CacheSize to the state size
 * (that much L2 is taken from everything else on the GPU; with several handles on one device the limit is the largest
 * request and is released when the last window goes) and fails with RMV_EINVAL if the state does not fit the
 * device's carve-out; 0 / -1 (auto = off) release it.  Pays when the state fits the carve-out but state + per-step
 * streams do not fit L2 (+10-13 % measured); launches that also write observation rows do not use the window.
 * Results are identical. */
RMV_API int rmv_set_l2_persist(rmv_env* env, int mode);
/* number of chunks rmv_step_host pipelines a batch through (H2D | kernel | D2H overlap); 0 = auto
 * (8 for n >= 65536, else 1) */
RMV_API int rmv_set_host_chunks(rmv_env* env, int chunks);

/* ---- state access (parity tests, checkpoint/resume, zero-copy views) --------------------- */
/* device pointers of the handle-owned buffers: state float32 [n_pad/128][S][128] (RMV_STATE_ROWS: [n_pad][S]); episode uint32 [n_pad]
 * (bit 31 = this env has returned done at least once — the reference's
 * `steps_beyond_done is not None`, quadrotor3d.py:68,114-122; bits 0..30 = number of resets,
 * the Philox episode index). */
RMV_API int rmv_state_ptr(rmv_env* env, float** state_dev, uint32_t** episode_dev);
/* copy state in/out as [n][S] rows (the reference's `env.state` tuples).  is_device != 0: the
 * buffer is a device pointer (transposed by a kernel on `stream`); == 0: a host pointer — the
 * call then synchronises `stream` before returning. */
RMV_API int rmv_set_state(rmv_env* env, const float* aos, int is_device, void* stream);
RMV_API int rmv_get_state(rmv_env* env, float* aos, int is_device, void* stream);
RMV_API int rmv_set_episode(rmv_env* env, const uint32_t* episode, int is_device, void* stream);
RMV_API int rmv_get_episode(rmv_env* env, uint32_t* episode, int is_device, void* stream);

/* ---- the hot path ------------------------------------------------------------------------ */
/* reset(): state <- U[-1,1)^S (quadrotor3d.py:182-185 and twins), episode index += 1.
 * mask_dev == NULL resets every env, else only envs with mask_dev[i] != 0.
 * obs_aos_dev, if not NULL, receives the new observation rows [n][S]. */
RMV_API int rmv_reset(rmv_env* env, const uint8_t* mask_dev, float* obs_aos_dev, void* stream);

/* step(): ONE fused kernel = dynamics + done + reward + (optional) reset-on-done for all n envs.
 * Replaces Quadrotor2D.step (quadrotor2d.py:74-113), Quadrotor2DSlungload.step
 * (quadrotor2d_slungload.py:79-154), Quadrotor3D.step (quadrotor3d.py:81-124),
 * Quadrotor3DSlungload.step (quadrotor3d_slungload.py:87-167) and, with auto_reset, the
 * reset-on-done of the caller (test/test_quadrotor3d.py:21-22) / of baselines' vec-envs behind
 * make_vec_env (run.py:89).
 *   act_dev      float32 actions, layout RMV_ACT_AOS ([n][A]) or RMV_ACT_SOA ([A][act_ld]).  RMV_ACT_AOS rows are
 *                fetched with one A*4-byte vector load each: act_dev must be aligned to A*4 bytes (16 for the 3D
 *                kinds, 8 for the planar ones; any tensor allocation is), else RMV_EALIGN — also for
 *                rmv_step_rows, rmv_step_many and rmv_step_host
 *   reward_dev   float32 [n]
 *   done_dev     uint8   [n]  (0 / 1)
 *   obs_aos_dev  NULL, or float32 [n][S]: observation rows after the step (on an RMV_STATE_ROWS handle the state
 *                buffer already is that matrix — pass NULL, or the pointer rmv_state_ptr returns; any other
 *                pointer receives a device copy) (obs == state in the
 *                reference; with auto_reset a done env reports its RESET state, the VecEnv
 *                convention).  NULL saves 4*S bytes per env: the state buffer itself is the
 *                observation in tiled component-major form.
 *   auto_reset   0: gym.Env semantics — the terminal state stays, the caller resets;
 *                1: VecEnv semantics — done envs are re-drawn inside the same kernel. */
RMV_API int rmv_step(rmv_env* env, const float* act_dev, int act_layout, int64_t act_ld,
             float* reward_dev, uint8_t* done_dev, float* obs_aos_dev, int auto_reset,
             void* stream);

/* step() on ROW-MAJOR state ("travelling state"): the state of env i is read from rows_in_dev[i][0..S) and the new
 * state (for a re-drawn env: its reset state) is written to rows_out_dev[i][0..S); the handle's own tiled state is
 * neither read nor written, its episode words / tracking buffers are used and updated as in rmv_step.  Made for
 * on-device rollout buffers (run.py:63-68 -> ppo2 Runner): obs == state in the reference (quadrotor3d.py:104), so
 * step t reads obs[t] and writes obs[t+1] — 4(2S+A+1)+1 bytes per env instead of the 4(3S+A+1)+1 of rmv_step with
 * observation rows, and the rows read are the ones the previous step just wrote.  rows_in_dev == rows_out_dev is
 * allowed (in place).  Same results as rmv_step bit for bit.  To continue with rmv_step afterwards, load the last
 * rows into the handle with rmv_set_state(env, rows, 1, stream). */
RMV_API int rmv_step_rows(rmv_env* env, const float* rows_in_dev, const float* act_dev, int act_layout, int64_t act_ld,
                          float* reward_dev, uint8_t* done_dev, float* rows_out_dev, int auto_reset, void* stream);

/* T consecutive steps with auto-reset from a pre-generated action block (random-policy or
 * replayed rollouts).  act_dev is [T][A][act_ld] (SOA) or [T][n][A] (AOS); reward_dev / done_dev
 * are [T][out_ld] or NULL.  `mode`:
 *   RMV_MANY_STREAM   (0): (the default; falls back to RMV_MANY_LAUNCHES where that is faster: batches whose step
 *               streams several hundred MB through HBM, rmv_stream_pays) the state makes the round trip through memory EVERY step (read from the state buffer,
 *               written back in place: 4(2S+A+1)+1 bytes per env-step, the traffic of T calls of rmv_step), but
 *               the T steps are ONE launch of the persistent step-stream kernel: every warp owns a fixed set of
 *               32-env units and walks (step, unit) in order, so a step depends on nothing but its own envs'
 *               previous step — no grid-wide barrier, one CTA barrier per step for the re-draws (csrc/rmv_stream.cuh);
 *               on an RMV_STATE_ROWS handle the unit's rows pass through the warp's shared tile both ways;
 *   RMV_MANY_ROLLOUT  (1): ONE launch, each thread keeps its envs' state in registers for all T steps
 *               (state is read and written once per call);
 *   RMV_MANY_LAUNCHES (2): T launches of the single-step kernel (what mode 0 was before the stream kernel; a
 *               repeated call is replayed from a CUDA graph).
 * Results are identical bit for bit in all three. */
enum { RMV_MANY_STREAM = 0, RMV_MANY_ROLLOUT = 1, RMV_MANY_LAUNCHES = 2 };
RMV_API int rmv_step_many(rmv_env* env, int n_steps, const float* act_dev, int act_layout, int64_t act_ld,
                  float* reward_dev, uint8_t* done_dev, int64_t out_ld, int mode, void* stream);
/* which kernel family the most recent step call of this handle launched (tests and benchmarks name the kernel they
 * measured; -1 before the first step) */
enum { RMV_VARIANT_STEP = 0,         /* k_step: one launch per step, tiled state                                  */
       RMV_VARIANT_ROWS = 1,         /* k_step_rows: row-major state                                              */
       RMV_VARIANT_TMA = 2,          /* k_step_tma: TMA bulk copies + mbarrier pipeline                           */
       RMV_VARIANT_ROLLOUT = 3,      /* k_rollout: state in registers for T steps                                 */
       RMV_VARIANT_STREAM = 4,       /* k_step_stream: persistent step stream, one env per thread (tiles or rows) */
       RMV_VARIANT_STREAM_PAIR = 5   /* k_step_stream_pair: step stream, two envs per thread on the packed f32 pipe */ };
RMV_API int rmv_last_variant(const rmv_env* env, int* variant);
/* 1 if rmv_step_many(RMV_MANY_STREAM) runs the stream kernel for n_envs of `kind`, 0 if it issues launches (device-free).
 * For tiled handles; an RMV_STATE_ROWS handle keeps the stream up to 64 MB of algorithmic bytes per step (its rows pass
 * through shared memory both ways) — rmv_last_variant tells what a call launched. */
RMV_API int rmv_stream_pays(int kind, int64_t n_envs);
/* Tuning of the step-stream kernel (sweeps): warps_per_sm caps the resident warps per SM (0 = as many as fit);
 * block = threads per CTA (0 auto, 32, 64, 128 — a CTA is only a scheduling unit of this kernel); action_prefetch = fetch
 * the warp's NEXT action row while the current unit is computed (-1 auto: on for small batches, 0, 1; row-per-env actions
 * only).  Results are identical. */
RMV_API int rmv_set_stream(rmv_env* env, int warps_per_sm, int block, int action_prefetch);
/* the shape the stream kernel takes for n_envs on a device with `sms` SMs and room for `resident_warps_per_sm` of its
 * warps per SM (device-free query): every warp owns units_per_warp 32-env units (the last few one less), `warps`
 * warps (a multiple of 4) own work, launched as grid x block threads. */
RMV_API int rmv_stream_shape(int64_t n_envs, int sms, int resident_warps_per_sm, int* units_per_warp, int* warps,
                             int* grid, int* block);

/* Same as rmv_step with HOST buffers (pinned memory recommended: rmv_host_alloc): copies the
 * actions host->device, runs the step kernel, copies reward/done/(obs) back and synchronises
 * `stream`.  This is what the Python VecEnv's numpy step() calls.  Large batches are pipelined in
 * chunks over internal streams (ordered after everything already queued on `stream`) so that both
 * PCIe directions and the kernel overlap. */
RMV_API int rmv_step_host(rmv_env* env, const float* act_host /*[n][A]*/, float* reward_host,
                  uint8_t* done_host, float* obs_host /*[n][S] or NULL*/, int auto_reset,
                  void* stream);
RMV_API int rmv_reset_host(rmv_env* env, float* obs_host /*[n][S]*/, void* stream);

/* ---- controllers (the caller on the input side of step() in every reference test) ---- */
/* control(): the reference's geometric controller for every env, actions written to act_dev
 * (layout as in rmv_step).  Replaces Quadrotor2D.control (quadrotor2d.py:115-138),
 * Quadrotor2DSlungload.control (quadrotor2d_slungload.py:156-183), Quadrotor3D.control
 * (quadrotor3d.py:126-180), Quadrotor3DSlungload.control (quadrotor3d_slungload.py:169-226).
 * ref_pos / ref_vel: D floats each (D = 2 or 3), NULL -> the reference's defaults
 * (ref_pos (0,0) | (0,0,2) | slung 3D (0,0,1); ref_vel 0). */
RMV_API int rmv_control(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_dev,
                        int act_layout, int64_t act_ld, void* stream);
/* control() + step() FUSED: the closed loop of the reference's own tests — `action = env.control();
 * env.step(action)` (test/test_quadrotor3d.py:16-22 and twins) — as ONE kernel: state in -> controller -> dynamics ->
 * reward / done -> reset-on-done -> state out.  No action buffer is read: 4(2S+1)+1 bytes per env-step.  act_out_dev
 * (NULL, or float32 [n][A] rows aligned to one row) receives the actions the controller produced; the other arguments
 * are those of rmv_step.  Same results bit for bit as rmv_control followed by rmv_step. */
RMV_API int rmv_control_step(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_out_dev,
                             float* reward_dev, uint8_t* done_dev, float* obs_aos_dev, int auto_reset, void* stream);
/* T closed-loop steps with reset-on-done (the 400-step loop of test/test_quadrotor3d.py:16-22): one launch of the step
 * stream where it pays (rmv_stream_pays), else one launch per step.  act_out_dev NULL or [T][n][A]; reward_dev / done_dev both
 * NULL or [T][out_ld]. */
RMV_API int rmv_control_step_many(rmv_env* env, int n_steps, const float* ref_pos, const float* ref_vel, float* act_out_dev,
                                  float* reward_dev, uint8_t* done_dev, int64_t out_ld, void* stream);
/* same with a host destination [n][A]; synchronises `stream` */
RMV_API int rmv_control_host(rmv_env* env, const float* ref_pos, const float* ref_vel, float* act_host,
                             void* stream);

/* ---- synthetic inputs --------------------------------------------------------------------- */
/* Random-policy actions ~ U[low, high) for n envs from Philox4x32-10 keyed by
 * (seed, global env id, step index): the `env.action_space.sample()` of README.md:124-131 made
 * lane-parallel.  Writes layout AOS [n][A] or SOA [A][act_ld]. */
RMV_API int rmv_fill_uniform_actions(float* act_dev, int64_t n_envs, int action_dim, int act_layout,
                             int64_t act_ld, float low, float high, uint64_t seed,
                             int64_t env_id_offset, uint32_t step_index, int device, void* stream);

/* ---- episode statistics (needs RMV_TRACK_EPISODES) ---------------------------------------- */
/* out[0] = number of finished episodes, out[1] = sum of their returns, out[2] = sum of their
 * lengths, out[3] = env-steps taken, all since creation.  Synchronises `stream`. */
RMV_API int rmv_episode_stats(rmv_env* env, double out[4], void* stream);
/* return / length of the most recently finished episode of the k envs listed in idx_host (host array of env
 * indices), gathered on the device and copied to the host arrays; synchronises `stream`.  This is how the
 * VecEnv fills info['episode'] for the envs whose done flag is set, without moving all n entries. */
RMV_API int rmv_gather_episodes(rmv_env* env, const int64_t* idx_host, int64_t k, float* return_host,
                                uint32_t* length_host, void* stream);
/* per-env return / length of the most recently finished episode, device pointers [ld] */
RMV_API int rmv_episode_ptr(rmv_env* env, float** last_return_dev, uint32_t** last_length_dev);

/* ---- VecNormalize on the device ------------------------------------------- */
/* Running observation / return normalisation of baselines' `VecNormalize` — the wrapper the reference's launcher puts
 * around a vec-env (gym_reinmav/run.py:91-92: for env_type 'mujoco'; an opt-in for the native envs).  Third-party
 * algorithm (openai/baselines@90d6677, baselines/common/vec_env/vec_normalize.py + running_mean_std.py, not vendored):
 * ret = ret*gamma + rew; ob_rms.update(obs); obs' = clip((obs - mean)/sqrt(var + eps), +-clip_obs); ret_rms.update(ret);
 * rew' = clip(rew/sqrt(ret_var + eps), +-clip_rew); ret[done] = 0 — statistics in float64, two kernels per step, no host
 * round trip.  An rmv_norm belongs to one device and one batch size; it is independent of any rmv_env. */
typedef struct rmv_norm rmv_norm;
RMV_API int rmv_norm_create(rmv_norm** out, int64_t n_envs, int obs_dim, int device, int norm_obs, int norm_ret,
                            float clip_obs /*10*/, float clip_rew /*10*/, double gamma /*0.99*/, double epsilon /*1e-8*/);
RMV_API int rmv_norm_destroy(rmv_norm* nm);
/* VecNormalize.step_wait(): raw obs rows [n][S] / reward [n] / done [n] in, normalised obs_out [n][S] / reward_out [n]
 * out (obs_dev and obs_out_dev both NULL: rewards only).  update == 0 freezes the statistics (evaluation). */
RMV_API int rmv_norm_step(rmv_norm* nm, const float* obs_dev, const float* reward_dev, const uint8_t* done_dev,
                          float* obs_out_dev, float* reward_out_dev, int update, void* stream);
/* VecNormalize.reset(): returns cleared, the reset observations enter the statistics and are normalised */
RMV_API int rmv_norm_reset(rmv_norm* nm, const float* obs_dev, float* obs_out_dev, int update, void* stream);
/* the running statistics (checkpoint / restore / logging); ob_mean, ob_var: obs_dim doubles.  Both synchronise `stream`. */
RMV_API int rmv_norm_get(rmv_norm* nm, double* ob_mean, double* ob_var, double* ob_count, double* ret_mean, double* ret_var,
                         double* ret_count, void* stream);
RMV_API int rmv_norm_set(rmv_norm* nm, const double* ob_mean, const double* ob_var, double ob_count, double ret_mean,
                         double ret_var, double ret_count, void* stream);

/* ---- pinned host memory helpers ------------------------------------------------------------ */
RMV_API int rmv_host_alloc(void** out, size_t bytes);
RMV_API int rmv_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RMV_H_ */
