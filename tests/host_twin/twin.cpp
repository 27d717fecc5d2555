// Host build of csrc/rmv_math.cuh — TEST INFRASTRUCTURE ONLY.
// Lets the CPU test-suite check, without a GPU, that the arithmetic the kernels inline
// equals the NumPy float32 model (oracle/batch.py, dtype=float32) bit for bit.
// Built by tests/host_twin/build.py with  g++ -O2 -ffp-contract=off -fno-fast-math.
#include "rmv_math.cuh"
#include "rmv_f64.cuh"
#include "rmv.h"
#include <string.h>

using namespace rmv;

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

template <int KIND>
static void run(const DevParams& P, long n, const float* s_in, const float* a_in, float* s_out,
                float* norms) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  for (long e = 0; e < n; ++e) {
    float s[S], a[A];
    memcpy(s, s_in + e * S, sizeof(s));
    memcpy(a, a_in + e * A, sizeof(a));
    EnvStep<KIND>::run(P, s, a, norms[3 * e], norms[3 * e + 1], norms[3 * e + 2]);
    memcpy(s_out + e * S, s, sizeof(s));
  }
}

extern "C" int twin_step(int kind, const rmv_params* p, long n, const float* s_in,
                         const float* a_in, float* s_out, float* norms) {
  const DevParams P = to_dev(*p);
  switch (kind) {
    case QUAD2D: run<QUAD2D>(P, n, s_in, a_in, s_out, norms); break;
    case QUAD2D_SLUNG: run<QUAD2D_SLUNG>(P, n, s_in, a_in, s_out, norms); break;
    case QUAD3D: run<QUAD3D>(P, n, s_in, a_in, s_out, norms); break;
    case QUAD3D_SLUNG: run<QUAD3D_SLUNG>(P, n, s_in, a_in, s_out, norms); break;
    default: return -1;
  }
  return 0;
}

extern "C" int twin_reset_draws(int kind, unsigned long long seed, unsigned long long env_id,
                                unsigned episode, float* out) {
  switch (kind) {
    case QUAD2D: { float s[5]; reset_draws<QUAD2D>(seed, env_id, episode, s); memcpy(out, s, sizeof(s)); } break;
    case QUAD2D_SLUNG: { float s[9]; reset_draws<QUAD2D_SLUNG>(seed, env_id, episode, s); memcpy(out, s, sizeof(s)); } break;
    case QUAD3D: { float s[10]; reset_draws<QUAD3D>(seed, env_id, episode, s); memcpy(out, s, sizeof(s)); } break;
    case QUAD3D_SLUNG: { float s[16]; reset_draws<QUAD3D_SLUNG>(seed, env_id, episode, s); memcpy(out, s, sizeof(s)); } break;
    default: return -1;
  }
  return 0;
}

// out[0..2] = pos_sq_max, vel_sq_max, inv_mass of the derived parameters
extern "C" void twin_derived(const rmv_params* p, float* out) {
  const DevParams P = to_dev(*p);
  out[0] = P.pos_sq_max; out[1] = P.vel_sq_max; out[2] = P.inv_mass; out[3] = P.inv_total_mass;
}
extern "C" float twin_sq_threshold(float t) { return sq_threshold(t); }

extern "C" void twin_sincos(long n, const float* x, float* s, float* c) {
  for (long i = 0; i < n; ++i) sincos_pinned(x[i], s[i], c[i]);
}

extern "C" void twin_fma(long n, const float* a, const float* b, const float* c, float* out) {
  for (long i = 0; i < n; ++i) out[i] = fma(rf(a[i]), rf(b[i]), rf(c[i])).v;
}

extern "C" void twin_philox(const unsigned* ctr, const unsigned* key, unsigned* out) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

// the re-draw pass of the step kernels: key-schedule Philox and the one-fma uniform
extern "C" void twin_philox_ks(unsigned long long seed, const unsigned* ctr, unsigned* out) {
  uint32_t ks[20];
  philox_key_schedule(seed, ks);
  philox4x32_10_ks(ctr[0], ctr[1], ctr[2], ctr[3], ks, out);
}
extern "C" void twin_uniform_pm1(long n, const unsigned* bits, float* fast, float* plain) {
  for (long i = 0; i < n; ++i) { fast[i] = uniform_pm1(bits[i]); plain[i] = uniform_from_bits(bits[i], -1.0f, 2.0f); }
}

// pinned atan2 and the geometric controllers (control_law, rmv_math.cuh)
extern "C" void twin_atan2(long n, const float* y, const float* x, float* out) {
  for (long i = 0; i < n; ++i) out[i] = atan2_pinned(y[i], x[i]);
}

template <int KIND>
static void run_control(const CtrlParams& C, long n, const float* s_in, float* a_out) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
  for (long e = 0; e < n; ++e) {
    float s[S], a[A];
    memcpy(s, s_in + e * S, sizeof(s));
    control_law<KIND>(C, s, a);
    memcpy(a_out + e * A, a, sizeof(a));
  }
}

// ctrl = {ref_pos[3], ref_vel[3], mass, gravity}
extern "C" int twin_control(int kind, const float* ctrl, long n, const float* s_in, float* a_out) {
  CtrlParams C;
  for (int i = 0; i < 3; ++i) { C.ref_pos[i] = ctrl[i]; C.ref_vel[i] = ctrl[3 + i]; }
  C.mass = ctrl[6]; C.gravity = ctrl[7];
  switch (kind) {
    case QUAD2D: run_control<QUAD2D>(C, n, s_in, a_out); break;
    case QUAD2D_SLUNG: run_control<QUAD2D_SLUNG>(C, n, s_in, a_out); break;
    case QUAD3D: run_control<QUAD3D>(C, n, s_in, a_out); break;
    case QUAD3D_SLUNG: run_control<QUAD3D_SLUNG>(C, n, s_in, a_out); break;
    default: return -1;
  }
  return 0;
}

// float64 asin / atan2 of k_reinmav_step (rmv_f64.cuh)
extern "C" void twin_atan2_f64(long n, const double* y, const double* x, double* out) {
  for (long i = 0; i < n; ++i) out[i] = atan2_f64(y[i], x[i]);
}
extern "C" void twin_asin_f64(long n, const double* x, double* out) {
  for (long i = 0; i < n; ++i) out[i] = asin_f64(x[i]);
}
