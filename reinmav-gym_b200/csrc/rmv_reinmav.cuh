// rmv_reinmav.cuh — reinmav-v0: the closed-loop 13-state rigid body of
// gym_reinmav/envs/native/reinmav_env.py (position, velocity, quaternion, body rates; built-in PD
// controller tracking a quintic trajectory; no action input; reward 90, done True every step).
//
// One thread per env, state float32 in the tiled layout [n_tiles][13][128]; ONE step = the reference's
// explicit Euler over np.arange(t, t+dt, 1/5000) (50 or 51 substeps, reinmav_env.py:90-98), evaluated in
// FLOAT64 registers: ~300 flops x 50 substeps per env-step is compute, not bandwidth, and float32 would not
// hold 1e-5 against the reference over the 20 000 substeps of its own 4 s test trajectory.
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

struct ReinmavArgs {
  float* state;        // tiled [n_tiles][13][128]
  float* reward;       // [n] or nullptr
  uint8_t* done;       // [n] or nullptr
  float* obs;          // [n][13] or nullptr
  int64_t n;
  double t0, delta;    // substep i runs at t0 + i * delta
  int n_sub;
  double inv_inertia[9];
};

constexpr int kReinmavS = 13;
constexpr int kReinmavMaxSub = 64;     // np.arange(t, t + 0.01, 1/5000) has 50 or 51 elements

// What every env shares at one substep: the quintic reference trajectory (trj_gen, reinmav_env.py:128-136) and
// sin/cos of the desired yaw (controller, :306-337) depend on the clock only, so a CTA evaluates them once per
// substep into shared memory instead of every thread doing it 50 times.
struct ReinmavRef {
  double d_pos, d_vel, d_acc, sp, cp;
};
__device__ __forceinline__ ReinmavRef reinmav_reference(double t) {
  ReinmavRef r;
  const double tau = fmax(0.0, fmin(t, 4.0)) / 4.0;
  const double tau2 = tau * tau, tau3 = tau2 * tau, tau4 = tau3 * tau, tau5 = tau4 * tau;
  r.d_pos = 10.0 * tau3 - 15.0 * tau4 + 6.0 * tau5;
  r.d_vel = (30 / 4.0) * tau2 - (60 / 4.0) * tau3 + (30 / 4.0) * tau4;
  r.d_acc = (60 / 16.0) * tau - (180 / 16.0) * tau2 + (120 / 16.0) * tau3;
  sincos(r.d_pos, &r.sp, &r.cp);                             // psi_des = desired yaw = d_pos
  return r;
}

__device__ __forceinline__ void reinmav_derivative(const double (&s)[13], const ReinmavRef& ref, const double (&invI)[9],
                                                   double (&ds)[13]) {
  constexpr double ARM = 0.0860, MASS = 0.1800, GRAV = 9.8100, MIN_F = 0.0, MAX_F = 3.5316;   // reinmav_env.py:55-59
  constexpr double I00 = 0.00025, I02 = 2.55e-06, I11 = 0.000232, I20 = 2.55e-06, I22 = 0.0003738;   // :60-62
  const double qw = s[6], qx = s[7], qy = s[8], qz = s[9], p = s[10], q = s[11], r = s[12];
  // quat2mat (:267-290); only the entries the controller and the force projection read
  const double nq = qw * qw + qx * qx + qy * qy + qz * qz;
  double R02 = 0.0, R10 = 0.0, R11 = 1.0, R12 = 0.0, R20 = 0.0, R21 = 0.0, R22 = 1.0;
  if (nq > 2.220446049250313e-16) {
    const double sc = 2.0 / nq;
    const double X = qx * sc, Y = qy * sc, Z = qz * sc;
    const double wX = qw * X, wY = qw * Y, wZ = qw * Z, xX = qx * X, xY = qx * Y, xZ = qx * Z;
    const double yY = qy * Y, yZ = qy * Z, zZ = qz * Z;
    R02 = xZ + wY;
    R10 = xY + wZ; R11 = 1.0 - (xX + zZ); R12 = yZ - wX;
    R20 = xZ - wY; R21 = yZ + wX; R22 = 1.0 - (xX + yY);
  }
  // RotToRPY (:341-346): phi = asin(R12), psi = atan2(-R10/cos(phi), R11/cos(phi)), theta = atan2(-R02/cos(phi),
  // R22/cos(phi)).  cos(asin(x)) = sqrt(1 - x^2) > 0 for |x| < 1, and atan2 is invariant under a common positive
  // factor, so the two divisions and the cosine are not evaluated (differences at the 1e-16 level, far inside the
  // stated 1e-5; at |R12| = 1 the reference itself divides by zero).
  const double phi = asin(R12);
  const double psi = atan2(-R10, R11);
  const double theta = atan2(-R02, R22);
  // controller (:306-337)
  const double kp[3] = {10.0, 10.0, 35.0}, kd[3] = {5.0, 5.0, 22.0};
  double ddr[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) ddr[i] = ref.d_acc + kd[i] * (ref.d_vel - s[3 + i]) + kp[i] * (ref.d_pos - s[i]);
  const double u1 = MASS * (GRAV + ddr[2]);
  const double phi_des = 1 / GRAV * (ddr[0] * ref.sp - ddr[1] * ref.cp);
  const double theta_des = 1 / GRAV * (ddr[0] * ref.cp + ddr[1] * ref.sp);
  const double mx = 100.0 * (phi_des - phi) - 0.1 * p;
  const double my = 100.0 * (theta_des - theta) - 0.1 * q;
  const double mz = 100.0 * (ref.d_pos - psi) + 0.1 * (ref.d_vel - r);
  // motor mixing + clamp (:204-216)
  const double k = 0.5 / ARM;
  double T[4] = {0.25 * u1 - k * my, 0.25 * u1 + k * mx, 0.25 * u1 + k * my, 0.25 * u1 - k * mx};
#pragma unroll
  for (int i = 0; i < 4; ++i) T[i] = fmax(fmin(T[i], MAX_F / 4.0), MIN_F / 4.0);
  const double force = T[0] + T[1] + T[2] + T[3];
  const double mom[3] = {ARM * T[1] - ARM * T[3], -ARM * T[0] + ARM * T[2], mz};
  ds[0] = s[3]; ds[1] = s[4]; ds[2] = s[5];
  ds[3] = 1.0 / MASS * (R20 * force);
  ds[4] = 1.0 / MASS * (R21 * force);
  ds[5] = 1.0 / MASS * (R22 * force - MASS * GRAV);
  const double err = 1 - nq;
  ds[6] = -0.5 * (-p * qx - q * qy - r * qz) + 2.0 * err * qw;
  ds[7] = -0.5 * (p * qw - r * qy + q * qz) + 2.0 * err * qx;
  ds[8] = -0.5 * (q * qw + r * qx - p * qz) + 2.0 * err * qy;
  ds[9] = -0.5 * (r * qw - q * qx + p * qy) + 2.0 * err * qz;
  const double Iw[3] = {I00 * p + I02 * r, I11 * q, I20 * p + I22 * r};
  const double rhs[3] = {mom[0] - (q * Iw[2] - r * Iw[1]), mom[1] - (r * Iw[0] - p * Iw[2]), mom[2] - (p * Iw[1] - q * Iw[0])};
#pragma unroll
  for (int i = 0; i < 3; ++i) ds[10 + i] = invI[3 * i] * rhs[0] + invI[3 * i + 1] * rhs[1] + invI[3 * i + 2] * rhs[2];
}

__global__ void __launch_bounds__(128) k_reinmav_step(const ReinmavArgs a) {
  __shared__ ReinmavRef ref[kReinmavMaxSub];
  if ((int)threadIdx.x < a.n_sub) ref[threadIdx.x] = reinmav_reference(a.t0 + (double)threadIdx.x * a.delta);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  float* const row = a.state + ((e >> 7) * (int64_t)(kReinmavS * kTile) + (e & (kTile - 1)));
  double s[13], ds[13], invI[9];
#pragma unroll
  for (int c = 0; c < 13; ++c) s[c] = (double)row[c * kTile];
#pragma unroll
  for (int i = 0; i < 9; ++i) invI[i] = a.inv_inertia[i];
  const double h = 1.0 / 5000;
  for (int i = 0; i < a.n_sub; ++i) {
    reinmav_derivative(s, ref[i], invI, ds);
#pragma unroll
    for (int c = 0; c < 13; ++c) s[c] = s[c] + h * ds[c];
  }
#pragma unroll
  for (int c = 0; c < 13; ++c) {
    const float x = (float)s[c];
    row[c * kTile] = x;
    if (a.obs) a.obs[e * kReinmavS + c] = x;
  }
  if (a.reward) a.reward[e] = 90.0f;            // reward = 100 - 10, done = True (reinmav_env.py:111-116)
  if (a.done) a.done[e] = 1;
}

__global__ void k_reinmav_init(float* state, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* const row = state + ((e >> 7) * (int64_t)(kReinmavS * kTile) + (e & (kTile - 1)));
  for (int c = 0; c < 13; ++c) row[c * kTile] = (c == 6) ? 1.0f : 0.0f;     // init_state (:80): unit quaternion at rest
}

}  // namespace rmv
