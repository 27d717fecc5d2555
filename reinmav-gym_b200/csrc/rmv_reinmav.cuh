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
#include "rmv_f64.cuh"

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

constexpr double kRmArm = 0.0860, kRmMass = 0.1800, kRmGrav = 9.8100, kRmMaxF = 3.5316;   // reinmav_env.py:55-59 (MIN_F = 0)
#define RMV_REINMAV_GAINS const double kRmKp[3] = {10.0, 10.0, 35.0}, kRmKd[3] = {5.0, 5.0, 22.0}   /* controller gains (:311-312) */
constexpr double kRmKpRot = 100.0, kRmKdRot = 0.1;                                // (:313-314)
constexpr double kRmMix = 0.5 / kRmArm;                                           // moment -> thrust difference (:204-212)
constexpr double kRmI00 = 0.00025, kRmI02 = 2.55e-06, kRmI11 = 0.000232, kRmI20 = 2.55e-06, kRmI22 = 0.0003738;   // (:60-62)

// The constants of a substep that are not exact in 32 bits.  In the constant bank an FP64 instruction reads them as an
// operand; written as literals each use costs two uniform-register moves inside the substep loop.
struct ReinmavConst {
  double quarter_mass, quarter_weight, neg_rot_p, neg_rot_d, neg_kd_rot, thrust_cap, inv_mass, neg_grav, arm;
  double i00, i02, i11, i20, i22, h;
};
__constant__ ReinmavConst kRmC = {0.25 * kRmMass, 0.25 * kRmMass * kRmGrav, -kRmKpRot * kRmMix, -kRmKdRot * kRmMix, -kRmKdRot,
                                  kRmMaxF / 4.0, 1.0 / kRmMass, -kRmGrav, kRmArm,
                                  kRmI00, kRmI02, kRmI11, kRmI20, kRmI22, 1.0 / 5000};

// What every env shares at one substep: the quintic reference trajectory (trj_gen, reinmav_env.py:128-136) and
// sin/cos of the desired yaw (controller, :306-337) depend on the clock only, so a CTA evaluates them once per
// substep into shared memory instead of every thread doing it 50 times — together with every product of a gain and a
// reference value, so that a thread's share of the PD laws is one fused multiply-add per state component:
//   ddr_i = d_acc + kd_i (d_vel - v_i) + kp_i (d_pos - x_i)            = acc0[i] - kd_i v_i - kp_i x_i
//   k mx  = k (100 (phi_des - phi) - 0.1 p),  phi_des = (ddr_0 sp - ddr_1 cp) / g
//                                                                       = ddr_0 sp_k - ddr_1 cp_k - 100 k phi - 0.1 k p
//   mz    = 100 (d_pos - psi) + 0.1 (d_vel - r)                        = mz0 - 100 psi - 0.1 r
// (k = 0.5 / arm, the motor-mixing factor; sp_k = 100 k sin(psi_des) / g).  The regrouping moves results at the
// 1e-16 level (see reinmav_derivative).
struct ReinmavRef {
  double acc0[3], sp_k, cp_k, mz0;
};
__device__ __forceinline__ ReinmavRef reinmav_reference(double t) {
  ReinmavRef r;
  RMV_REINMAV_GAINS;
  const double tau = fmax(0.0, fmin(t, 4.0)) / 4.0;
  const double tau2 = tau * tau, tau3 = tau2 * tau, tau4 = tau3 * tau, tau5 = tau4 * tau;
  const double d_pos = 10.0 * tau3 - 15.0 * tau4 + 6.0 * tau5;
  const double d_vel = (30 / 4.0) * tau2 - (60 / 4.0) * tau3 + (30 / 4.0) * tau4;
  const double d_acc = (60 / 16.0) * tau - (180 / 16.0) * tau2 + (120 / 16.0) * tau3;
  double sp, cp;
  sincos(d_pos, &sp, &cp);                                   // psi_des = desired yaw = d_pos
#pragma unroll
  for (int i = 0; i < 3; ++i) r.acc0[i] = d_acc + kRmKd[i] * d_vel + kRmKp[i] * d_pos;
  r.sp_k = kRmKpRot * kRmMix * (1 / kRmGrav * sp);
  r.cp_k = kRmKpRot * kRmMix * (1 / kRmGrav * cp);
  r.mz0 = kRmKpRot * d_pos + kRmKdRot * d_vel;
  return r;
}

// The arithmetic below is the reference's, regrouped and contracted by hand into fused multiply-adds (the library is
// compiled with --fmad=false for the float32 step kernels, and this kernel is bound by the FP64 pipe: a DMUL + DADD
// pair costs twice a DFMA).  That changes results at the 1e-16 level; reinmav-v0 is held to 1e-5 against the
// reference and 2e-6 against the float64 oracle with the same float32 storage rule (tests/test_reinmav.py).
__device__ __forceinline__ void reinmav_derivative(const double (&s)[13], const ReinmavRef& ref, const double (&invI)[9],
                                                   double (&ds)[13]) {
  RMV_REINMAV_GAINS;
  const double qw = s[6], qx = s[7], qy = s[8], qz = s[9], p = s[10], q = s[11], r = s[12];
  // quat2mat (:267-290); only the entries the controller and the force projection read
  const double nq = ::fma(qz, qz, ::fma(qy, qy, ::fma(qx, qx, qw * qw)));
  double R02 = 0.0, R10 = 0.0, R11 = 1.0, R12 = 0.0, R20 = 0.0, R21 = 0.0, R22 = 1.0;
  if (nq > 2.220446049250313e-16) {                          // np.finfo(float).eps (:271); false for NaN
    const double sc = normal_range(nq) ? div_newton(2.0, nq) : div_rare(2.0, nq);
    const double X = qx * sc, Y = qy * sc, Z = qz * sc;
    const double xX = qx * X, xY = qx * Y, xZ = qx * Z, yZ = qy * Z;
    R02 = ::fma(qw, Y, xZ);
    R10 = ::fma(qw, Z, xY); R11 = 1.0 - ::fma(qz, Z, xX); R12 = ::fma(-qw, X, yZ);
    R20 = ::fma(-qw, Y, xZ); R21 = ::fma(qw, X, yZ); R22 = 1.0 - ::fma(qy, Y, xX);
  }
  // RotToRPY (:341-346): phi = asin(R12), psi = atan2(-R10/cos(phi), R11/cos(phi)), theta = atan2(-R02/cos(phi),
  // R22/cos(phi)).  cos(asin(x)) = sqrt(1 - x^2) > 0 for |x| < 1, and atan2 is invariant under a common positive
  // factor, so the two divisions and the cosine are not evaluated (differences at the 1e-16 level, far inside the
  // stated 1e-5; at |R12| = 1 the reference itself divides by zero).  asin_f64 / atan2_f64 (rmv_f64.cuh): double
  // accuracy at a quarter of the library versions' FP64 instructions.
  double phi, psi, theta;
  if (asin_fast_ok(R12) & atan2_fast_ok(-R10, R11) & atan2_fast_ok(-R02, R22)) {          // one branch for the three
    phi = asin_core(R12); psi = atan2_core(-R10, R11); theta = atan2_core(-R02, R22);
  } else {
    phi = asin_f64(R12); psi = atan2_f64(-R10, R11); theta = atan2_f64(-R02, R22);
  }
  // controller (:306-337) and motor mixing + clamp (:204-216)
  double ddr[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) ddr[i] = ::fma(-kRmKp[i], s[i], ::fma(-kRmKd[i], s[3 + i], ref.acc0[i]));
  const double u4 = ::fma(kRmC.quarter_mass, ddr[2], kRmC.quarter_weight);             // u1 / 4, u1 = m (g + ddr_z)
  const double kmx = ::fma(kRmC.neg_rot_p, phi, ::fma(kRmC.neg_rot_d, p, ::fma(ddr[0], ref.sp_k, -(ddr[1] * ref.cp_k))));
  const double kmy = ::fma(kRmC.neg_rot_p, theta, ::fma(kRmC.neg_rot_d, q, ::fma(ddr[0], ref.cp_k, ddr[1] * ref.sp_k)));
  const double mz = ::fma(-kRmKpRot, psi, ::fma(kRmC.neg_kd_rot, r, ref.mz0));
  double T[4] = {u4 - kmy, u4 + kmx, u4 + kmy, u4 - kmx};
#pragma unroll
  for (int i = 0; i < 4; ++i) T[i] = clamp_zero_to(T[i], kRmC.thrust_cap);      // np.clip(T, MIN_F / 4 = 0, MAX_F / 4)
  const double fm = kRmC.inv_mass * (T[0] + T[1] + T[2] + T[3]);                        // force / mass
  ds[0] = s[3]; ds[1] = s[4]; ds[2] = s[5];
  ds[3] = R20 * fm;
  ds[4] = R21 * fm;
  ds[5] = ::fma(R22, fm, kRmC.neg_grav);
  const double err2 = 2.0 * (1 - nq), hp = 0.5 * p, hq = 0.5 * q, hr = 0.5 * r;
  ds[6] = ::fma(hr, qz, ::fma(hq, qy, ::fma(hp, qx, err2 * qw)));
  ds[7] = ::fma(-hq, qz, ::fma(hr, qy, ::fma(-hp, qw, err2 * qx)));
  ds[8] = ::fma(hp, qz, ::fma(-hr, qx, ::fma(-hq, qw, err2 * qy)));
  ds[9] = ::fma(-hp, qy, ::fma(hq, qx, ::fma(-hr, qw, err2 * qz)));
  const double Iw[3] = {::fma(kRmC.i02, r, kRmC.i00 * p), kRmC.i11 * q, ::fma(kRmC.i22, r, kRmC.i20 * p)};
  // moments minus omega x (I omega)
  const double rhs[3] = {::fma(-q, Iw[2], ::fma(r, Iw[1], kRmC.arm * (T[1] - T[3]))),
                         ::fma(-r, Iw[0], ::fma(p, Iw[2], kRmC.arm * (T[2] - T[0]))),
                         ::fma(-p, Iw[1], ::fma(q, Iw[0], mz))};
  // I^-1 rhs: the pitch axis decouples ([[a,0,c],[0,e,0],[g,0,i]] and its inverse share the zeros, exact on the host side)
  ds[10] = ::fma(invI[2], rhs[2], invI[0] * rhs[0]);
  ds[11] = invI[4] * rhs[1];
  ds[12] = ::fma(invI[8], rhs[2], invI[6] * rhs[0]);
}

__global__ void __launch_bounds__(128, 7) k_reinmav_step(const ReinmavArgs a) {
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
  const double h = kRmC.h;                                   // 1/5000 (:91)
  for (int i = 0; i < a.n_sub; ++i) {
    reinmav_derivative(s, ref[i], invI, ds);
#pragma unroll
    for (int c = 0; c < 13; ++c) s[c] = ::fma(h, ds[c], s[c]);
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
