// rmv_math.cuh — per-env arithmetic of the four native quadrotor step() functions.
//
// One env = one set of registers; no memory access in this file.  Every floating-point
// operation goes through rf / rd below, which map to the individually rounded IEEE
// intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn, __fmaf_rn ...).  Those are never
// contracted or re-associated by the compiler, so the evaluation order written here IS the
// arithmetic: it is the reference's own left-to-right NumPy expression order
// (gym_reinmav/envs/native/*.py, cited per function) in float32, with `x*y + acc` steps of that
// order fused into one FFMA where written as fma(x, y, acc) (one rounding instead of two: fewer
// instructions AND closer to the float64 reference).  tests/ check the kernel against an
// independent NumPy float32 evaluation of the same sequence (exact fma emulation,
// oracle/fma32.py) bit for bit.
//
// The file also compiles as plain C++ (tests/host_twin builds it with g++ -ffp-contract=off)
// so the arithmetic can be checked in a container without a GPU.  That build is test
// infrastructure only; the product has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RMV_HD __host__ __device__ __forceinline__
#else
#define RMV_HD inline
#endif

namespace rmv {

// ------------------------------------------------------------------------------------------
// rounded scalar types
// ------------------------------------------------------------------------------------------
struct rf {
  float v;
  RMV_HD rf() : v(0.0f) {}
  RMV_HD rf(float x) : v(x) {}
};
#if defined(__CUDA_ARCH__)
RMV_HD rf operator+(rf a, rf b) { return rf(__fadd_rn(a.v, b.v)); }
RMV_HD rf operator-(rf a, rf b) { return rf(__fsub_rn(a.v, b.v)); }
RMV_HD rf operator*(rf a, rf b) { return rf(__fmul_rn(a.v, b.v)); }
RMV_HD rf operator/(rf a, rf b) { return rf(__fdiv_rn(a.v, b.v)); }
RMV_HD rf rsqrt_exact(rf a) { return rf(__fsqrt_rn(a.v)); }
#else
RMV_HD rf operator+(rf a, rf b) { return rf(a.v + b.v); }
RMV_HD rf operator-(rf a, rf b) { return rf(a.v - b.v); }
RMV_HD rf operator*(rf a, rf b) { return rf(a.v * b.v); }
RMV_HD rf operator/(rf a, rf b) { return rf(a.v / b.v); }
RMV_HD rf rsqrt_exact(rf a) { return rf(sqrtf(a.v)); }
#endif
RMV_HD rf operator-(rf a) { return rf(-a.v); }
RMV_HD rf sqrt(rf a) { return rsqrt_exact(a); }
// 1/a correctly rounded: the same value as the IEEE quotient 1.0f/a for every input; __frcp_rn is the shorter sequence
RMV_HD rf rcp(rf a) {
#if defined(__CUDA_ARCH__)
  return rf(__frcp_rn(a.v));
#else
  return rf(1.0f / a.v);
#endif
}
// a*b + c with ONE rounding (FFMA).  Used exactly where the float32 model (oracle/batch.py `_fma`)
// uses its exact fma; everything else stays individually rounded.
RMV_HD rf fma(rf a, rf b, rf c) {
#if defined(__CUDA_ARCH__)
  return rf(__fmaf_rn(a.v, b.v, c.v));
#else
  return rf(fmaf(a.v, b.v, c.v));
#endif
}

struct rd {
  double v;
  RMV_HD rd() : v(0.0) {}
  RMV_HD rd(double x) : v(x) {}
};
#if defined(__CUDA_ARCH__)
RMV_HD rd operator+(rd a, rd b) { return rd(__dadd_rn(a.v, b.v)); }
RMV_HD rd operator-(rd a, rd b) { return rd(__dsub_rn(a.v, b.v)); }
RMV_HD rd operator*(rd a, rd b) { return rd(__dmul_rn(a.v, b.v)); }
RMV_HD rd sqrt(rd a) { return rd(__dsqrt_rn(a.v)); }
#else
RMV_HD rd operator+(rd a, rd b) { return rd(a.v + b.v); }
RMV_HD rd operator-(rd a, rd b) { return rd(a.v - b.v); }
RMV_HD rd operator*(rd a, rd b) { return rd(a.v * b.v); }
RMV_HD rd sqrt(rd a) { return rd(::sqrt(a.v)); }
#endif

// ------------------------------------------------------------------------------------------
// constants of one handle, pre-rounded to float32 on the host
// ------------------------------------------------------------------------------------------
struct DevParams {
  float mass, load_mass, dt, neg_gravity, tether_len;
  float pos_done, vel_done, thrust_scale;
  float first_done_reward, later_done_reward;
  float reward_scale;     // baselines RewardScaler (run.py:89 reward_scale=): returned reward = reward * reward_scale
  float inv_total_mass;   // 1.0f / (mass + load_mass), two float32 ops
  float inv_mass;         // 1.0f / mass (exactly 1 for the reference's mass = 1)
  float pos_sq_max;       // largest float32 y with sqrt_rn(y) <= pos_done: |p| > pos_done  <=>  p.p > pos_sq_max
  float vel_sq_max;       // same for vel_done
  int thrust_clamp;
};

// Host-side completion of the derived fields.  Every value is produced by single float32 IEEE
// operations so that the NumPy float32 model (oracle/batch.py) computes the same bits.
inline float sq_threshold(float t) {
  // sqrt_rn is monotone, so {y : sqrt_rn(y) <= t} is a down-set of the non-negative floats; its
  // maximum turns the reference's `norm > t` into a comparison of the squared norm, bit for bit.
  if (t != t) return t;                       // NaN threshold: never done
  if (t < 0.0f) return -1.0f;                 // any norm exceeds a negative threshold
  if (t == INFINITY) return INFINITY;
  volatile float y = t * t;
  if (y == INFINITY) y = 3.402823466e+38f;
  while (sqrtf(nextafterf(y, INFINITY)) <= t && y < 3.402823466e+38f) y = nextafterf(y, INFINITY);
  while (sqrtf(y) > t) y = nextafterf(y, -INFINITY);
  return y;
}
inline void finish_params(DevParams& d) {
  volatile float total = d.mass + d.load_mass;
  d.inv_total_mass = 1.0f / total;
  volatile float m = d.mass;
  d.inv_mass = 1.0f / m;
  d.pos_sq_max = sq_threshold(d.pos_done);
  d.vel_sq_max = sq_threshold(d.vel_done);
}

enum { QUAD2D = 0, QUAD2D_SLUNG = 1, QUAD3D = 2, QUAD3D_SLUNG = 3 };

template <int KIND> struct Dims;
template <> struct Dims<QUAD2D> { static constexpr int S = 5, A = 2, D = 2; };
template <> struct Dims<QUAD2D_SLUNG> { static constexpr int S = 9, A = 2, D = 2; };
template <> struct Dims<QUAD3D> { static constexpr int S = 10, A = 4, D = 3; };
template <> struct Dims<QUAD3D_SLUNG> { static constexpr int S = 16, A = 4, D = 3; };

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter/key convention: oracle/philox.py.
// ------------------------------------------------------------------------------------------
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t STREAM_RESET_BASE = 0u, STREAM_ACTION = 16u;

RMV_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// top 24 bits -> [0,1), exact; then lo + (hi-lo)*u with two rounded float32 ops
RMV_HD float uniform_from_bits(uint32_t bits, float lo, float width) {
  const rf u = rf((float)(bits >> 8)) * rf(5.9604644775390625e-08f);  // 2^-24, exact product
  return (rf(lo) + rf(width) * u).v;
}

// Philox4x32-10 with the key schedule ks[2r] = k0 + r*W0, ks[2r+1] = k1 + r*W1 precomputed by the host (the kernels
// read it from the kernel-parameter constant bank): per round two wide multiplies and two three-input XORs.
// Same function as philox4x32_10 (tests/test_philox.py).
RMV_HD void philox_key_schedule(uint64_t seed, uint32_t (&ks)[20]) {
  for (uint32_t r = 0; r < 10; ++r) {
    ks[2 * r] = (uint32_t)seed + r * PHILOX_W0;
    ks[2 * r + 1] = (uint32_t)(seed >> 32) + r * PHILOX_W1;
  }
}
RMV_HD void philox4x32_10_ks(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&ks)[20],
                             uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
    const uint32_t hi0 = __umulhi(PHILOX_M0, c0), hi1 = __umulhi(PHILOX_M1, c2);
#else
    const uint32_t hi0 = (uint32_t)(((uint64_t)PHILOX_M0 * c0) >> 32), hi1 = (uint32_t)(((uint64_t)PHILOX_M1 * c2) >> 32);
#endif
    const uint32_t lo0 = PHILOX_M0 * c0, lo1 = PHILOX_M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ ks[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ ks[2 * r + 1];
    c1 = lo1; c3 = lo0; c0 = n0; c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// U[-1,1) from the top 24 bits: -1 + u24 * 2^-23 in ONE fma.  Every intermediate of
// uniform_from_bits(bits, -1, 2) is exact (multiples of 2^-23 of magnitude <= 1), so this is the same value
// (checked for all 2^24 inputs in tests/test_philox.py).
RMV_HD float uniform_pm1(uint32_t bits) {
  return fma(rf((float)(bits >> 8)), rf(1.1920928955078125e-07f), rf(-1.0f)).v;
}

// reset(): S draws of U[-1,1) for (seed, global env id, episode) — replaces
// self.np_random.uniform(low=-1.0, high=1.0, size=(S,)), quadrotor3d.py:184 and twins.
template <int KIND>
RMV_HD void reset_draws(uint64_t seed, uint64_t env_id, uint32_t episode, float s[Dims<KIND>::S]) {
  constexpr int S = Dims<KIND>::S;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int b = 0; b < (S + 3) / 4; ++b) {
    uint32_t w[4];
    philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), episode,
                  ((STREAM_RESET_BASE + (uint32_t)KIND) << 8) | (uint32_t)b, k0, k1, w);
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (4 * b + l < S) s[4 * b + l] = uniform_from_bits(w[l], -1.0f, 2.0f);
  }
}

// ------------------------------------------------------------------------------------------
// sin / cos in float32 with a pinned sequence (mirror of oracle/sincos32.py)
// ------------------------------------------------------------------------------------------
RMV_HD uint32_t bits_of(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
RMV_HD float float_of(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

RMV_HD void sincos_pinned(float xin, float& s_out, float& c_out) {
  const rf x(xin);
  rf r;
  int quad;
  if (fabsf(xin) <= 1.0e5f) {
    const rf kf(rintf((x * rf(0.63661977236758134f)).v));
    r = fma(-kf, rf(7.54978995489188216e-8f), fma(-kf, rf(4.837512969970703125e-4f), fma(-kf, rf(1.5703125f), x)));
    quad = ((int)kf.v) & 3;
  } else {  // huge, inf or NaN: reduce in float64 (rare, exact-sequence as well)
    const rd xd((double)xin);
    const rd kd(rint((xd * rd(0.6366197723675814)).v));
    const rd rr = (xd - kd * rd(1.5707963267948966)) - kd * rd(6.123233995736766e-17);
    const rd qd = kd - rd(4.0) * rd(floor((kd * rd(0.25)).v));
    r = rf((float)rr.v);
    quad = (qd.v == qd.v && fabs(qd.v) < 8.0) ? (int)qd.v : 0;
  }
  const rf z = r * r;
  rf t = fma(rf(-1.9515295891e-4f), z, rf(8.3321608736e-3f));
  t = fma(t, z, rf(-1.6666654611e-1f));
  t = t * z;
  const rf sr = fma(t, r, r);
  t = fma(rf(2.443315711809948e-5f), z, rf(-1.388731625493765e-3f));
  t = fma(t, z, rf(4.166664568298827e-2f));
  t = t * z;
  t = t * z;
  t = fma(rf(-0.5f), z, t);
  const rf cr = t + rf(1.0f);
  // quadrant fix-up without branches: odd quadrants swap sin/cos, the sign bits follow the quadrant
  //   q:      0    1    2    3
  //   sin:   sr   cr  -sr  -cr
  //   cos:   cr  -sr  -cr   sr
  const bool swap = (quad & 1) != 0;
  const uint32_t sb = swap ? bits_of(cr.v) : bits_of(sr.v);
  const uint32_t cb = swap ? bits_of(sr.v) : bits_of(cr.v);
  s_out = float_of(sb ^ (((uint32_t)quad & 2u) << 30));
  c_out = float_of(cb ^ ((((uint32_t)quad + 1u) & 2u) << 30));
}

// ------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------
template <int D> RMV_HD rf sumsq(const rf (&c)[D]) {
  rf acc = c[0] * c[0];
#pragma unroll
  for (int i = 1; i < D; ++i) acc = fma(c[i], c[i], acc);
  return acc;
}
template <int D> RMV_HD rf norm(const rf (&c)[D]) { return sqrt(sumsq<D>(c)); }

// x' and v' of one body.  VEL_FIRST = quadrotor2d_slungload.py:100-101,109-110 (velocity, then
// position with the NEW velocity); otherwise quadrotor3d.py:98-99 (position with the OLD velocity).
template <int D, bool VEL_FIRST>
RMV_HD void advance(rf (&x)[D], rf (&v)[D], const rf (&a)[D], rf dt) {
  const rf half(0.5f);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    // x + v*dt + ((0.5*a)*dt)*dt, left to right, each "+ product" step fused
    if (VEL_FIRST) {
      v[i] = fma(a[i], dt, v[i]);
      x[i] = fma(half * a[i] * dt, dt, fma(v[i], dt, x[i]));
    } else {
      x[i] = fma(half * a[i] * dt, dt, fma(v[i], dt, x[i]));
      v[i] = fma(a[i], dt, v[i]);
    }
  }
}

// pyquaternion semantics of Quaternion(att).rotation_matrix.dot([0,0,1]) followed by
// .derivative(w) on the same (now normalised) object — quadrotor3d.py:94-96,101-102.
// in:  q (un-normalised state quaternion), om (body rates)
// out: e3 = R(q^) e_z, qdot = (0.5 q^) (x) (0, om)
RMV_HD void attitude_3d(const rf (&q)[4], const rf (&om)[3], rf (&e3)[3], rf (&qdot)[4]) {
  rf w = q[0], x = q[1], y = q[2], z = q[3];
  const rf n2 = fma(z, z, fma(y, y, fma(x, x, w * w)));
  const rf n = sqrt(n2);
  // pyquaternion skips the division when |1 - q.q| < 1e-14; in float32 that means q.q == 1, where
  // multiplying by 1/sqrt(1) = 1 is exact, so only the n > 0 guard is left.  float32 normalises
  // with ONE IEEE reciprocal and four products (x/n -> x*(1/n)): <= 1 ulp from the quotient.
  if (n.v > 0.0f) {
    const rf inv = rcp(n);
    w = w * inv; x = x * inv; y = y * inv; z = z * inv;
  }
  // column 3 of (Q Qbar^T)[1:,1:], four-term sums left to right
  e3[0] = fma(y, w, fma(z, x, fma(w, y, x * z)));
  e3[1] = fma(-x, w, fma(-w, x, fma(z, y, y * z)));
  e3[2] = fma(w, w, fma(-x, x, fma(-y, y, z * z)));
  const rf h(0.5f);
  const rf hw = h * w, hx = h * x, hy = h * y, hz = h * z;
  qdot[0] = fma(-hz, om[2], fma(-hy, om[1], (-hx) * om[0]));
  qdot[1] = fma(hy, om[2], fma(-hz, om[1], hw * om[0]));
  qdot[2] = fma(-hx, om[2], fma(hw, om[1], hz * om[0]));
  qdot[3] = fma(hw, om[2], fma(hx, om[1], (-hy) * om[0]));
}

// [cos(att + pi/2), sin(att + pi/2)] = [-sin att, cos att]  (quadrotor2d.py:88)
RMV_HD void attitude_2d(rf theta, rf (&axis)[2]) {
  float s, c;
  sincos_pinned(theta.v, s, c);
  axis[0] = rf(-s);
  axis[1] = rf(c);
}

RMV_HD rf commanded_thrust(const DevParams& P, float a0) {
  rf T = rf(P.thrust_scale) * rf(a0);               // quadrotor2d.py:75 (scale is 1 elsewhere)
  if (P.thrust_clamp && T.v < 0.0f) T = rf(0.0f);    // quadrotor2d.py:76-77
  return T;
}

// free quadrotor: acc = T/m * axis + g ; advance (quadrotor3d.py:96-99, quadrotor2d.py:88-90)
template <int D>
RMV_HD void quad_free(const DevParams& P, rf T, const rf (&axis)[D], rf (&pos)[D], rf (&vel)[D]) {
  const rf tm = T * rf(P.inv_mass);   // T/m with m = 1: exact; general m: reciprocal-multiply
  rf acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = fma(tm, axis[i], rf(i == D - 1 ? P.neg_gravity : 0.0f));
  advance<D, false>(pos, vel, acc, rf(P.dt));
}

// quadrotor + tethered point load (quadrotor3d_slungload.py:101-145, quadrotor2d_slungload.py:93-132)
template <int D, bool VEL_FIRST>
RMV_HD void quad_slung(const DevParams& P, rf T, const rf (&axis)[D], rf (&pos)[D], rf (&vel)[D],
                       rf (&lp)[D], rf (&lv)[D]) {
  const rf dt(P.dt), m(P.mass), ml(P.load_mass), L(P.tether_len);
  rf g[D];
#pragma unroll
  for (int i = 0; i < D; ++i) g[i] = rf(i == D - 1 ? P.neg_gravity : 0.0f);
  rf tv[D];
#pragma unroll
  for (int i = 0; i < D; ++i) tv[i] = lp[i] - pos[i];
  // Switching predicate |tether| >= L (quadrotor3d_slungload.py:104).  After a taut step the
  // projection (:126-127) leaves |tether| = L(1 +- ulp), so the branch rides on rounding noise;
  // it is evaluated in float64 from the float32 state so that, given the same state, this path
  // takes the float64 reference's branch (DESIGN.md "Switching surface").
  rd acc2;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const rd d = rd((double)lp[i].v) - rd((double)pos[i].v);
    acc2 = (i == 0) ? d * d : acc2 + d * d;
  }
  const bool taut = sqrt(acc2).v >= (double)P.tether_len;
  const rf tm = T * rf(P.inv_mass);
  if (taut) {
    const rf inv_nt = rcp(norm<D>(tv));
    rf u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) u[i] = tv[i] * inv_nt;
    const rf lv2 = sumsq<D>(lv);
    const rf c = m * L * lv2;  // SCALAR, broadcast-subtracted from every component (:110)
    rf inner = u[0] * (T * axis[0] - c);
#pragma unroll
    for (int i = 1; i < D; ++i) inner = fma(u[i], T * axis[i] - c, inner);
    const rf inv(P.inv_total_mass);
    rf la[D], back[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      la[i] = fma(inv, inner * u[i], g[i]);
      back[i] = (-g[i]) + la[i];  // the (x + g) - g rounding of :111,115
    }
    advance<D, VEL_FIRST>(lp, lv, la, dt);
    const rf tn = ml * norm<D>(back);
    rf acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = fma(tn * u[i], rf(P.inv_mass), fma(tm, axis[i], g[i]));
    advance<D, VEL_FIRST>(pos, vel, acc, dt);
    // kinematic constraint (:126-128)
    rf dv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dv[i] = lp[i] - pos[i];
    const rf inv_nd = rcp(norm<D>(dv));
    rf dir[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dir[i] = dv[i] * inv_nd;
    rf rel = (lv[0] - vel[0]) * dir[0];
#pragma unroll
    for (int i = 1; i < D; ++i) rel = fma(lv[i] - vel[i], dir[i], rel);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      lp[i] = fma(dir[i], L, pos[i]);
      lv[i] = fma(-rel, dir[i], lv[i]);
    }
  } else {
    advance<D, VEL_FIRST>(lp, lv, g, dt);
    rf acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = fma(tm, axis[i], g[i]);
    advance<D, VEL_FIRST>(pos, vel, acc, dt);
  }
}

// ------------------------------------------------------------------------------------------
// one env, one step.  s: state components in the reference's tuple order; a: action.
// Outputs what the reward/done block needs (quadrotor3d.py:106-113 and twins): the SQUARED norms
// that `done` thresholds (compared against DevParams::*_sq_max, equivalent to norm > threshold)
// and the norm the reward negates.
// ------------------------------------------------------------------------------------------
template <int KIND> struct EnvStep;

template <> struct EnvStep<QUAD3D> {  // quadrotor3d.py:81-113
  static RMV_HD void run(const DevParams& P, float (&s)[10], const float (&a)[4], float& n_pos,
                         float& n_vel, float& n_rew) {
    rf pos[3] = {s[0], s[1], s[2]}, q[4] = {s[3], s[4], s[5], s[6]}, vel[3] = {s[7], s[8], s[9]};
    const rf om[3] = {a[1], a[2], a[3]};
    rf e3[3], qdot[4];
    attitude_3d(q, om, e3, qdot);
    quad_free<3>(P, commanded_thrust(P, a[0]), e3, pos, vel);
    const rf dt(P.dt);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[3 + i] = fma(qdot[i], dt, q[i]).v;
#pragma unroll
    for (int i = 0; i < 3; ++i) { s[i] = pos[i].v; s[7 + i] = vel[i].v; }
    const rf p2 = sumsq<3>(pos);
    n_pos = p2.v;
    n_vel = sumsq<3>(vel).v;
    n_rew = sqrt(p2).v;
  }
};

template <> struct EnvStep<QUAD3D_SLUNG> {  // quadrotor3d_slungload.py:87-156
  static RMV_HD void run(const DevParams& P, float (&s)[16], const float (&a)[4], float& n_pos,
                         float& n_vel, float& n_rew) {
    rf pos[3] = {s[0], s[1], s[2]}, q[4] = {s[3], s[4], s[5], s[6]}, vel[3] = {s[7], s[8], s[9]};
    rf lp[3] = {s[10], s[11], s[12]}, lv[3] = {s[13], s[14], s[15]};
    const rf om[3] = {a[1], a[2], a[3]};
    rf e3[3], qdot[4];
    attitude_3d(q, om, e3, qdot);
    quad_slung<3, false>(P, commanded_thrust(P, a[0]), e3, pos, vel, lp, lv);
    const rf dt(P.dt);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[3 + i] = fma(qdot[i], dt, q[i]).v;
#pragma unroll
    for (int i = 0; i < 3; ++i) { s[i] = pos[i].v; s[7 + i] = vel[i].v; s[10 + i] = lp[i].v; s[13 + i] = lv[i].v; }
    const rf l2 = sumsq<3>(lp);
    n_pos = l2.v;            // done: LOAD position, QUAD velocity (:149-152)
    n_vel = sumsq<3>(vel).v;
    n_rew = sqrt(l2).v;      // reward: load position (:156)
  }
};

template <> struct EnvStep<QUAD2D> {  // quadrotor2d.py:74-102
  static RMV_HD void run(const DevParams& P, float (&s)[5], const float (&a)[2], float& n_pos,
                         float& n_vel, float& n_rew) {
    rf pos[2] = {s[0], s[1]}, vel[2] = {s[3], s[4]};
    const rf th(s[2]);
    rf axis[2];
    attitude_2d(th, axis);
    quad_free<2>(P, commanded_thrust(P, a[0]), axis, pos, vel);
    s[2] = fma(rf(a[1]), rf(P.dt), th).v;
    s[0] = pos[0].v; s[1] = pos[1].v; s[3] = vel[0].v; s[4] = vel[1].v;
    const rf p2 = sumsq<2>(pos);
    n_pos = p2.v;
    n_vel = sumsq<2>(vel).v;
    n_rew = sqrt(p2).v;
  }
};

template <> struct EnvStep<QUAD2D_SLUNG> {  // quadrotor2d_slungload.py:79-143
  static RMV_HD void run(const DevParams& P, float (&s)[9], const float (&a)[2], float& n_pos,
                         float& n_vel, float& n_rew) {
    rf pos[2] = {s[0], s[1]}, vel[2] = {s[3], s[4]}, lp[2] = {s[5], s[6]}, lv[2] = {s[7], s[8]};
    const rf th(s[2]);
    rf axis[2];
    attitude_2d(th, axis);
    quad_slung<2, true>(P, commanded_thrust(P, a[0]), axis, pos, vel, lp, lv);
    s[2] = fma(rf(a[1]), rf(P.dt), th).v;
    s[0] = pos[0].v; s[1] = pos[1].v; s[3] = vel[0].v; s[4] = vel[1].v;
    s[5] = lp[0].v; s[6] = lp[1].v; s[7] = lv[0].v; s[8] = lv[1].v;
    n_pos = sumsq<2>(lp).v;  // done: load position / load velocity (:136-139)
    n_vel = sumsq<2>(lv).v;
    n_rew = norm<2>(pos).v;  // reward: QUAD position (:143)
  }
};

// ------------------------------------------------------------------------------------------
// control(): the reference's geometric controllers, one env, pinned arithmetic.
//   2D kinds: quadrotor2d.py:115-138, quadrotor2d_slungload.py:156-183
//   3D kinds: quadrotor3d.py:126-180, quadrotor3d_slungload.py:169-226
// Every operation is individually rounded and in the reference's own left-to-right order (the only fused steps are
// the sum-of-squares inside the norms and attitude_3d, as in step()); oracle/batch.py control() in float32 is the same
// sequence, so the kernels are checked bit for bit (tests/test_host_twin.py on the CPU, tests/test_gpu_vec_env.py on
// the GPU) and within 1e-5 of the float64 reference.
// ------------------------------------------------------------------------------------------
struct CtrlParams {
  float ref_pos[3], ref_vel[3];   // self.ref_pos / self.ref_vel (quadrotor3d.py:55-56)
  float mass, gravity;
};

// atan / atan2 in float32 with a pinned sequence (mirror of oracle/atan32.py): octant reduction + degree-4 polynomial
// in x^2 (Cephes atanf coefficients), no fused multiply-add.  <= 2.8e-7 absolute on [-pi, pi].
RMV_HD float atan_pinned(float xin) {
  const bool neg = (bits_of(xin) >> 31) != 0u;
  const rf ax(fabsf(xin));
  const bool big = ax.v > 2.414213562373095f, mid = ax.v > 0.4142135623730950f;
  const rf y0(big ? 1.57079632679489661923f : (mid ? 0.78539816339744830962f : 0.0f));
  rf r = ax;
  if (big) r = -(rf(1.0f) / ax);
  else if (mid) r = (ax - rf(1.0f)) / (ax + rf(1.0f));
  const rf z = r * r;
  const rf p = (((rf(8.05374449538e-2f) * z - rf(1.38776856032e-1f)) * z + rf(1.99777106478e-1f)) * z - rf(3.33329491539e-1f)) * z * r + r;
  const rf y = y0 + p;
  return neg ? -y.v : y.v;
}
RMV_HD float atan2_pinned(float y, float x) {
  const float pi = 3.14159265358979323846f, pio2 = 1.57079632679489661923f;
  if (x != x || y != y) return float_of(0x7fc00000u);
  if (x == 0.0f) return y < 0.0f ? -pio2 : (y == 0.0f ? 0.0f : pio2);
  if (y == 0.0f) return x < 0.0f ? ((bits_of(y) >> 31) ? -pi : pi) : y;      // +-0 -> +-pi or +-0, as libm
  const rf z(atan_pinned((rf(y) / rf(x)).v));
  const rf w(x < 0.0f ? (y < 0.0f ? -pi : pi) : 0.0f);
  return (w + z).v;
}

// pyquaternion Quaternion(matrix=R) on m = R^T (rows xb, yb, zb): trace method, largest component first
RMV_HD void quat_from_rows(const rf (&m)[3][3], rf (&q)[4]) {
  const rf one(1.0f);
  rf t;
  if (m[2][2].v < 0.0f) {
    if (m[0][0].v > m[1][1].v) {
      t = one + m[0][0] - m[1][1] - m[2][2];
      q[0] = m[1][2] - m[2][1]; q[1] = t; q[2] = m[0][1] + m[1][0]; q[3] = m[2][0] + m[0][2];
    } else {
      t = one - m[0][0] + m[1][1] - m[2][2];
      q[0] = m[2][0] - m[0][2]; q[1] = m[0][1] + m[1][0]; q[2] = t; q[3] = m[1][2] + m[2][1];
    }
  } else {
    if (m[0][0].v < -m[1][1].v) {
      t = one - m[0][0] - m[1][1] + m[2][2];
      q[0] = m[0][1] - m[1][0]; q[1] = m[2][0] + m[0][2]; q[2] = m[1][2] + m[2][1]; q[3] = t;
    } else {
      t = one + m[0][0] + m[1][1] + m[2][2];
      q[0] = t; q[1] = m[1][2] - m[2][1]; q[2] = m[2][0] - m[0][2]; q[3] = m[0][1] - m[1][0];
    }
  }
  const rf sc = rf(0.5f) / sqrt(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = q[i] * sc;
}

template <int KIND>
RMV_HD void control_law(const CtrlParams& C, const float (&s)[Dims<KIND>::S], float (&out)[Dims<KIND>::A]) {
  constexpr int D = Dims<KIND>::D;
  const rf kp(-5.0f), kv(-4.0f), g(C.gravity);
  if constexpr (D == 2) {
    const rf tau(0.1f);
    const rf pos[2] = {s[0], s[1]}, th(s[2]), vel[2] = {s[3], s[4]};
    rf acc[2];
    acc[0] = kp * (pos[0] - rf(C.ref_pos[0])) + kv * (vel[0] - rf(C.ref_vel[0])) + rf(0.0f);
    acc[1] = kp * (pos[1] - rf(C.ref_pos[1])) + kv * (vel[1] - rf(C.ref_vel[1])) + g;
    const rf att = rf(atan2_pinned(acc[1].v, acc[0].v)) - rf(1.57079632679489661923f);
    out[1] = ((rf(-1.0f) / tau) * (th - att)).v;
    out[0] = (rf(C.mass) * norm<2>(acc)).v;
  } else {
    const rf tau(0.3f);
    const rf pos[3] = {s[0], s[1], s[2]}, q[4] = {s[3], s[4], s[5], s[6]}, vel[3] = {s[7], s[8], s[9]};
    rf acc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      acc[i] = kp * (pos[i] - rf(C.ref_pos[i])) + kv * (vel[i] - rf(C.ref_vel[i])) - (i == 2 ? -g : rf(0.0f));
    const rf na = norm<3>(acc);
    rf zb[3] = {acc[0] / na, acc[1] / na, acc[2] / na};
    rf xb[3] = {zb[2], rf(0.0f) * zb[0], -zb[0]};                  // (0,1,0) x zb
    const rf nx = norm<3>(xb);
#pragma unroll
    for (int i = 0; i < 3; ++i) xb[i] = xb[i] / nx;
    const rf yb[3] = {zb[1] * xb[2] - zb[2] * xb[1], zb[2] * xb[0] - zb[0] * xb[2], zb[0] * xb[1] - zb[1] * xb[0]};
    const rf nz = norm<3>(zb);
#pragma unroll
    for (int i = 0; i < 3; ++i) zb[i] = zb[i] / nz;
    const rf m[3][3] = {{xb[0], xb[1], xb[2]}, {yb[0], yb[1], yb[2]}, {zb[0], zb[1], zb[2]}};
    rf qd[4];
    quat_from_rows(m, qd);
    // error quaternion = conj(q) * qd with the RAW state quaternion (no normalisation, as the reference)
    const rf w0 = q[0], x0 = -q[1], y0 = -q[2], z0 = -q[3];
    const rf ew = w0 * qd[0] - x0 * qd[1] - y0 * qd[2] - z0 * qd[3];
    const rf ex = x0 * qd[0] + w0 * qd[1] - z0 * qd[2] + y0 * qd[3];
    const rf ey = y0 * qd[0] + z0 * qd[1] + w0 * qd[2] - x0 * qd[3];
    const rf ez = z0 * qd[0] - y0 * qd[1] + x0 * qd[2] + w0 * qd[3];
    const rf sgn(ew.v > 0.0f ? 1.0f : (ew.v < 0.0f ? -1.0f : (ew.v == 0.0f ? 0.0f : ew.v)));   // np.sign (NaN stays NaN)
    const rf gain = (rf(2.0f) / tau) * sgn;
    // thrust = desired_acc . R(q^) e3
    const rf om[3] = {0.0f, 0.0f, 0.0f};
    rf e3[3], qdot[4];
    attitude_3d(q, om, e3, qdot);
    out[0] = (acc[0] * e3[0] + acc[1] * e3[1] + acc[2] * e3[2]).v;
    out[1] = (gain * ex).v; out[2] = (gain * ey).v; out[3] = (gain * ez).v;
  }
}

// done / reward block shared by the four envs (quadrotor3d.py:106-122).  Strict '>' so that a
// NaN norm is "not done", as in the reference; `ever_done` is steps_beyond_done is not None.
RMV_HD bool is_done(const DevParams& P, float pos_sq, float vel_sq) {
  return (pos_sq > P.pos_sq_max) || (vel_sq > P.vel_sq_max);
}
RMV_HD float reward_of(const DevParams& P, bool done, bool ever_done, float n_rew) {
  return done ? (ever_done ? P.later_done_reward : P.first_done_reward) : -n_rew;
}
// what step() hands back: baselines' RewardScaler sits OUTSIDE Monitor (make_vec_env, run.py:89), so the episode
// return keeps the raw reward and only the returned value is scaled.  One float32 product; exact for scale = 1.
RMV_HD float scaled_reward(const DevParams& P, float r) { return (rf(r) * rf(P.reward_scale)).v; }

}  // namespace rmv
