// rmv_pair.cuh — quadrotor3d step() for TWO envs per thread on Blackwell's packed float32 pipe (sm_100a FFMA2 / FMUL2 /
// FADD2: `fma.rn.f32x2` and friends, one instruction = the same IEEE round-to-nearest operation on both halves of a
// 64-bit register pair).
//
// The step kernel is bound by instruction issue and latency, not by arithmetic throughput: ~190 issued instructions per
// env-step of which ~80 are FFMA / FMUL / FADD.  Two envs in one thread share every non-arithmetic instruction (address
// arithmetic, predicates, branches, uniform loads), turn 10 + 10 row accesses into 10 + 10 64-bit ones, and the packed
// pipe halves the arithmetic issue slots — at the same number of envs in flight per SM (32 warps x 64 envs instead of
// 64 x 32).  Each half goes through exactly the operation sequence of EnvStep<QUAD3D> (rmv_math.cuh), individually
// rounded, so the results are bit-identical to the one-env-per-thread kernels (tests/test_gpu_parity.py).
//
// Device-only (the host twin of the arithmetic is rmv_math.cuh; this file is checked against it bit for bit on the GPU).
#pragma once
#include "rmv_math.cuh"

#if defined(__CUDACC__)
namespace rmv {

struct p2 {
  float2 v;
  __device__ __forceinline__ p2() : v(make_float2(0.0f, 0.0f)) {}
  __device__ __forceinline__ p2(float2 x) : v(x) {}
  __device__ __forceinline__ p2(float lo, float hi) : v(make_float2(lo, hi)) {}
  __device__ __forceinline__ explicit p2(float both) : v(make_float2(both, both)) {}
};
__device__ __forceinline__ p2 operator+(p2 a, p2 b) { return p2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ p2 operator*(p2 a, p2 b) { return p2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ p2 operator-(p2 a) { return p2(-a.v.x, -a.v.y); }
__device__ __forceinline__ p2 operator-(p2 a, p2 b) { return p2(__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))); }   // a - b == a + (-b) exactly
__device__ __forceinline__ p2 fma(p2 a, p2 b, p2 c) { return p2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ p2 sqrt(p2 a) { return p2(__fsqrt_rn(a.v.x), __fsqrt_rn(a.v.y)); }

__device__ __forceinline__ p2 sumsq3(const p2 (&c)[3]) { return fma(c[2], c[2], fma(c[1], c[1], c[0] * c[0])); }

// EnvStep<QUAD3D>::run for two envs: s[c] = component c of (env 0, env 1), a[c] likewise.
// Out: squared position / velocity norms (done thresholds) and the position norm (reward), per env.
__device__ __forceinline__ void quad3d_pair_step(const DevParams& P, p2 (&s)[10], const p2 (&a)[4], p2& n_pos, p2& n_vel, p2& n_rew) {
  p2 pos[3] = {s[0], s[1], s[2]}, vel[3] = {s[7], s[8], s[9]};
  const p2 q[4] = {s[3], s[4], s[5], s[6]};
  // attitude_3d: normalise (guarded per env), R(q^) e_z, (0.5 q^) (x) (0, om)
  p2 w = q[0], x = q[1], y = q[2], z = q[3];
  const p2 n2 = fma(z, z, fma(y, y, fma(x, x, w * w)));
  const p2 n = sqrt(n2);
  // `if (n > 0) q *= 1/n`: a factor of exactly 1 leaves the env whose guard fails untouched (x * 1.0f == x for every x)
  const p2 inv(n.v.x > 0.0f ? __frcp_rn(n.v.x) : 1.0f, n.v.y > 0.0f ? __frcp_rn(n.v.y) : 1.0f);
  w = w * inv; x = x * inv; y = y * inv; z = z * inv;
  p2 e3[3];
  e3[0] = fma(y, w, fma(z, x, fma(w, y, x * z)));
  e3[1] = fma(-x, w, fma(-w, x, fma(z, y, y * z)));
  e3[2] = fma(w, w, fma(-x, x, fma(-y, y, z * z)));
  const p2 h(0.5f);
  const p2 hw = h * w, hx = h * x, hy = h * y, hz = h * z;
  const p2 om[3] = {a[1], a[2], a[3]};
  p2 qdot[4];
  qdot[0] = fma(-hz, om[2], fma(-hy, om[1], (-hx) * om[0]));
  qdot[1] = fma(hy, om[2], fma(-hz, om[1], hw * om[0]));
  qdot[2] = fma(-hx, om[2], fma(hw, om[1], hz * om[0]));
  qdot[3] = fma(hw, om[2], fma(hx, om[1], (-hy) * om[0]));
  // commanded_thrust + quad_free<3>
  p2 T = p2(P.thrust_scale) * a[0];
  if (P.thrust_clamp) T = p2(T.v.x < 0.0f ? 0.0f : T.v.x, T.v.y < 0.0f ? 0.0f : T.v.y);
  const p2 tm = T * p2(P.inv_mass);
  const p2 dt(P.dt);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const p2 acc = fma(tm, e3[i], p2(i == 2 ? P.neg_gravity : 0.0f));
    pos[i] = fma(h * acc * dt, dt, fma(vel[i], dt, pos[i]));       // position with the OLD velocity (quadrotor3d.py:98-99)
    vel[i] = fma(acc, dt, vel[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) s[3 + i] = fma(qdot[i], dt, q[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { s[i] = pos[i]; s[7 + i] = vel[i]; }
  n_pos = sumsq3(pos);
  n_vel = sumsq3(vel);
  n_rew = sqrt(n_pos);
}

}  // namespace rmv
#endif
