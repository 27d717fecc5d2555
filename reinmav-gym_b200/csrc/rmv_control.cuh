// rmv_control.cuh — batched geometric controllers: the reference's control()
// methods, one thread per env, state tile in -> action row out.
//   2D kinds: quadrotor2d.py:115-138, quadrotor2d_slungload.py:156-183
//   3D kinds: quadrotor3d.py:126-180, quadrotor3d_slungload.py:169-226
// Parity bar: within 1e-5 (relative to max(1,|a|)) of the float64 oracle — atan2f/sqrtf are the CUDA
// library versions here, so this path is tolerance-checked, not bit-pinned like step().
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

struct ControlArgs {
  const float* state;     // tiled [n_tiles][S][128]
  float* act;             // [n][A] or [A][act_ld]
  int64_t n, act_ld;
  int act_soa;
  int rows;               // state is [n][S] rows (RMV_STATE_ROWS handle)
  float ref_pos[3], ref_vel[3];
  float mass, gravity;
};

// pyquaternion Quaternion(matrix=R) on m = R^T (rows xb, yb, zb): trace method, largest component first
__device__ __forceinline__ void quat_from_rows(const float (&m)[3][3], float (&q)[4]) {
  float t;
  if (m[2][2] < 0.0f) {
    if (m[0][0] > m[1][1]) {
      t = 1.0f + m[0][0] - m[1][1] - m[2][2];
      q[0] = m[1][2] - m[2][1]; q[1] = t; q[2] = m[0][1] + m[1][0]; q[3] = m[2][0] + m[0][2];
    } else {
      t = 1.0f - m[0][0] + m[1][1] - m[2][2];
      q[0] = m[2][0] - m[0][2]; q[1] = m[0][1] + m[1][0]; q[2] = t; q[3] = m[1][2] + m[2][1];
    }
  } else {
    if (m[0][0] < -m[1][1]) {
      t = 1.0f - m[0][0] - m[1][1] + m[2][2];
      q[0] = m[0][1] - m[1][0]; q[1] = m[2][0] + m[0][2]; q[2] = m[1][2] + m[2][1]; q[3] = t;
    } else {
      t = 1.0f + m[0][0] + m[1][1] + m[2][2];
      q[0] = t; q[1] = m[1][2] - m[2][1]; q[2] = m[2][0] - m[0][2]; q[3] = m[0][1] - m[1][0];
    }
  }
  const float sc = 0.5f / sqrtf(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] *= sc;
}

template <int KIND>
__global__ void __launch_bounds__(128) k_control(const ControlArgs a) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A, D = Dims<KIND>::D;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  float st[S];
  if (a.rows) {
#pragma unroll
    for (int c = 0; c < S; ++c) st[c] = a.state[e * S + c];
  } else {
    const float* const tile = a.state + ((e >> 7) * (int64_t)(S * kTile) + (e & (kTile - 1)));
#pragma unroll
    for (int c = 0; c < S; ++c) st[c] = tile[c * kTile];
  }
  const float* const row = st;
  float out[A];
  if constexpr (D == 2) {
    const float kp = -5.0f, kv = -4.0f, tau = 0.1f;
    const float px = row[0], py = row[1], th = row[2], vx = row[3], vy = row[4];
    const float ax = kp * (px - a.ref_pos[0]) + kv * (vx - a.ref_vel[0]) + 0.0f;
    const float ay = kp * (py - a.ref_pos[1]) + kv * (vy - a.ref_vel[1]) + a.gravity;
    const float att = atan2f(ay, ax) - 1.57079632679489661923f;
    out[1] = (-1.0f / tau) * (th - att);
    out[0] = a.mass * sqrtf(ax * ax + ay * ay);
  } else {
    const float kp = -5.0f, kv = -4.0f, tau = 0.3f;
    float p[3], q[4], v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { p[i] = row[i]; v[i] = row[7 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = row[3 + i];
    float acc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = kp * (p[i] - a.ref_pos[i]) + kv * (v[i] - a.ref_vel[i]) + (i == 2 ? a.gravity : 0.0f);
    const float na = sqrtf(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]);
    float zb[3] = {acc[0] / na, acc[1] / na, acc[2] / na};
    float xb[3] = {zb[2], 0.0f, -zb[0]};                       // (0,1,0) x zb
    const float nx = sqrtf(xb[0] * xb[0] + xb[1] * xb[1] + xb[2] * xb[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) xb[i] /= nx;
    const float yb[3] = {zb[1] * xb[2] - zb[2] * xb[1], zb[2] * xb[0] - zb[0] * xb[2], zb[0] * xb[1] - zb[1] * xb[0]};
    const float nz = sqrtf(zb[0] * zb[0] + zb[1] * zb[1] + zb[2] * zb[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) zb[i] /= nz;
    const float m[3][3] = {{xb[0], xb[1], xb[2]}, {yb[0], yb[1], yb[2]}, {zb[0], zb[1], zb[2]}};
    float qd[4];
    quat_from_rows(m, qd);
    // error quaternion = conj(q) * qd with the RAW state quaternion (no normalisation, as the reference)
    const float w0 = q[0], x0 = -q[1], y0 = -q[2], z0 = -q[3];
    const float ew = w0 * qd[0] - x0 * qd[1] - y0 * qd[2] - z0 * qd[3];
    const float ex = x0 * qd[0] + w0 * qd[1] - z0 * qd[2] + y0 * qd[3];
    const float ey = y0 * qd[0] + z0 * qd[1] + w0 * qd[2] - x0 * qd[3];
    const float ez = z0 * qd[0] - y0 * qd[1] + x0 * qd[2] + w0 * qd[3];
    const float sgn = (ew > 0.0f) ? 1.0f : ((ew < 0.0f) ? -1.0f : (ew == 0.0f ? 0.0f : ew));   // np.sign (NaN stays NaN)
    const float gain = (2.0f / tau) * sgn;
    // thrust = desired_acc . R(q^) e3
    const rf qq[4] = {q[0], q[1], q[2], q[3]}, om[3] = {0.0f, 0.0f, 0.0f};
    rf e3[3], qdot[4];
    attitude_3d(qq, om, e3, qdot);
    out[0] = acc[0] * e3[0].v + acc[1] * e3[1].v + acc[2] * e3[2].v;
    out[1] = gain * ex; out[2] = gain * ey; out[3] = gain * ez;
  }
#pragma unroll
  for (int c = 0; c < A; ++c) {
    if (a.act_soa) a.act[c * a.act_ld + e] = out[c]; else a.act[e * A + c] = out[c];
  }
}

}  // namespace rmv
