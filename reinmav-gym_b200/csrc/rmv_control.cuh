// rmv_control.cuh — batched geometric controllers: the reference's control()
// methods, one thread per env, state tile in -> action row out.
//   2D kinds: quadrotor2d.py:115-138, quadrotor2d_slungload.py:156-183
//   3D kinds: quadrotor3d.py:126-180, quadrotor3d_slungload.py:169-226
// The arithmetic is control_law<KIND> (rmv_math.cuh): pinned operation order, pinned atan2 — checked bit for bit against
// the float32 model and within 1e-5 (relative to max(1,|a|)) of the float64 reference.
#pragma once
#include "rmv_kernels.cuh"

namespace rmv {

struct ControlArgs {
  const float* state;     // tiled [n_tiles][S][128]
  float* act;             // [n][A] or [A][act_ld]
  int64_t n, act_ld;
  int act_soa;
  int rows;               // state is [n][S] rows (RMV_STATE_ROWS handle)
  CtrlParams C;           // ref_pos / ref_vel / mass / gravity
};

template <int KIND>
__global__ void __launch_bounds__(128) k_control(const ControlArgs a) {
  constexpr int S = Dims<KIND>::S, A = Dims<KIND>::A;
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
  float out[A];
  control_law<KIND>(a.C, st, out);
#pragma unroll
  for (int c = 0; c < A; ++c) {
    if (a.act_soa) a.act[c * a.act_ld + e] = out[c]; else a.act[e * A + c] = out[c];
  }
}

}  // namespace rmv
