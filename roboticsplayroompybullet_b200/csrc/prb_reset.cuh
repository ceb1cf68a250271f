// prb_reset.cuh — playEnv.reset() on the split step pipeline.
//
// Reference: environments.py:173-187 (reset until the sampled state is not already a success), :599-603
// (instance.reset), :519-556 (reset_object_pos: re-seat drawer / door / button / dial, drop the block at a
// uniform position, 100 settle substeps, again if it left the bounds), :575-596 (reset_arm: rest pose -> one
// IK call on the live arm -> hard reset of joints [0:6]), :492-516 (reset_goal_pos).
//
// A reset is a sequence of ROUNDS over the envs that are still pending:
//   prb_reset_place_kernel    thread per env: (re)seat the objects of the pending envs
//   masked step pipeline      `settle_steps` substeps of prb_setup_kernel + solver kernels, pending envs only
//   prb_reset_finish_kernel   warp per env: bounds check -> another try, else arm reset, goal, observation,
//                             reward -> another attempt while the state is already a success; counts the envs
//                             that stay pending
// The host reads that one counter after each round (reset is a synchronous call in the reference too); a full-batch
// reset of the play world needs one round for ~69 % of the envs and 8-9 shrinking rounds for the rest.  The finish
// kernel also writes the list of the envs that stay pending: the next round's step pipeline runs over that list
// (env_of() in prb_stream.cuh), so a round costs what its envs cost, not what the whole batch costs.
// Sampling is counter-based: (seed, global env id, attempt, draw), attempt = the env's lifetime reset counter,
// so results do not depend on sharding or on which envs are reset together.
#pragma once
#include "prb_stream.cuh"

#define RESET_MAX_ATTEMPTS 16      // guard of the "already successful" loop (the reference loops unbounded)
#define RESET_MAX_TRIES 4          // object placements per attempt (the reference recurses unbounded)

// ctl[2 e] = attempts made in this call | try index << 8;  ctl[2 e + 1] = RNG attempt id of the current attempt
__global__ void prb_reset_place_kernel(const DevModel* __restrict__ Mp, float* __restrict__ state, int* __restrict__ ctl,
                                       const unsigned char* __restrict__ mask, unsigned char* __restrict__ pending, int N,
                                       unsigned long long seed, unsigned env_offset, int first,
                                       int* __restrict__ list_out = nullptr, int* __restrict__ n_list = nullptr) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  const DevModel& M = *Mp;
  if (first) {
    pending[e] = (mask == nullptr || mask[e] != 0) ? 1 : 0;
    ctl[2 * e] = 0;
  }
  if (!pending[e]) return;
  if (list_out != nullptr) list_out[atomicAdd(n_list, 1)] = e;      // masked reset: the first round runs over a list too
  float* st = state + (size_t)e * M.state_stride;
  const int nd = M.nd;
  const int o_free = 5 * nd, o_slide = o_free + 13 * M.n_free, o_cnt = o_slide + 2 * M.n_slide + M.goal_dim + 8 + 1;
  const int t = (ctl[2 * e] >> 8) & 0xff;
  if (t == 0) {                                   // a new attempt draws a new RNG attempt id
    const float rc = st[o_cnt];
    ctl[2 * e + 1] = (int)(uint32_t)rc;
    st[o_cnt] = rc + 1.0f;
  }
  const uint32_t attempt = (uint32_t)ctl[2 * e + 1];
  if (M.play) {                                   // drawer back to its default pose, door / button / dial to 0
    float* d = st + o_free + 13;
    for (int k = 0; k < 3; k++) { d[k] = M.free_pos0[1][k]; d[7 + k] = 0.f; d[10 + k] = 0.f; }
    for (int k = 0; k < 4; k++) d[3 + k] = M.free_quat0[1][k];
    for (int s = 0; s < 2 * M.n_slide; s++) st[o_slide + s] = 0.f;
  }
  if (M.n_free > 0) {                             // the block: uniform in the object bounds, dropped from +dz
    float u[4];
    rng4(seed, env_offset + (uint32_t)e, attempt, (uint32_t)t, u);
    float* b = st + o_free;
    for (int k = 0; k < 3; k++) { b[k] = M.obj_lo[k] + (M.obj_hi[k] - M.obj_lo[k]) * u[k]; b[7 + k] = 0.f; b[10 + k] = 0.f; }
    b[2] += M.params[P_OBJ_RESET_DZ];
    b[3] = 0.f; b[4] = 0.f; b[5] = 0.7071f; b[6] = 0.7071f;
  }
}

template <int ND>
__global__ void __launch_bounds__(32 * SetupCfg::WPB) prb_reset_finish_kernel(const DevModel* __restrict__ Mp, float* __restrict__ state, DevOut O,
                                                                                int* __restrict__ ctl, unsigned char* __restrict__ pending,
                                                                                int* __restrict__ n_pending, int N, unsigned long long seed,
                                                                                unsigned env_offset, int* __restrict__ list_out = nullptr) {
  typedef SetupMemT<SetupCfg> WM;
  PRB_SMEM_DECL2;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * SetupCfg::WPB + wib;
  if (e >= N) return;
  if (!pending[e]) return;
  const DevModel& M = *Mp;
  WM& W = wm[wib];
  float* st = state + (size_t)e * M.state_stride;
  load_state(M, W, st, lane);
  const int c0 = ctl[2 * e];
  const int made = c0 & 0xff, t = (c0 >> 8) & 0xff;
  const uint32_t attempt = (uint32_t)ctl[2 * e + 1];
  __syncwarp();
  // the settle substeps may have dropped contacts: count the env once
  if (lane == 0 && O.ovf_env && O.ovf_env[e]) { O.ovf_env[e] = 0; if (O.overflow) atomicAdd(O.overflow, 1ull); }
  bool oob = false;
  if (M.n_free > 0) for (int k = 0; k < 3; k++) if (W.fpos[0][k] > M.env_hi[k]) oob = true;
  if (oob && t + 1 < RESET_MAX_TRIES) {           // uniform: every lane reads the same shared values
    if (lane == 0) { ctl[2 * e] = made | ((t + 1) << 8); const int idx = atomicAdd(n_pending, 1); if (list_out) list_out[idx] = e; }
    return;                                       // stays pending: the next round re-seats the objects (draw t + 1)
  }
  float u[4];
  // reset_arm (environments.py:575-596)
  if (lane == 0) {
    rng4(seed, env_offset + (uint32_t)e, attempt, 4u, u);
    float np_[3];
    for (int k = 0; k < 3; k++) np_[k] = M.goal_lo[k] + (M.goal_hi[k] - M.goal_lo[k]) * u[k];
    np_[2] += M.params[P_RESET_Z_OFFSET];
    float qq[7];
    for (int i = 0; i < M.n_ik; i++) { qq[i] = M.rest[i]; W.q[i] = M.rest[i]; W.qd[i] = 0.f; }
    if (M.arm_kind == 1) { W.q[M.n_ik] = 0.f; W.qd[M.n_ik] = 0.f; }
    if (M.n_ik == 6) ik_world<6>(M, qq, np_, M.default_orn, 1, M.ik_reset_iters);
    else ik_world<7>(M, qq, np_, M.default_orn, 1, M.ik_reset_iters);
    for (int i = 0; i < 6; i++) { W.q[i] = qq[i]; W.qd[i] = 0.f; }
  }
  __syncwarp();
  // reset_goal_pos (environments.py:492-516)
  rng4(seed, env_offset + (uint32_t)e, attempt, 5u, u);
  if (!M.play) {
    if (lane < 3) W.goal[lane] = M.goal_lo[lane] + (M.goal_hi[lane] - M.goal_lo[lane]) * u[lane];
    __syncwarp();
  } else {
    phase_observe(M, W, lane, O, (size_t)e, true);   // writes achieved_goal for this env
    __syncwarp();
    int idx = (int)(u[0] * M.goal_dim);
    if (idx >= M.goal_dim) idx = M.goal_dim - 1;
    if (lane < M.goal_dim) {
      float g = O.achieved_goal[(size_t)e * M.goal_dim + lane];
      if (lane == idx) g = g + u[1];
      W.goal[lane] = g;
    }
    __syncwarp();
  }
  const float r = phase_observe(M, W, lane, O, (size_t)e, true);
  __syncwarp();
  store_state(M, W, st, lane);
  if (lane == 0) {
    if (r > -1.f && made + 1 < RESET_MAX_ATTEMPTS) { ctl[2 * e] = made + 1; const int idx = atomicAdd(n_pending, 1); if (list_out) list_out[idx] = e; }   // already a success: again
    else pending[e] = 0;
  }
}

// playEnv.reset(o) (environments.py:173-187 with an observation): objects re-seated from obs_quat without settle steps
// (:541-556), the arm from the rest pose through one IK call to the observed end-effector pose (:582-596), a new goal
// (:492-516), again while the state already satisfies it (only the goal draw changes between attempts).  The object is
// read at its real offset in the obs_quat layout (the reference's 11 / 10 indexing is wrong for the 19-D play layout);
// restore_env also restores drawer y / door / button / dial (the reference leaves them at their defaults): INTEGRATION.md.
template <int ND>
__global__ void __launch_bounds__(32 * SetupCfg::WPB) prb_reset_to_kernel(const DevModel* __restrict__ Mp, float* __restrict__ state, DevOut O,
                                                                            const float* __restrict__ obs, const unsigned char* __restrict__ mask,
                                                                            int N, unsigned long long seed, unsigned env_offset, int restore_env) {
  typedef SetupMemT<SetupCfg> WM;
  PRB_SMEM_DECL2;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * SetupCfg::WPB + wib;
  if (e >= N) return;
  if (mask != nullptr && mask[e] == 0) return;
  const DevModel& M = *Mp;
  WM& W = wm[wib];
  float* st = state + (size_t)e * M.state_stride;
  load_state(M, W, st, lane);
  const float* ob = obs + (size_t)e * M.obs_dim;
  const int rv = M.return_velocity, uo = M.use_orientation;
  const int o_obj = 3 + (rv ? 3 : 0) + (uo ? 4 : 0) + 1;
  if (lane == 0) {
    if (M.play) {
      for (int k = 0; k < 3; k++) { W.fpos[1][k] = M.free_pos0[1][k]; W.fvel[1][k] = 0.f; W.fang[1][k] = 0.f; }
      for (int k = 0; k < 4; k++) W.fquat[1][k] = M.free_quat0[1][k];
      for (int s = 0; s < M.n_slide; s++) { W.sq[s] = 0.f; W.sqd[s] = 0.f; }
    }
    if (M.n_free > 0) {
      for (int k = 0; k < 3; k++) { W.fpos[0][k] = ob[o_obj + k]; W.fvel[0][k] = 0.f; W.fang[0][k] = 0.f; }
      if (uo) for (int k = 0; k < 4; k++) W.fquat[0][k] = ob[o_obj + 3 + k];
      else { W.fquat[0][0] = 0.f; W.fquat[0][1] = 0.f; W.fquat[0][2] = 0.f; W.fquat[0][3] = 1.f; }
    }
    if (M.play && restore_env) {
      const int o_env = o_obj + 7;
      W.fpos[1][1] = ob[o_env];
      W.sq[0] = ob[o_env + 1]; W.sq[1] = ob[o_env + 2]; W.sq[2] = ob[o_env + 3] * 2.2f;
    }
    float tp[3] = {ob[0], ob[1], ob[2]}, tq[4];
    for (int k = 0; k < 4; k++) tq[k] = uo ? ob[(rv ? 6 : 3) + k] : M.default_orn[k];
    float qq[7];
    for (int i = 0; i < M.n_ik; i++) { qq[i] = M.rest[i]; W.q[i] = M.rest[i]; W.qd[i] = 0.f; }
    if (M.arm_kind == 1) { W.q[M.n_ik] = 0.f; W.qd[M.n_ik] = 0.f; }
    if (M.n_ik == 6) ik_world<6>(M, qq, tp, tq, 1, M.ik_reset_iters);
    else ik_world<7>(M, qq, tp, tq, 1, M.ik_reset_iters);
    for (int i = 0; i < 6; i++) { W.q[i] = qq[i]; W.qd[i] = 0.f; }
  }
  __syncwarp();
  float r = 0.f;
  for (int guard = 0; guard < RESET_MAX_ATTEMPTS && r > -1.f; guard++) {
    const uint32_t attempt = (uint32_t)W.reset_count;
    __syncwarp();
    if (lane == 0) W.reset_count += 1.0f;
    float u[4];
    rng4(seed, env_offset + (uint32_t)e, attempt, 5u, u);
    if (!M.play) {
      if (lane < 3) W.goal[lane] = M.goal_lo[lane] + (M.goal_hi[lane] - M.goal_lo[lane]) * u[lane];
      __syncwarp();
    } else {
      phase_observe(M, W, lane, O, (size_t)e, true);
      __syncwarp();
      int idx = (int)(u[0] * M.goal_dim);
      if (idx >= M.goal_dim) idx = M.goal_dim - 1;
      if (lane < M.goal_dim) {
        float g = O.achieved_goal[(size_t)e * M.goal_dim + lane];
        if (lane == idx) g = g + u[1];
        W.goal[lane] = g;
      }
      __syncwarp();
    }
    r = phase_observe(M, W, lane, O, (size_t)e, true);
    __syncwarp();
  }
  store_state(M, W, st, lane);
}

