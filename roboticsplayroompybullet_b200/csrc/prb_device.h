// prb_device.h — fp32 device-side model layout and small vector algebra.
// Included by the CUDA kernels (nvcc, sm_100a) and by the CPU SIMT emulator used in tests.
#pragma once
#include <stdint.h>

#ifndef PRB_EMU
#include <cuda_runtime.h>
#define PRB_HD __host__ __device__ __forceinline__
#define PRB_D __device__ __forceinline__
#define PRB_DN __device__ __noinline__
#else
#define PRB_HD inline
#define PRB_D inline
#define PRB_DN inline
#endif

#define PRB_MAXD 12
#define PRB_MAXFREE 2
#define PRB_MAXSLIDE 3
#define PRB_MAXCOL 64
#define PRB_MAXPAIR 1024
#define PRB_MAXDEPTH 8
#define PRB_NPARAM 25
#define PRB_PI_F 3.14159265358979323846f

// indices into DevModel::params (mirror of include/prb_model.h: enum prb_param)
enum {
  P_DT = 0, P_GRAVITY_Z, P_ERP_JOINT, P_ERP_CONTACT, P_LINEAR_SLOP, P_IK_DAMPING, P_IK_THRESHOLD,
  P_ARM_FORCE, P_SPARSE_THRESH, P_RESET_Z_OFFSET, P_DEFAULT_MOTOR_IMPULSE, P_MOTOR_KP, P_MOTOR_KD,
  P_LIMIT_MAX_IMPULSE, P_GEAR_RATIO, P_GEAR_ERP, P_GEAR_MAX_IMPULSE, P_MAX_COORD_VEL,
  P_ACTION_HIGH_XYZ, P_ACTION_HIGH_GRIP, P_OBJ_RESET_DZ, P_ARM_LIN_DAMP, P_ARM_ANG_DAMP,
  P_CONTACT_BREAKING, P_ACTION_TYPE
};

struct DevModel {
  int env_kind, arm_kind, nd, n_ik, n_free, n_slide, n_col, n_pair, n_grip;
  int ik_calls, ik_iters, ik_reset_iters, n_substeps, solver_iters, settle_steps;
  int obs_dim, goal_dim, fps_dim, observation_dim, use_orientation, return_velocity, play;
  int grip_obs_dof, gear_a, gear_b;
  int nv;             // total velocity DoF = nd + 6 n_free + n_slide  (<= 32: one lane each)
  int state_dim;      // floats of simulation state per env
  int state_stride;   // padded to a multiple of 32 floats: one env = whole 128-byte lines
  // ---- arm
  int parent[PRB_MAXD], jtype[PRB_MAXD], depth[PRB_MAXD];
  unsigned char path[PRB_MAXD][PRB_MAXDEPTH];   // root -> link (inclusive)
  unsigned sub_mask[PRB_MAXD];                  // bit i set: link i in subtree(j) (incl. j)
  unsigned anc_mask[PRB_MAXD];                  // bit j set: j ancestor-or-self of link i
  float jpos[PRB_MAXD][3], jrot[PRB_MAXD][9], axis[PRB_MAXD][3], com[PRB_MAXD][3];
  float mass[PRB_MAXD], inertia[PRB_MAXD][6];   // xx xy xz yy yz zz about COM, link frame
  float lo[PRB_MAXD], hi[PRB_MAXD], jdamp[PRB_MAXD], rest[PRB_MAXD];
  float base_pos[3], base_rot[9], base_quat[4];
  int site_link[4];
  float site_pos[4][3], site_rot[4][9];
  int joints_obs_dof[8];
  // ---- colliders (boxes)
  signed char col_body[PRB_MAXCOL], col_link[PRB_MAXCOL], col_urdf[PRB_MAXCOL];
  unsigned char col_obj[PRB_MAXCOL];   // collision-object id (manifold reduction groups)
  float col_pos[PRB_MAXCOL][3], col_rot[PRB_MAXCOL][9], col_half[PRB_MAXCOL][3];
  float col_fric[PRB_MAXCOL], col_spin[PRB_MAXCOL], col_stiff[PRB_MAXCOL], col_damp[PRB_MAXCOL];
  unsigned char pair_a[PRB_MAXPAIR], pair_b[PRB_MAXPAIR];
  // ---- free bodies, slide bodies
  float free_mass[PRB_MAXFREE], free_inertia[PRB_MAXFREE][3], free_ld[PRB_MAXFREE], free_ad[PRB_MAXFREE];
  float free_pos0[PRB_MAXFREE][3], free_quat0[PRB_MAXFREE][4];
  int slide_jtype[PRB_MAXSLIDE];
  float slide_pos[PRB_MAXSLIDE][3], slide_rot[PRB_MAXSLIDE][9], slide_axis[PRB_MAXSLIDE][3];
  float slide_axis_w[PRB_MAXSLIDE][3], slide_minv[PRB_MAXSLIDE], slide_ad[PRB_MAXSLIDE];
  float slide_motor[PRB_MAXSLIDE][4];
  // ---- control / env constants
  float ctrl_ll[8], ctrl_ul[8], ctrl_inc[8];
  int grip_dof[8], grip_mimic[8];
  float grip_scale[8], grip_offset[8], grip_force[8];
  float goal_lo[3], goal_hi[3], obj_lo[3], obj_hi[3], env_hi[3], default_orn[4];
  float params[PRB_NPARAM];
};

// Output buffers: one [N, dim] fp32 array per key of the reference's observation dict
// (environments.py:849-861) plus reward / success / target_poses (environments.py:211-214).
struct DevOut {
  float *obs_quat, *achieved_goal, *desired_goal, *cag, *fps, *joints, *velocity, *observation;
  float *proprio, *reward, *success, *target_poses;
  unsigned long long* overflow;   // env steps (since creation) in which a contact had to be dropped (capacity)
  unsigned char* ovf_env;         // optional [N]: env dropped a contact since its last observation (counted once per env step)
  int* dbg;                       // optional [N,4] per-env-step maxima: A floats, contacts, pool floats, units
};

// ------------------------------------------------------------------ vector algebra
struct v3 { float x, y, z; };
struct m3 { float m[9]; };  // row-major

PRB_HD v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PRB_HD v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
PRB_HD v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
PRB_HD v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
PRB_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PRB_HD v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PRB_HD float norm(v3 a) { return sqrtf(dot(a, a)); }
PRB_HD float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
PRB_HD v3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
PRB_HD void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
PRB_HD m3 ldm(const float* p) { m3 r; for (int i = 0; i < 9; i++) r.m[i] = p[i]; return r; }
PRB_HD void stm(float* p, const m3& a) { for (int i = 0; i < 9; i++) p[i] = a.m[i]; }
PRB_HD v3 mul(const m3& A, v3 v) {
  return V3(A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
            A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z);
}
PRB_HD v3 tmul(const m3& A, v3 v) {  // A^T v
  return V3(A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
            A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z);
}
PRB_HD m3 mul(const m3& A, const m3& B) {
  m3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
PRB_HD v3 col(const m3& A, int j) { return V3(A.m[j], A.m[3 + j], A.m[6 + j]); }
PRB_HD m3 ident3() { m3 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.f : 0.f; return r; }
PRB_HD m3 axis_angle(v3 a, float q) {
  float s, c;
  sincosf(q, &s, &c);
  float t = 1.f - c;
  m3 R;
  R.m[0] = t * a.x * a.x + c; R.m[1] = t * a.x * a.y - s * a.z; R.m[2] = t * a.x * a.z + s * a.y;
  R.m[3] = t * a.x * a.y + s * a.z; R.m[4] = t * a.y * a.y + c; R.m[5] = t * a.y * a.z - s * a.x;
  R.m[6] = t * a.x * a.z - s * a.y; R.m[7] = t * a.y * a.z + s * a.x; R.m[8] = t * a.z * a.z + c;
  return R;
}
// symmetric 3x3 (xx xy xz yy yz zz) times vector
PRB_HD v3 symmul(const float* S, v3 v) {
  return V3(S[0] * v.x + S[1] * v.y + S[2] * v.z, S[1] * v.x + S[3] * v.y + S[4] * v.z, S[2] * v.x + S[4] * v.y + S[5] * v.z);
}
// R diag/sym R^T  -> symmetric 6
PRB_HD void rot_sym(const m3& R, const float* S, float* out) {
  // T = R*S
  float T[9];
  for (int i = 0; i < 3; i++) {
    float a = R.m[3 * i], b = R.m[3 * i + 1], c = R.m[3 * i + 2];
    T[3 * i] = a * S[0] + b * S[1] + c * S[2];
    T[3 * i + 1] = a * S[1] + b * S[3] + c * S[4];
    T[3 * i + 2] = a * S[2] + b * S[4] + c * S[5];
  }
  out[0] = T[0] * R.m[0] + T[1] * R.m[1] + T[2] * R.m[2];
  out[1] = T[0] * R.m[3] + T[1] * R.m[4] + T[2] * R.m[5];
  out[2] = T[0] * R.m[6] + T[1] * R.m[7] + T[2] * R.m[8];
  out[3] = T[3] * R.m[3] + T[4] * R.m[4] + T[5] * R.m[5];
  out[4] = T[3] * R.m[6] + T[4] * R.m[7] + T[5] * R.m[8];
  out[5] = T[6] * R.m[6] + T[7] * R.m[7] + T[8] * R.m[8];
}
// quaternions [x,y,z,w]
PRB_HD void quat_to_mat(const float* q, m3& R) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float s = 2.0f / (x * x + y * y + z * z + w * w);
  R.m[0] = 1 - s * (y * y + z * z); R.m[1] = s * (x * y - w * z); R.m[2] = s * (x * z + w * y);
  R.m[3] = s * (x * y + w * z); R.m[4] = 1 - s * (x * x + z * z); R.m[5] = s * (y * z - w * x);
  R.m[6] = s * (x * z - w * y); R.m[7] = s * (y * z + w * x); R.m[8] = 1 - s * (x * x + y * y);
}
PRB_HD void mat_to_quat(const m3& R, float* q) {
  float t = R.m[0] + R.m[4] + R.m[8];
  if (t > 0) {
    float s = sqrtf(t + 1.0f);
    q[3] = s * 0.5f; s = 0.5f / s;
    q[0] = (R.m[7] - R.m[5]) * s; q[1] = (R.m[2] - R.m[6]) * s; q[2] = (R.m[3] - R.m[1]) * s;
  } else {
    int i = R.m[0] < R.m[4] ? (R.m[4] < R.m[8] ? 2 : 1) : (R.m[0] < R.m[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    float s = sqrtf(R.m[4 * i] - R.m[4 * j] - R.m[4 * k] + 1.0f);
    float qq[4];
    qq[i] = s * 0.5f; s = 0.5f / s;
    qq[3] = (R.m[3 * k + j] - R.m[3 * j + k]) * s;
    qq[j] = (R.m[3 * j + i] + R.m[3 * i + j]) * s;
    qq[k] = (R.m[3 * k + i] + R.m[3 * i + k]) * s;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
  }
}
PRB_HD void quat_mul(const float* a, const float* b, float* o) {
  float x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  float y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  float z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  float w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
PRB_HD void quat_from_euler(const float* rpy, float* q) {  // reference: getQuaternionFromEuler, environments.py:960
  float sr, cr, sp, cp, sy, cy;
  sincosf(rpy[0] * 0.5f, &sr, &cr); sincosf(rpy[1] * 0.5f, &sp, &cp); sincosf(rpy[2] * 0.5f, &sy, &cy);
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
}
PRB_HD void euler_from_quat(const float* q, float* rpy) {  // reference: getEulerFromQuaternion, environments.py:859
  float sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], sqw = q[3] * q[3];
  float sarg = -2.f * (q[0] * q[2] - q[3] * q[1]);
  if (sarg <= -0.99999f) { rpy[0] = 0; rpy[1] = -0.5f * PRB_PI_F; rpy[2] = 2 * atan2f(q[0], -q[1]); }
  else if (sarg >= 0.99999f) { rpy[0] = 0; rpy[1] = 0.5f * PRB_PI_F; rpy[2] = 2 * atan2f(-q[0], q[1]); }
  else {
    rpy[0] = atan2f(2 * (q[1] * q[2] + q[3] * q[0]), sqw - sqx - sqy + sqz);
    rpy[1] = asinf(sarg);
    rpy[2] = atan2f(2 * (q[0] * q[1] + q[3] * q[2]), sqw + sqx - sqy - sqz);
  }
}
PRB_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// Philox4x32-10 counter RNG; same keying as the oracle so sampled resets are comparable
PRB_HD void philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
PRB_HD void rng4(uint64_t seed, uint32_t env, uint32_t attempt, uint32_t block, float* u) {
  uint32_t o[4];
  philox4(env, attempt, block, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  for (int i = 0; i < 4; i++) u[i] = (float)(o[i] >> 8) * (1.0f / 16777216.0f);
}
