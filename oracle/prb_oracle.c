/* prb_oracle.c — CPU restatement (fp64, scalar C) of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under roboticsplayroompybullet_b200/ may import,
 * link or execute this file; it is the checker for tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the arithmetic of the reference path lives in PyBullet/Bullet3
 * (third party, un-vendored, un-versioned: reference setup.py:5 lists only 'gym'),
 * which cannot be imported in the authoring container and ships no golden vectors
 * (the reference has no tests).  This file therefore restates (a) the reference's own
 * Python line by line and (b) the PUBLISHED Bullet3 algorithms it calls, from their
 * documented behaviour (Bullet3 2.8x/3.x: btMultiBodyDynamicsWorld, btMultiBody ABA,
 * btMultiBodyConstraintSolver PGS, btMultiBodyJointMotor/JointLimitConstraint,
 * btBoxBoxDetector, BussIK DLS as driven by PhysicsServerCommandProcessor).
 *
 * Reference call sites followed (roboticsPlayroomPybullet/envs/…):
 *   step            environments.py:206-214
 *   action -> IK    environments.py:915-934, 955-961, 984-1007; inverseKinematics.py:44-50
 *   motors          environments.py:1010-1034 (arm), 1037-1073 (gripper)
 *   stepSimulation  environments.py:485-490 (12 substeps), :326 (no residual early exit)
 *   observation     environments.py:720-894
 *   reward          environments.py:269-304, playRewardFunc.py:16-77
 *   reset           environments.py:173-187, 492-603
 * Independent of the CUDA product on purpose: articulated-body algorithm (Featherstone,
 * world coordinates) here versus composite-rigid-body + Cholesky in the kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/prb_model.h"

#define MAXD 12
#define MAXFREE 2
#define MAXSLIDE 3
#define MAXV (MAXD + 6 * MAXFREE + MAXSLIDE)
#define MAXCOL 64
#define MAXCONTACT 96
#define MAXROW (2 * MAXD + MAXD + MAXSLIDE + 1 + 4 * MAXCONTACT)
#define PI 3.14159265358979323846

/* ORC_REAL=float builds the SAME restatement in single precision (liborc f32): used only by the conditioning tests,
 * which bound |CUDA - oracle| on stiff contact states by |oracle(fp32) - oracle(fp64)| of this independent algorithm. */
#ifndef ORC_REAL
#define ORC_REAL double
#endif
typedef ORC_REAL real;
typedef struct { real x, y, z; } v3;
typedef struct { real m[3][3]; } m3;

/* ---------------------------------------------------------------- small algebra */
static v3 V(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vscale(v3 a, real s) { return V(a.x * s, a.y * s, a.z * s); }
static real vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static real vnorm(v3 a) { return sqrt(vdot(a, a)); }
static real vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static v3 mmulv(const m3* A, v3 v) {
  return V(A->m[0][0] * v.x + A->m[0][1] * v.y + A->m[0][2] * v.z,
           A->m[1][0] * v.x + A->m[1][1] * v.y + A->m[1][2] * v.z,
           A->m[2][0] * v.x + A->m[2][1] * v.y + A->m[2][2] * v.z);
}
static v3 mtmulv(const m3* A, v3 v) {
  return V(A->m[0][0] * v.x + A->m[1][0] * v.y + A->m[2][0] * v.z,
           A->m[0][1] * v.x + A->m[1][1] * v.y + A->m[2][1] * v.z,
           A->m[0][2] * v.x + A->m[1][2] * v.y + A->m[2][2] * v.z);
}
static m3 mmul(const m3* A, const m3* B) {
  m3 C;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real s = 0; for (int k = 0; k < 3; k++) s += A->m[i][k] * B->m[k][j];
    C.m[i][j] = s;
  }
  return C;
}
static m3 mtrans(const m3* A) { m3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.m[i][j] = A->m[j][i]; return C; }
static m3 mload(const double* p) { m3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.m[i][j] = p[3 * i + j]; return C; }
static v3 vload(const double* p) { return V(p[0], p[1], p[2]); }
static m3 mloadr(const real* p) { m3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.m[i][j] = p[3 * i + j]; return C; }
static v3 vloadr(const real* p) { return V(p[0], p[1], p[2]); }
static v3 mcol(const m3* A, int j) { return V(A->m[0][j], A->m[1][j], A->m[2][j]); }
static m3 axis_angle(v3 a, real q) {
  real c = cos(q), s = sin(q), t = 1 - c;
  m3 R = {{{t * a.x * a.x + c, t * a.x * a.y - s * a.z, t * a.x * a.z + s * a.y},
           {t * a.x * a.y + s * a.z, t * a.y * a.y + c, t * a.y * a.z - s * a.x},
           {t * a.x * a.z - s * a.y, t * a.y * a.z + s * a.x, t * a.z * a.z + c}}};
  return R;
}
/* quaternions are [x,y,z,w] (Bullet) */
static void quat_to_mat(const real* q, m3* R) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  real n = x * x + y * y + z * z + w * w, s = 2.0 / n;
  R->m[0][0] = 1 - s * (y * y + z * z); R->m[0][1] = s * (x * y - w * z); R->m[0][2] = s * (x * z + w * y);
  R->m[1][0] = s * (x * y + w * z); R->m[1][1] = 1 - s * (x * x + z * z); R->m[1][2] = s * (y * z - w * x);
  R->m[2][0] = s * (x * z - w * y); R->m[2][1] = s * (y * z + w * x); R->m[2][2] = 1 - s * (x * x + y * y);
}
static void mat_to_quat(const m3* R, real* q) { /* btMatrix3x3::getRotation */
  real t = R->m[0][0] + R->m[1][1] + R->m[2][2];
  if (t > 0) {
    real s = sqrt(t + 1.0);
    q[3] = s * 0.5; s = 0.5 / s;
    q[0] = (R->m[2][1] - R->m[1][2]) * s; q[1] = (R->m[0][2] - R->m[2][0]) * s; q[2] = (R->m[1][0] - R->m[0][1]) * s;
  } else {
    int i = R->m[0][0] < R->m[1][1] ? (R->m[1][1] < R->m[2][2] ? 2 : 1) : (R->m[0][0] < R->m[2][2] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    real s = sqrt(R->m[i][i] - R->m[j][j] - R->m[k][k] + 1.0);
    q[i] = s * 0.5; s = 0.5 / s;
    q[3] = (R->m[k][j] - R->m[j][k]) * s; q[j] = (R->m[j][i] + R->m[i][j]) * s; q[k] = (R->m[k][i] + R->m[i][k]) * s;
  }
}
static void quat_mul(const real* a, const real* b, real* o) { /* o = a*b */
  real x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  real y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  real z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  real w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
/* getQuaternionFromEuler (environments.py:960): R = Rz(yaw) Ry(pitch) Rx(roll) */
void orc_quat_from_euler(const real* rpy, real* q) {
  real hr = rpy[0] * 0.5, hp = rpy[1] * 0.5, hy = rpy[2] * 0.5;
  real cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
}
/* getEulerFromQuaternion (environments.py:859, playRewardFunc.py:25-26); no normalisation */
void orc_euler_from_quat(const real* q, real* rpy) {
  real sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], sqw = q[3] * q[3];
  real sarg = -2 * (q[0] * q[2] - q[3] * q[1]);
  if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * PI; rpy[2] = 2 * atan2(q[0], -q[1]); }
  else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * PI; rpy[2] = 2 * atan2(-q[0], q[1]); }
  else {
    rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), sqw - sqx - sqy + sqz);
    rpy[1] = asin(sarg);
    rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), sqw + sqx - sqy - sqz);
  }
}

/* ---------------------------------------------------------------- Philox4x32-10 */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* uniform in [0,1) with 24 random bits: exactly representable in fp32 and fp64 */
static void rng4(uint64_t seed, uint32_t env, uint32_t attempt, uint32_t block, real* u) {
  uint32_t o[4];
  philox(env, attempt, block, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  for (int i = 0; i < 4; i++) u[i] = (real)(o[i] >> 8) * (1.0 / 16777216.0);
}
void orc_rng4(uint64_t seed, uint32_t env, uint32_t attempt, uint32_t block, real* u) { rng4(seed, env, attempt, block, u); }

/* ---------------------------------------------------------------- state */
typedef struct {
  real q[MAXD], qd[MAXD], mtarget[MAXD], mkp[MAXD], mmaximp[MAXD];
  real fpos[MAXFREE][3], fquat[MAXFREE][4], fvel[MAXFREE][3], fang[MAXFREE][3];
  real sq[MAXSLIDE], sqd[MAXSLIDE];
  real goal[16];
  real lastq[8];
  real last_valid;
  real reset_count;
} State;

static int state_dim(const prb_model* M) { return 5 * M->nd + 13 * M->n_free + 2 * M->n_slide + M->goal_dim + 10; }
int orc_state_dim(const prb_model* M) { return state_dim(M); }

static void state_unpack(const prb_model* M, const real* s, State* S) {
  int nd = M->nd, o = 0;
  memset(S, 0, sizeof(*S));
  for (int i = 0; i < nd; i++) S->q[i] = s[o++];
  for (int i = 0; i < nd; i++) S->qd[i] = s[o++];
  for (int i = 0; i < nd; i++) S->mtarget[i] = s[o++];
  for (int i = 0; i < nd; i++) S->mkp[i] = s[o++];
  for (int i = 0; i < nd; i++) S->mmaximp[i] = s[o++];
  for (int b = 0; b < M->n_free; b++) {
    for (int k = 0; k < 3; k++) S->fpos[b][k] = s[o++];
    for (int k = 0; k < 4; k++) S->fquat[b][k] = s[o++];
    for (int k = 0; k < 3; k++) S->fvel[b][k] = s[o++];
    for (int k = 0; k < 3; k++) S->fang[b][k] = s[o++];
  }
  for (int b = 0; b < M->n_slide; b++) { S->sq[b] = s[o++]; S->sqd[b] = s[o++]; }
  for (int k = 0; k < M->goal_dim; k++) S->goal[k] = s[o++];
  for (int k = 0; k < 8; k++) S->lastq[k] = s[o++];
  S->last_valid = s[o++];
  S->reset_count = s[o++];
}
static void state_pack(const prb_model* M, const State* S, real* s) {
  int nd = M->nd, o = 0;
  for (int i = 0; i < nd; i++) s[o++] = S->q[i];
  for (int i = 0; i < nd; i++) s[o++] = S->qd[i];
  for (int i = 0; i < nd; i++) s[o++] = S->mtarget[i];
  for (int i = 0; i < nd; i++) s[o++] = S->mkp[i];
  for (int i = 0; i < nd; i++) s[o++] = S->mmaximp[i];
  for (int b = 0; b < M->n_free; b++) {
    for (int k = 0; k < 3; k++) s[o++] = S->fpos[b][k];
    for (int k = 0; k < 4; k++) s[o++] = S->fquat[b][k];
    for (int k = 0; k < 3; k++) s[o++] = S->fvel[b][k];
    for (int k = 0; k < 3; k++) s[o++] = S->fang[b][k];
  }
  for (int b = 0; b < M->n_slide; b++) { s[o++] = S->sq[b]; s[o++] = S->sqd[b]; }
  for (int k = 0; k < M->goal_dim; k++) s[o++] = S->goal[k];
  for (int k = 0; k < 8; k++) s[o++] = S->lastq[k];
  s[o++] = S->last_valid;
  s[o++] = S->reset_count;
}

/* ---------------------------------------------------------------- kinematics */
typedef struct {
  m3 R[MAXD]; v3 p[MAXD];      /* link frame (joint frame after motion) in world */
  v3 a[MAXD];                  /* joint axis, world */
  v3 c[MAXD];                  /* link COM, world */
} Kin;

/* forward kinematics of the reduced arm; base pose (bR,bp) */
static void arm_fk(const prb_model* M, const real* q, const m3* bR, v3 bp, Kin* K) {
  for (int i = 0; i < M->nd; i++) {
    int par = M->arm_parent[i];
    const m3* pR = par < 0 ? bR : &K->R[par];
    v3 pp = par < 0 ? bp : K->p[par];
    m3 jR = mload(M->arm_jrot + 9 * i);
    v3 jp = vload(M->arm_jpos + 3 * i), ax = vload(M->arm_axis + 3 * i);
    m3 R0 = mmul(pR, &jR);
    if (M->arm_jtype[i] == 0) {
      m3 Rq = axis_angle(ax, q[i]);
      K->R[i] = mmul(&R0, &Rq);
      K->p[i] = vadd(mmulv(pR, jp), pp);
    } else {
      K->R[i] = R0;
      K->p[i] = vadd(vadd(mmulv(pR, jp), pp), vscale(mmulv(&R0, ax), q[i]));
    }
    K->a[i] = mmulv(&K->R[i], ax);
    K->c[i] = vadd(K->p[i], mmulv(&K->R[i], vload(M->arm_com + 3 * i)));
  }
}
static int is_ancestor(const prb_model* M, int j, int link) { /* j ancestor-or-self of link */
  while (link >= 0) { if (link == j) return 1; link = M->arm_parent[link]; }
  return 0;
}
static void site_pose(const prb_model* M, const Kin* K, int site, v3* pos, m3* R) {
  int l = M->site_link[site];
  m3 sR = mload(M->site_rot + 9 * site);
  *pos = vadd(K->p[l], mmulv(&K->R[l], vload(M->site_pos + 3 * site)));
  *R = mmul(&K->R[l], &sR);
}
/* world-frame FK of all sites, for tests: out[site] = pos3 + quat4 */
void orc_fk_sites(const prb_model* M, const real* q, real* out) {
  Kin K; m3 bR = mload(M->arm_base_rot);
  arm_fk(M, q, &bR, vload(M->arm_base_pos), &K);
  for (int s = 0; s < 4; s++) {
    v3 p; m3 R; site_pose(M, &K, s, &p, &R);
    out[7 * s] = p.x; out[7 * s + 1] = p.y; out[7 * s + 2] = p.z;
    mat_to_quat(&R, out + 7 * s + 3);
  }
}

/* ---------------------------------------------------------------- inverse kinematics
 * One pybullet.calculateInverseKinematics(body, ee, pos, orn) call (call sites
 * inverseKinematics.py:48,50; environments.py:593,995-997): velocity DLS with
 * orientation, per-joint damping 0.5, <= max_iters iterations, exit when the
 * POSITION error drops below 1e-4 (checked before each iteration on the previous
 * iterate's error), angle step clamp 45 deg.  Everything in the base frame. */
static void solve_linear(int n, real* A, real* b) { /* Gaussian elimination, partial pivoting; A n x n row-major (stride MAXD) */
  for (int c = 0; c < n; c++) {
    int piv = c; real best = fabs(A[c * MAXD + c]);
    for (int r = c + 1; r < n; r++) if (fabs(A[r * MAXD + c]) > best) { best = fabs(A[r * MAXD + c]); piv = r; }
    if (piv != c) {
      for (int k = 0; k < n; k++) { real t = A[c * MAXD + k]; A[c * MAXD + k] = A[piv * MAXD + k]; A[piv * MAXD + k] = t; }
      real t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; r++) {
      real f = A[r * MAXD + c] / A[c * MAXD + c];
      for (int k = c; k < n; k++) A[r * MAXD + k] -= f * A[c * MAXD + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    real s = b[r];
    for (int k = r + 1; k < n; k++) s -= A[r * MAXD + k] * b[k];
    b[r] = s / A[r * MAXD + r];
  }
}
void orc_ik(const prb_model* M, const real* q_in, const real* tpos_w, const real* tquat_w, int max_iters, real* q_out) {
  int nd = M->nd;
  real lambda = M->params[PRB_P_IK_DAMPING], thresh = M->params[PRB_P_IK_THRESHOLD];
  m3 bR = mload(M->arm_base_rot); v3 bp = vload(M->arm_base_pos);
  /* target into base coordinates */
  v3 tp = mtmulv(&bR, vsub(vloadr(tpos_w), bp));
  real bq[4], bqi[4], tq[4];
  mat_to_quat(&bR, bq);
  bqi[0] = -bq[0]; bqi[1] = -bq[1]; bqi[2] = -bq[2]; bqi[3] = bq[3];
  quat_mul(bqi, tquat_w, tq);
  m3 I3 = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  real q[MAXD];
  for (int i = 0; i < nd; i++) q[i] = q_in[i];
  real diff = 1e30;
  int ee_link = M->site_link[0];
  for (int it = 0; it < max_iters && diff > thresh; it++) {
    Kin K; arm_fk(M, q, &I3, V(0, 0, 0), &K);
    v3 ep; m3 eR; site_pose(M, &K, 0, &ep, &eR);
    real eq[4]; mat_to_quat(&eR, eq);
    diff = vnorm(vsub(ep, tp));
    /* 6 x nd Jacobian of the EE frame origin (world/base axes) */
    real J[6][MAXD];
    for (int j = 0; j < nd; j++) {
      v3 jl = V(0, 0, 0), ja = V(0, 0, 0);
      if (is_ancestor(M, j, ee_link)) {
        if (M->arm_jtype[j] == 0) { ja = K.a[j]; jl = vcross(K.a[j], vsub(ep, K.p[j])); }
        else jl = K.a[j];
      }
      J[0][j] = jl.x; J[1][j] = jl.y; J[2][j] = jl.z; J[3][j] = ja.x; J[4][j] = ja.y; J[5][j] = ja.z;
    }
    real e[6];
    v3 dp = vsub(tp, ep);
    e[0] = dp.x; e[1] = dp.y; e[2] = dp.z;
    { /* deltaQ = endQ * startQ^-1 ; angle/axis; Bullet keeps the angle in a float */
      real si[4] = {-eq[0], -eq[1], -eq[2], eq[3]}, dq[4];
      real n2 = eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2] + eq[3] * eq[3];
      for (int k = 0; k < 4; k++) si[k] /= n2;
      quat_mul(tq, si, dq);
      real w = dq[3]; if (w > 1) w = 1; if (w < -1) w = -1;
      float angle = (float)(2.0 * acos(w));
      real s2 = 1.0 - dq[3] * dq[3];
      v3 axis;
      if (s2 < 10.0 * 2.220446049250313e-16) axis = V(1, 0, 0);
      else { real s = 1.0 / sqrt(s2); axis = V(dq[0] * s, dq[1] * s, dq[2] * s); }
      if (angle > PI) angle -= (float)(2.0 * PI); else if (angle < -PI) angle += (float)(2.0 * PI);
      real an = vnorm(axis);
      axis = vscale(axis, 1.0 / an);
      e[3] = angle * axis.x; e[4] = angle * axis.y; e[5] = angle * axis.z;
    }
    /* (J^T J + diag(lambda)) dtheta = J^T e   (Jacobian::CalcDeltaThetasDLS2) */
    real A[MAXD * MAXD], b[MAXD];
    for (int i = 0; i < nd; i++) {
      for (int j = 0; j < nd; j++) {
        real s = 0; for (int k = 0; k < 6; k++) s += J[k][i] * J[k][j];
        A[i * MAXD + j] = s + (i == j ? lambda : 0.0);
      }
      real s = 0; for (int k = 0; k < 6; k++) s += J[k][i] * e[k];
      b[i] = s;
    }
    solve_linear(nd, A, b);
    real mx = 0; for (int i = 0; i < nd; i++) if (fabs(b[i]) > mx) mx = fabs(b[i]);
    real maxang = 45.0 * PI / 180.0;
    if (mx > maxang) for (int i = 0; i < nd; i++) b[i] *= maxang / mx;
    for (int i = 0; i < nd; i++) q[i] += b[i];
  }
  for (int i = 0; i < nd; i++) q_out[i] = q[i];
}
/* InverseKinematicsSolver.calc_angles (inverseKinematics.py:44-50): the IK client's arm
 * has its first n_ik joints set to the current state, all other joints stay at 0;
 * `calls` chained solves, each restarted from the previous result's first n_ik entries. */
void orc_calc_angles(const prb_model* M, const real* q_cur, const real* tpos, const real* tquat, real* q_out) {
  real q[MAXD], r[MAXD];
  for (int i = 0; i < M->nd; i++) q[i] = i < M->n_ik ? q_cur[i] : 0.0;
  for (int c = 0; c < M->ik_calls; c++) {
    orc_ik(M, q, tpos, tquat, M->ik_iters, r);
    for (int i = 0; i < M->n_ik; i++) q[i] = r[i];
  }
  for (int i = 0; i < M->nd; i++) q_out[i] = r[i];
}

/* ---------------------------------------------------------------- collision: box-box
 * Separating-axis test over the 15 candidate axes with the classic face/edge bias
 * (an edge axis must beat the best face axis by 5 %), then either one edge-edge
 * point or the incident face clipped against the reference face, culled to <= 4
 * points (deepest first, rest spread by angle).  Mirrors btBoxBoxDetector's results:
 * normal points from B to A, points lie on B, depth >= 0. */
typedef struct { v3 pos; v3 n; real depth; } CPoint;

static int clip_quad(const real h[2], const real p_in[8], real ret[16]) { /* clip quad p (4 pts) by rectangle +-h */
  int nq = 4, nr = 0;
  real buffer[16];
  const real* q = p_in; real* r = ret;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      const real* pq = q; real* pr = r; nr = 0;
      for (int i = nq; i > 0; i--) {
        if (sign * pq[dir] < h[dir]) { pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++; if (nr & 8) { q = r; goto done; } }
        const real* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          pr[dir] = sign * h[dir];
          pr += 2; nr++; if (nr & 8) { q = r; goto done; }
        }
        pq += 2;
      }
      q = r; r = (q == ret) ? buffer : ret; nq = nr;
    }
  }
done:
  if (q != ret) memcpy(ret, q, nr * 2 * sizeof(real));
  return nr;
}
static void cull_points(int n, const real p[], int m, int i0, int iret[]) {
  real a, cx, cy, q;
  if (n == 1) { cx = p[0]; cy = p[1]; }
  else if (n == 2) { cx = 0.5 * (p[0] + p[2]); cy = 0.5 * (p[1] + p[3]); }
  else {
    a = 0; cx = 0; cy = 0;
    for (int i = 0; i < n - 1; i++) {
      q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
      a += q; cx += q * (p[i * 2] + p[i * 2 + 2]); cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
    }
    q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
    if (fabs(a + q) > 2.220446049250313e-16) a = 1.0 / (3.0 * (a + q)); else a = 1e18;
    cx = a * (cx + q * (p[n * 2 - 2] + p[0])); cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
  }
  real A[8]; int avail[8];
  for (int i = 0; i < n; i++) { A[i] = atan2(p[i * 2 + 1] - cy, p[i * 2] - cx); avail[i] = 1; }
  avail[i0] = 0; iret[0] = i0; iret++;
  for (int j = 1; j < m; j++) {
    a = j * (2 * PI / m) + A[i0];
    if (a > PI) a -= 2 * PI;
    real maxdiff = 1e9, diff; *iret = i0;
    for (int i = 0; i < n; i++) if (avail[i]) {
      diff = fabs(A[i] - a); if (diff > PI) diff = 2 * PI - diff;
      if (diff < maxdiff - 1e-5) { maxdiff = diff; *iret = i; }   /* ties: the first point wins */
    }
    avail[*iret] = 0; iret++;
  }
}
static void line_closest(v3 pa, v3 ua, v3 pb, v3 ub, real* alpha, real* beta) {
  v3 p = vsub(pb, pa);
  real uaub = vdot(ua, ub), q1 = vdot(ua, p), q2 = -vdot(ub, p), d = 1 - uaub * uaub;
  if (d <= 0.0001) { *alpha = 0; *beta = 0; }
  else { d = 1.0 / d; *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}
int orc_box_box_impl(v3 p1, const m3* R1, v3 side1h, v3 p2, const m3* R2, v3 side2h, CPoint* out) {
  /* TIE: a later axis replaces the current best only if it separates by 1 um more (structural ties of resting boxes are
   * otherwise decided by rounding, differently in fp32 and fp64: see the kernels' box_box) */
  const real fudge = 1.05, TIE = 1e-6;
  real A[3] = {side1h.x, side1h.y, side1h.z}, B[3] = {side2h.x, side2h.y, side2h.z};
  v3 p = vsub(p2, p1);
  v3 pp = mtmulv(R1, p);
  real Rm[3][3], Q[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rm[i][j] = vdot(mcol(R1, i), mcol(R2, j)); Q[i][j] = fabs(Rm[i][j]); }
  real s = -1e300, s2, l; int invert_normal = 0, code = 0;
  v3 normalC = V(0, 0, 0); const m3* normalRm = 0; int normalRcol = 0;
  real ppv[3] = {pp.x, pp.y, pp.z};
#define TST(expr1, expr2, Rmat, col, cc) \
  s2 = fabs(expr1) - (expr2); if (s2 > 0) return 0; \
  if (s2 > s + TIE) { s = s2; normalRm = Rmat; normalRcol = col; invert_normal = ((expr1) < 0); code = (cc); }
  TST(ppv[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), R1, 0, 1);
  TST(ppv[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), R1, 1, 2);
  TST(ppv[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), R1, 2, 3);
  TST(vdot(mcol(R2, 0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), R2, 0, 4);
  TST(vdot(mcol(R2, 1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), R2, 1, 5);
  TST(vdot(mcol(R2, 2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), R2, 2, 6);
#undef TST
  const real eps = 1.0e-5;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] += eps;
#define TST(expr1, expr2, n1, n2, n3, cc) \
  s2 = fabs(expr1) - (expr2); if (s2 > 2.220446049250313e-16) return 0; \
  l = sqrt((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
  if (l > 2.220446049250313e-16) { s2 /= l; \
    if (s2 * fudge > s + TIE) { s = s2; normalRm = 0; normalC = V((n1) / l, (n2) / l, (n3) / l); invert_normal = ((expr1) < 0); code = (cc); } }
  TST(ppv[2] * Rm[1][0] - ppv[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -Rm[2][0], Rm[1][0], 7);
  TST(ppv[2] * Rm[1][1] - ppv[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -Rm[2][1], Rm[1][1], 8);
  TST(ppv[2] * Rm[1][2] - ppv[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -Rm[2][2], Rm[1][2], 9);
  TST(ppv[0] * Rm[2][0] - ppv[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), Rm[2][0], 0, -Rm[0][0], 10);
  TST(ppv[0] * Rm[2][1] - ppv[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), Rm[2][1], 0, -Rm[0][1], 11);
  TST(ppv[0] * Rm[2][2] - ppv[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), Rm[2][2], 0, -Rm[0][2], 12);
  TST(ppv[1] * Rm[0][0] - ppv[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -Rm[1][0], Rm[0][0], 0, 13);
  TST(ppv[1] * Rm[0][1] - ppv[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -Rm[1][1], Rm[0][1], 0, 14);
  TST(ppv[1] * Rm[0][2] - ppv[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -Rm[1][2], Rm[0][2], 0, 15);
#undef TST
  if (!code) return 0;
  v3 normal;
  if (normalRm) normal = mcol(normalRm, normalRcol); else normal = mmulv(R1, normalC);
  if (invert_normal) normal = vscale(normal, -1);
  real depth = -s;
  if (code > 6) { /* edge-edge: one point */
    v3 pa = p1, pb = p2;
    for (int j = 0; j < 3; j++) { real sign = vdot(normal, mcol(R1, j)) > 0 ? 1.0 : -1.0; pa = vadd(pa, vscale(mcol(R1, j), sign * A[j])); }
    for (int j = 0; j < 3; j++) { real sign = vdot(normal, mcol(R2, j)) > 0 ? -1.0 : 1.0; pb = vadd(pb, vscale(mcol(R2, j), sign * B[j])); }
    v3 ua = mcol(R1, (code - 7) / 3), ub = mcol(R2, (code - 7) % 3);
    real alpha, beta; line_closest(pa, ua, pb, ub, &alpha, &beta);
    pb = vadd(pb, vscale(ub, beta));
    out[0].pos = pb; out[0].n = vscale(normal, -1); out[0].depth = depth;
    return 1;
  }
  /* face-something: reference box a, incident box b */
  const m3 *Ra, *Rb; v3 pa, pb; const real *Sa, *Sb;
  if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; }
  else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; }
  v3 normal2 = code <= 3 ? normal : vscale(normal, -1);
  v3 nr = mtmulv(Rb, normal2);
  real anr[3] = {fabs(nr.x), fabs(nr.y), fabs(nr.z)}, nrv[3] = {nr.x, nr.y, nr.z};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  v3 center;
  if (nrv[lanr] < 0) center = vadd(vsub(pb, pa), vscale(mcol(Rb, lanr), Sb[lanr]));
  else center = vsub(vsub(pb, pa), vscale(mcol(Rb, lanr), Sb[lanr]));
  int codeN = code <= 3 ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  real quad[8], c1 = vdot(center, mcol(Ra, code1)), c2 = vdot(center, mcol(Ra, code2));
  real m11 = vdot(mcol(Ra, code1), mcol(Rb, a1)), m12 = vdot(mcol(Ra, code1), mcol(Rb, a2));
  real m21 = vdot(mcol(Ra, code2), mcol(Rb, a1)), m22 = vdot(mcol(Ra, code2), mcol(Rb, a2));
  {
    real k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  real rect[2] = {Sa[code1], Sa[code2]}, ret[16];
  int n = clip_quad(rect, quad, ret);
  if (n < 1) return 0;
  real point[24], dep[8], det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    real k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    real k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    v3 pt = vadd(vadd(center, vscale(mcol(Rb, a1), k1)), vscale(mcol(Rb, a2), k2));
    real d = Sa[codeN] - vdot(normal2, pt);
    if (d >= 0) {
      point[cnum * 3] = pt.x; point[cnum * 3 + 1] = pt.y; point[cnum * 3 + 2] = pt.z;
      dep[cnum] = d; ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++;
    }
  }
  if (cnum < 1) return 0;
  int maxc = 4, idx[8], m = cnum;
  if (cnum > maxc) {
    int i1 = 0; real maxd = dep[0];
    for (int i = 1; i < cnum; i++) if (dep[i] > maxd + 1e-7) { maxd = dep[i]; i1 = i; }   /* ties: the first point wins (same rule as the kernels) */
    cull_points(cnum, ret, maxc, i1, idx); m = maxc;
  } else for (int i = 0; i < cnum; i++) idx[i] = i;
  for (int j = 0; j < m; j++) {
    int i = idx[j];
    v3 pw = vadd(V(point[i * 3], point[i * 3 + 1], point[i * 3 + 2]), pa);
    if (code >= 4) pw = vsub(pw, vscale(normal, dep[i]));
    out[j].pos = pw; out[j].n = vscale(normal, -1); out[j].depth = dep[i];
  }
  return m;
}
/* test entry: boxes given as pos3, rot9 (row-major), half3; out = n x (pos3, normal3, depth) */
int orc_box_box(const real* p1, const real* R1, const real* h1, const real* p2, const real* R2, const real* h2, real* out) {
  CPoint c[8]; m3 Ra = mloadr(R1), Rb = mloadr(R2);
  int n = orc_box_box_impl(vloadr(p1), &Ra, vloadr(h1), vloadr(p2), &Rb, vloadr(h2), c);
  for (int i = 0; i < n; i++) {
    out[7 * i] = c[i].pos.x; out[7 * i + 1] = c[i].pos.y; out[7 * i + 2] = c[i].pos.z;
    out[7 * i + 3] = c[i].n.x; out[7 * i + 4] = c[i].n.y; out[7 * i + 5] = c[i].n.z; out[7 * i + 6] = c[i].depth;
  }
  return n;
}

/* ---------------------------------------------------------------- world poses */
typedef struct {
  Kin K;
  m3 fR[MAXFREE];
  m3 sR[MAXSLIDE]; v3 sp[MAXSLIDE]; v3 sa[MAXSLIDE];
  m3 cR[MAXCOL]; v3 cp[MAXCOL];
  v3 lo[MAXCOL], hi[MAXCOL];
} Poses;

static void body_frame(const prb_model* M, const State* S, const Poses* P, int body, int link, m3* R, v3* p) {
  if (body < 0) { m3 I = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; *R = I; *p = V(0, 0, 0); }
  else if (body == 0 && link < 0) { *R = mload(M->arm_base_rot); *p = vload(M->arm_base_pos); }
  else if (body == 0) { *R = P->K.R[link]; *p = P->K.p[link]; }
  else if (body <= M->n_free) { *R = P->fR[body - 1]; *p = vloadr(S->fpos[body - 1]); }
  else { *R = P->sR[body - 1 - M->n_free]; *p = P->sp[body - 1 - M->n_free]; }
}
static void compute_poses(const prb_model* M, const State* S, Poses* P) {
  m3 bR = mload(M->arm_base_rot);
  arm_fk(M, S->q, &bR, vload(M->arm_base_pos), &P->K);
  for (int b = 0; b < M->n_free; b++) quat_to_mat(S->fquat[b], &P->fR[b]);
  for (int b = 0; b < M->n_slide; b++) {
    m3 R0 = mload(M->slide_rot + 9 * b); v3 ax = vload(M->slide_axis + 3 * b), p0 = vload(M->slide_pos + 3 * b);
    P->sa[b] = mmulv(&R0, ax);
    if (M->slide_jtype[b] == 0) { m3 Rq = axis_angle(ax, S->sq[b]); P->sR[b] = mmul(&R0, &Rq); P->sp[b] = p0; }
    else { P->sR[b] = R0; P->sp[b] = vadd(p0, vscale(P->sa[b], S->sq[b])); }
  }
  for (int c = 0; c < M->n_col; c++) {
    m3 R; v3 p; body_frame(M, S, P, M->col_body[c], M->col_link[c], &R, &p);
    m3 lR = mload(M->col_rot + 9 * c);
    P->cR[c] = mmul(&R, &lR);
    P->cp[c] = vadd(p, mmulv(&R, vload(M->col_pos + 3 * c)));
    v3 h = vload(M->col_half + 3 * c), e;
    e.x = fabs(P->cR[c].m[0][0]) * h.x + fabs(P->cR[c].m[0][1]) * h.y + fabs(P->cR[c].m[0][2]) * h.z;
    e.y = fabs(P->cR[c].m[1][0]) * h.x + fabs(P->cR[c].m[1][1]) * h.y + fabs(P->cR[c].m[1][2]) * h.z;
    e.z = fabs(P->cR[c].m[2][0]) * h.x + fabs(P->cR[c].m[2][1]) * h.y + fabs(P->cR[c].m[2][2]) * h.z;
    P->lo[c] = vsub(P->cp[c], e); P->hi[c] = vadd(P->cp[c], e);
  }
}

/* ---------------------------------------------------------------- contacts */
typedef struct { int ca, cb; v3 pa, pb, n; real dist; } Contact;

/* Persistent-manifold capacity: Bullet keeps at most 4 points per pair of collision objects
 * (btPersistentManifold, MANIFOLD_CACHE_SIZE 4) and, when a 5th arrives, keeps the deepest and the
 * three that span the largest area (sortCachedPoints).  Restated here as a batch reduction over
 * the candidates of one object pair: deepest, farthest from it, largest triangle, largest gain. */
#define MTIE 2e-6   /* ties: within 2 um (depth, distance, height over the longest edge) the first candidate wins; same rule as the kernels */
static int reduce_manifold(const Contact* c, int n, int* keep) {
  if (n <= 4) { for (int i = 0; i < n; i++) keep[i] = i; return n; }
  int i0 = 0;
  for (int i = 1; i < n; i++) if (c[i].dist < c[i0].dist - MTIE) i0 = i;   /* ties: the first candidate wins (same rule as the kernels) */
  int i1 = -1; real best = -1;
  for (int i = 0; i < n; i++) if (i != i0) { v3 d = vsub(c[i].pb, c[i0].pb); real v = vnorm(d); if (v > best + MTIE) { best = v; i1 = i; } }
  int i2 = -1; best = -1;
  v3 e01 = vsub(c[i1].pb, c[i0].pb);
  const real atol = vnorm(e01) * MTIE;
  for (int i = 0; i < n; i++) if (i != i0 && i != i1) { v3 x = vcross(vsub(c[i].pb, c[i0].pb), e01); real v = vnorm(x); if (v > best + atol) { best = v; i2 = i; } }
  int i3 = -1; best = -1;
  for (int i = 0; i < n; i++) if (i != i0 && i != i1 && i != i2) {
    v3 a = vsub(c[i].pb, c[i0].pb), b = vsub(c[i].pb, c[i1].pb), d = vsub(c[i].pb, c[i2].pb);
    real v = vnorm(vcross(a, b)) + vnorm(vcross(b, d)) + vnorm(vcross(d, a));
    if (v > best + 4 * atol) { best = v; i3 = i; }
  }
  int sel[4] = {i0, i1, i2, i3}, m = 0;
  for (int i = 0; i < n; i++) if (i == sel[0] || i == sel[1] || i == sel[2] || i == sel[3]) keep[m++] = i;
  return m;
}
/* all narrow-phase candidates of the last detect_contacts call, before manifold reduction (tests) */
static real g_last_cand[4 * MAXCONTACT][10]; static int g_last_ncand = 0;
int orc_last_candidates(real* out) { for (int i = 0; i < g_last_ncand; i++) for (int k = 0; k < 10; k++) out[10 * i + k] = g_last_cand[i][k]; return g_last_ncand; }
static int detect_contacts(const prb_model* M, const Poses* P, Contact* C, int maxc) {
  g_last_ncand = 0;
  static Contact cand[4 * MAXCONTACT];
  int ncand = 0, nc = 0;
  int run_start = 0, run_oa = -1, run_ob = -1;
  for (int k = 0; k <= M->n_pair; k++) {
    int oa = -2, ob = -2;
    if (k < M->n_pair) { oa = M->col_obj[M->pair_a[k]]; ob = M->col_obj[M->pair_b[k]]; }
    if (k == M->n_pair || oa != run_oa || ob != run_ob) {   /* close the run of the previous object pair */
      int n = ncand - run_start, keep[4 * MAXCONTACT];
      if (n > 0) {
        int m = reduce_manifold(cand + run_start, n, keep);
        for (int i = 0; i < m && nc < maxc; i++) C[nc++] = cand[run_start + keep[i]];
      }
      ncand = 0; run_start = 0; run_oa = oa; run_ob = ob;
      if (k == M->n_pair) break;
    }
    int a = M->pair_a[k], b = M->pair_b[k];
    if (P->lo[a].x > P->hi[b].x || P->hi[a].x < P->lo[b].x || P->lo[a].y > P->hi[b].y || P->hi[a].y < P->lo[b].y ||
        P->lo[a].z > P->hi[b].z || P->hi[a].z < P->lo[b].z) continue;
    CPoint cp[8];
    int n = orc_box_box_impl(P->cp[a], &P->cR[a], vload(M->col_half + 3 * a), P->cp[b], &P->cR[b], vload(M->col_half + 3 * b), cp);
    for (int i = 0; i < n && ncand < 4 * MAXCONTACT; i++) {
      cand[ncand].ca = a; cand[ncand].cb = b; cand[ncand].n = cp[i].n; cand[ncand].pb = cp[i].pos; cand[ncand].dist = -cp[i].depth;
      cand[ncand].pa = vadd(cp[i].pos, vscale(cp[i].n, -cp[i].depth));
      if (g_last_ncand < 4 * MAXCONTACT) { real* o = g_last_cand[g_last_ncand++]; o[0] = a; o[1] = b; o[2] = cp[i].pos.x; o[3] = cp[i].pos.y; o[4] = cp[i].pos.z; o[5] = cp[i].n.x; o[6] = cp[i].n.y; o[7] = cp[i].n.z; o[8] = -cp[i].depth; o[9] = 0; }
      ncand++;
    }
  }
  return nc;
}

/* ---------------------------------------------------------------- dynamics (world-frame ABA) */
typedef struct { real v[6]; } sv;           /* spatial vector: motion (w, vO) or force (n, f) about the world origin */
typedef struct { real m[6][6]; } sm;
static sv sv_zero(void) { sv r; memset(&r, 0, sizeof(r)); return r; }
static sv sv_make(v3 a, v3 b) { sv r = {{a.x, a.y, a.z, b.x, b.y, b.z}}; return r; }
static v3 sv_top(const sv* s) { return V(s->v[0], s->v[1], s->v[2]); }
static v3 sv_bot(const sv* s) { return V(s->v[3], s->v[4], s->v[5]); }
static sv sv_add(sv a, sv b) { for (int i = 0; i < 6; i++) a.v[i] += b.v[i]; return a; }
static sv sv_scale(sv a, real s) { for (int i = 0; i < 6; i++) a.v[i] *= s; return a; }
static real sv_dot(const sv* a, const sv* b) { real s = 0; for (int i = 0; i < 6; i++) s += a->v[i] * b->v[i]; return s; }
static sv crm(const sv* v, const sv* m) { /* motion cross motion */
  v3 w = sv_top(v), vo = sv_bot(v), mw = sv_top(m), mv = sv_bot(m);
  return sv_make(vcross(w, mw), vadd(vcross(w, mv), vcross(vo, mw)));
}
static sv crf(const sv* v, const sv* f) { /* motion cross force */
  v3 w = sv_top(v), vo = sv_bot(v), n = sv_top(f), fl = sv_bot(f);
  return sv_make(vadd(vcross(w, n), vcross(vo, fl)), vcross(w, fl));
}
static sv sm_mul(const sm* A, const sv* x) { sv r; for (int i = 0; i < 6; i++) { real s = 0; for (int j = 0; j < 6; j++) s += A->m[i][j] * x->v[j]; r.v[i] = s; } return r; }
static void spatial_inertia(real m, v3 c, const m3* Ic, sm* I) {
  real cx[3][3] = {{0, -c.z, c.y}, {c.z, 0, -c.x}, {-c.y, c.x, 0}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real cc = 0; for (int k = 0; k < 3; k++) cc += cx[i][k] * cx[k][j];
    I->m[i][j] = Ic->m[i][j] - m * cc;
    I->m[i][j + 3] = m * cx[i][j];
    I->m[i + 3][j] = -m * cx[i][j];
    I->m[i + 3][j + 3] = (i == j) ? m : 0;
  }
}
typedef struct {
  sv S[MAXD], U[MAXD];
  real D[MAXD];
  sm IA[MAXD];
} Aba;
static void arm_S(const prb_model* M, const Kin* K, sv* S) {
  for (int i = 0; i < M->nd; i++)
    S[i] = M->arm_jtype[i] == 0 ? sv_make(K->a[i], vcross(K->p[i], K->a[i])) : sv_make(V(0, 0, 0), K->a[i]);
}
/* forward dynamics: qdd from (q, qd, tau=0) with gravity; also leaves IA/U/D for impulse responses */
static void arm_aba(const prb_model* M, const Kin* K, const real* qd, real* qdd, Aba* A) {
  int nd = M->nd;
  sv vel[MAXD], c[MAXD], pA[MAXD], acc[MAXD];
  real u[MAXD];
  real g = M->params[PRB_P_GRAVITY_Z];
  real klin = M->params[PRB_P_ARM_LIN_DAMP], kang = M->params[PRB_P_ARM_ANG_DAMP];
  arm_S(M, K, A->S);
  for (int i = 0; i < nd; i++) {
    int par = M->arm_parent[i];
    sv vj = sv_scale(A->S[i], qd[i]);
    vel[i] = par < 0 ? vj : sv_add(vel[par], vj);
    c[i] = crm(&vel[i], &vj);
    m3 Il = mload(M->arm_inertia + 9 * i), t = mmul(&K->R[i], &Il), Rt = mtrans(&K->R[i]), Iw = mmul(&t, &Rt);
    spatial_inertia(M->arm_mass[i], K->c[i], &Iw, &A->IA[i]);
    sv Iv = sm_mul(&A->IA[i], &vel[i]);
    pA[i] = crf(&vel[i], &Iv);
    /* Bullet's per-link damping (linear on COM velocity, angular with the link inertia); zeroed
       for the arm by changeDynamics (environments.py:421-422) but kept general */
    if (klin != 0 || kang != 0) {
      v3 w = sv_top(&vel[i]), vc = vadd(sv_bot(&vel[i]), vcross(w, K->c[i]));
      v3 fl = vscale(vc, M->arm_mass[i] * (klin + klin * vnorm(vc)));
      v3 na = vscale(mmulv(&Iw, w), (kang + kang * vnorm(w)));
      sv fd = sv_make(vadd(na, vcross(K->c[i], fl)), fl);
      pA[i] = sv_add(pA[i], fd);
    }
  }
  for (int i = nd - 1; i >= 0; i--) {
    int par = M->arm_parent[i];
    A->U[i] = sm_mul(&A->IA[i], &A->S[i]);
    A->D[i] = sv_dot(&A->S[i], &A->U[i]);
    u[i] = -M->arm_jdamp[i] * qd[i] - sv_dot(&A->S[i], &pA[i]);
    if (par >= 0) {
      sm Ia = A->IA[i];
      for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) Ia.m[r][s] -= A->U[i].v[r] * A->U[i].v[s] / A->D[i];
      sv Iac = sm_mul(&Ia, &c[i]);
      sv pa = sv_add(sv_add(pA[i], Iac), sv_scale(A->U[i], u[i] / A->D[i]));
      for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) A->IA[par].m[r][s] += Ia.m[r][s];
      pA[par] = sv_add(pA[par], pa);
    }
  }
  sv a0 = sv_make(V(0, 0, 0), V(0, 0, -g)); /* gravity as base acceleration */
  for (int i = 0; i < nd; i++) {
    int par = M->arm_parent[i];
    sv ap = sv_add(par < 0 ? a0 : acc[par], c[i]);
    qdd[i] = (u[i] - sv_dot(&A->U[i], &ap)) / A->D[i];
    acc[i] = sv_add(ap, sv_scale(A->S[i], qdd[i]));
  }
}
/* response of the arm to a generalized impulse tau: out = M^-1 tau (Bullet: calcAccelerationDeltasMultiDof) */
static void arm_minv(const prb_model* M, const Aba* A, const real* tau, real* out) {
  int nd = M->nd;
  sv pA[MAXD], acc[MAXD]; real u[MAXD];
  for (int i = 0; i < nd; i++) pA[i] = sv_zero();
  for (int i = nd - 1; i >= 0; i--) {
    int par = M->arm_parent[i];
    u[i] = tau[i] - sv_dot(&A->S[i], &pA[i]);
    if (par >= 0) pA[par] = sv_add(pA[par], sv_add(pA[i], sv_scale(A->U[i], u[i] / A->D[i])));
  }
  for (int i = 0; i < nd; i++) {
    int par = M->arm_parent[i];
    sv ap = par < 0 ? sv_zero() : acc[par];
    out[i] = (u[i] - sv_dot(&A->U[i], &ap)) / A->D[i];
    acc[i] = sv_add(ap, sv_scale(A->S[i], out[i]));
  }
}
/* joint-space inertia matrix via unit impulse responses, for tests (column j = M^-1 e_j inverted by caller) */
void orc_arm_minv_matrix(const prb_model* M, const real* q, real* out /* nd*nd */) {
  Kin K; m3 bR = mload(M->arm_base_rot); Aba A; real qd[MAXD] = {0}, qdd[MAXD];
  arm_fk(M, q, &bR, vload(M->arm_base_pos), &K);
  arm_aba(M, &K, qd, qdd, &A);
  for (int j = 0; j < M->nd; j++) {
    real tau[MAXD] = {0}, col[MAXD]; tau[j] = 1;
    arm_minv(M, &A, tau, col);
    for (int i = 0; i < M->nd; i++) out[i * M->nd + j] = col[i];
  }
}
void orc_arm_qdd(const prb_model* M, const real* q, const real* qd, real* qdd) {
  Kin K; m3 bR = mload(M->arm_base_rot); Aba A;
  arm_fk(M, q, &bR, vload(M->arm_base_pos), &K);
  arm_aba(M, &K, qd, qdd, &A);
}

/* ---------------------------------------------------------------- constraint rows + PGS */
typedef struct {
  real J[MAXV], B[MAXV];
  real rhs, cfm, invD, lo, hi, lambda, mu;
  int normal_row;    /* friction / spin rows: index of their normal row */
} Row;

typedef struct {
  const prb_model* M; const State* S; const Poses* P; const Aba* A;
  real v[MAXV];     /* velocities after the unconstrained update */
  m3 fIinv[MAXFREE];
} SolveCtx;

static int nv_total(const prb_model* M) { return M->nd + 6 * M->n_free + M->n_slide; }
static void apply_minv(const SolveCtx* X, const real* J, real* B) {
  const prb_model* M = X->M;
  int nd = M->nd;
  arm_minv(M, X->A, J, B);
  for (int b = 0; b < M->n_free; b++) {
    int o = nd + 6 * b;
    for (int k = 0; k < 3; k++) B[o + k] = J[o + k] / M->free_mass[b];
    v3 t = mmulv(&X->fIinv[b], V(J[o + 3], J[o + 4], J[o + 5]));
    B[o + 3] = t.x; B[o + 4] = t.y; B[o + 5] = t.z;
  }
  for (int s = 0; s < M->n_slide; s++) {
    int o = nd + 6 * M->n_free + s;
    real minv = M->slide_jtype[s] == 1 ? 1.0 / M->slide_mass[s] : 1.0 / M->slide_inertia[s];
    B[o] = J[o] * minv;
  }
}
/* add the Jacobian of "unit force `dir` at world point `pt`" (or unit torque if angular_only) on a collider's body */
static void add_point_jac(const SolveCtx* X, int col, v3 pt, v3 dir, real sign, int angular_only, real* J) {
  const prb_model* M = X->M;
  int body = M->col_body[col], nd = M->nd;
  if (body < 0) return;
  if (body == 0) {
    int link = M->col_link[col];
    if (link < 0) return;
    for (int j = 0; j < nd; j++) if (is_ancestor(M, j, link)) {
      real g;
      if (M->arm_jtype[j] == 0) g = angular_only ? vdot(X->P->K.a[j], dir) : vdot(X->P->K.a[j], vcross(vsub(pt, X->P->K.p[j]), dir));
      else g = angular_only ? 0.0 : vdot(X->P->K.a[j], dir);
      J[j] += sign * g;
    }
  } else if (body <= M->n_free) {
    int b = body - 1, o = nd + 6 * b;
    v3 r = vsub(pt, vloadr(X->S->fpos[b]));
    v3 t = angular_only ? dir : vcross(r, dir);
    if (!angular_only) { J[o] += sign * dir.x; J[o + 1] += sign * dir.y; J[o + 2] += sign * dir.z; }
    J[o + 3] += sign * t.x; J[o + 4] += sign * t.y; J[o + 5] += sign * t.z;
  } else {
    int s = body - 1 - M->n_free, o = nd + 6 * M->n_free + s;
    real g;
    if (M->slide_jtype[s] == 0) g = angular_only ? vdot(X->P->sa[s], dir) : vdot(X->P->sa[s], vcross(vsub(pt, X->P->sp[s]), dir));
    else g = angular_only ? 0.0 : vdot(X->P->sa[s], dir);
    J[o] += sign * g;
  }
}
static real row_finish(const SolveCtx* X, Row* r, real cfm_raw) {
  int nv = nv_total(X->M);
  apply_minv(X, r->J, r->B);
  real d = 0, rel = 0;
  for (int i = 0; i < nv; i++) { d += r->J[i] * r->B[i]; rel += r->J[i] * X->v[i]; }
  d += cfm_raw;
  r->invD = d > 2.220446049250313e-16 ? 1.0 / d : 0.0;
  r->lambda = 0; r->normal_row = -1; r->mu = 0;
  return rel;
}
static void btPlaneSpace1(v3 n, v3* p, v3* q) {
  if (fabs(n.z) > 0.7071067811865475244) {
    real a = n.y * n.y + n.z * n.z, k = 1.0 / sqrt(a);
    *p = V(0, -n.z * k, n.y * k);
    *q = V(a * k, -n.x * p->z, n.x * p->y);
  } else {
    real a = n.x * n.x + n.y * n.y, k = 1.0 / sqrt(a);
    *p = V(-n.y * k, n.x * k, 0);
    *q = V(-n.z * p->y, n.z * p->x, a * k);
  }
}
static real resolve_row(Row* r, real* dv, int nv) {
  real jdv = 0;
  for (int i = 0; i < nv; i++) jdv += r->J[i] * dv[i];
  real delta = r->rhs - r->lambda * r->cfm - jdv * r->invD;
  real sum = r->lambda + delta;
  if (sum < r->lo) { delta = r->lo - r->lambda; r->lambda = r->lo; }
  else if (sum > r->hi) { delta = r->hi - r->lambda; r->lambda = r->hi; }
  else r->lambda = sum;
  for (int i = 0; i < nv; i++) dv[i] += r->B[i] * delta;
  return delta;
}
static void resolve_cone(Row* ra, Row* rb, real* dv, int nv) { /* btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows */
  real ja = 0, jb = 0;
  for (int i = 0; i < nv; i++) { ja += ra->J[i] * dv[i]; jb += rb->J[i] * dv[i]; }
  real dB = rb->rhs - rb->lambda * rb->cfm - jb * rb->invD, sumB = rb->lambda + dB;
  real dA = ra->rhs - ra->lambda * ra->cfm - ja * ra->invD, sumA = ra->lambda + dA;
  if (sumA < ra->lo || sumA > ra->hi || sumB < rb->lo || sumB > rb->hi) {
    /* Bullet: angle = atan2(sumA, sumB); limits |lo*sin(angle)|, |lo*cos(angle)|; written without
       the trigonometric round trip: sin = sumA/r, cos = sumB/r */
    real rr = sqrt(sumA * sumA + sumB * sumB);
    real ca = rr > 0 ? fabs(ra->lo * sumA / rr) : 0.0, cb = rr > 0 ? fabs(rb->lo * sumB / rr) : fabs(rb->lo);
    if (sumA < -ca) { dA = -ca - ra->lambda; ra->lambda = -ca; }
    else if (sumA > ca) { dA = ca - ra->lambda; ra->lambda = ca; }
    else ra->lambda = sumA;
    if (sumB < -cb) { dB = -cb - rb->lambda; rb->lambda = -cb; }
    else if (sumB > cb) { dB = cb - rb->lambda; rb->lambda = cb; }
    else rb->lambda = sumB;
  } else { ra->lambda = sumA; rb->lambda = sumB; }
  for (int i = 0; i < nv; i++) dv[i] += ra->B[i] * dA + rb->B[i] * dB;
}

/* diagnostics of the last substep (tests) */
static int g_last_contacts = 0, g_last_rows = 0;
static int g_last_pairs[MAXCONTACT][2];
static real g_last_cdata[MAXCONTACT][10];   /* pa, pb, n, dist */
void orc_last_contact_pairs(int* out) { for (int i = 0; i < g_last_contacts; i++) { out[2 * i] = g_last_pairs[i][0]; out[2 * i + 1] = g_last_pairs[i][1]; } }
int orc_last_contacts(void) { return g_last_contacts; }
/* contacts of the last substep: {pa, pb, n, dist} each (tests) */
void orc_last_contact_data(real* out) { for (int i = 0; i < g_last_contacts; i++) for (int k = 0; k < 10; k++) out[10 * i + k] = g_last_cdata[i][k]; }
int orc_last_rows(void) { return g_last_rows; }

/* One stepSimulation() (environments.py:490,535): btMultiBodyDynamicsWorld::
 * internalSingleStepSimulation = collide -> ABA (v += dt*qdd) -> build rows -> 50 PGS
 * iterations -> v += dv -> integrate positions. */
static Row g_rows[MAXROW];
static void substep(const prb_model* M, State* S) {
  const real dt = M->params[PRB_P_DT], g = M->params[PRB_P_GRAVITY_Z];
  const real vmax = M->params[PRB_P_MAX_COORD_VEL];
  int nd = M->nd, nv = nv_total(M);
  static Poses P; static Aba A; static Contact C[MAXCONTACT];
  SolveCtx X; X.M = M; X.S = S; X.P = &P; X.A = &A;
  compute_poses(M, S, &P);
  int nc = detect_contacts(M, &P, C, MAXCONTACT);
  /* ---- unconstrained velocity update */
  real qdd[MAXD];
  arm_aba(M, &P.K, S->qd, qdd, &A);
  for (int i = 0; i < nd; i++) { real v = S->qd[i] + dt * qdd[i]; if (v > vmax) v = vmax; if (v < -vmax) v = -vmax; X.v[i] = v; }
  for (int b = 0; b < M->n_free; b++) {
    int o = nd + 6 * b;
    real kl = M->free_lin_damp[b], ka = M->free_ang_damp[b];
    v3 vl = vloadr(S->fvel[b]), w = vloadr(S->fang[b]);
    v3 acc = vadd(V(0, 0, g), vscale(vl, -(kl + kl * vnorm(vl))));
    v3 wb = mtmulv(&P.fR[b], w);
    v3 Id = vload(M->free_inertia + 3 * b);
    v3 Iw = V(Id.x * wb.x, Id.y * wb.y, Id.z * wb.z);
    v3 gy = vcross(wb, Iw);
    v3 wdb = V(-gy.x / Id.x, -gy.y / Id.y, -gy.z / Id.z);
    wdb = vadd(wdb, vscale(wb, -(ka + ka * vnorm(wb))));
    v3 wd = mmulv(&P.fR[b], wdb);
    real nvv[6] = {vl.x + dt * acc.x, vl.y + dt * acc.y, vl.z + dt * acc.z, w.x + dt * wd.x, w.y + dt * wd.y, w.z + dt * wd.z};
    for (int k = 0; k < 6; k++) { real v = nvv[k]; if (v > vmax) v = vmax; if (v < -vmax) v = -vmax; X.v[o + k] = v; }
    m3 D = {{{1 / Id.x, 0, 0}, {0, 1 / Id.y, 0}, {0, 0, 1 / Id.z}}}, t = mmul(&P.fR[b], &D), Rt = mtrans(&P.fR[b]);
    X.fIinv[b] = mmul(&t, &Rt);
  }
  for (int s = 0; s < M->n_slide; s++) {
    int o = nd + 6 * M->n_free + s;
    real qdd_s;
    if (M->slide_jtype[s] == 1) qdd_s = g * P.sa[s].z;                       /* gravity along the axis; linearDamping=0 (scenes.py:171-175) */
    else { real ka = M->slide_ang_damp[s], w = S->sqd[s]; qdd_s = -w * (ka + ka * fabs(w)); }
    real v = S->sqd[s] + dt * qdd_s; if (v > vmax) v = vmax; if (v < -vmax) v = -vmax;
    X.v[o] = v;
  }
  /* ---- rows */
  int nr = 0;
  Row* rows = g_rows;
  const real erp = M->params[PRB_P_ERP_JOINT], erp2 = M->params[PRB_P_ERP_CONTACT];
  /* joint limit rows (btMultiBodyJointLimitConstraint; created at URDF load => before the motors) */
  for (int i = 0; i < nd; i++) {
    if (M->arm_lo[i] > M->arm_hi[i]) continue;
    for (int side = 0; side < 2; side++) {
      real pen = side == 0 ? S->q[i] - M->arm_lo[i] : M->arm_hi[i] - S->q[i];
      if (pen > 0) continue;
      Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
      r->J[i] = side == 0 ? 1.0 : -1.0;
      real rel = row_finish(&X, r, 0.0);
      real e = pen > -0.04 ? erp : erp2;   /* split-impulse threshold picks m_erp vs m_erp2 */
      real poserr = -pen * e / dt, velerr = -rel;
      r->rhs = (poserr + velerr) * r->invD; r->cfm = 0; r->lo = 0; r->hi = M->params[PRB_P_LIMIT_MAX_IMPULSE];
    }
  }
  int n_joint_rows_start = nr;
  (void)n_joint_rows_start;
  /* arm joint motors (btMultiBodyJointMotor): default velocity motor or POSITION_CONTROL (kp, kd=1) */
  for (int i = 0; i < nd; i++) {
    if (S->mmaximp[i] <= 0) continue;
    Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
    r->J[i] = 1.0;
    real rel = row_finish(&X, r, 0.0);
    real kd = M->params[PRB_P_MOTOR_KD];
    real target_v = S->mkp[i] * (S->mtarget[i] - S->q[i]) / dt + X.v[i] + kd * (0.0 - X.v[i]);
    r->rhs = (target_v - rel) * r->invD; r->cfm = 0; r->lo = -S->mmaximp[i]; r->hi = S->mmaximp[i];
  }
  for (int s = 0; s < M->n_slide; s++) {
    int o = nd + 6 * M->n_free + s;
    const double* mot = M->slide_motor + 4 * s;
    real maximp = mot[3] < 0 ? M->params[PRB_P_DEFAULT_MOTOR_IMPULSE] : mot[3];
    if (maximp <= 0) continue;
    Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
    r->J[o] = 1.0;
    real rel = row_finish(&X, r, 0.0);
    real target_v = mot[1] * (mot[0] - S->sq[s]) / dt + X.v[o] + mot[2] * (0.0 - X.v[o]);
    r->rhs = (target_v - rel) * r->invD; r->cfm = 0; r->lo = -maximp; r->hi = maximp;
  }
  if (M->gear_a >= 0) { /* btMultiBodyGearConstraint between the Panda fingers (environments.py:400-405) */
    Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
    real ratio = M->params[PRB_P_GEAR_RATIO];
    r->J[M->gear_a] = 1.0; r->J[M->gear_b] = ratio;
    real rel = row_finish(&X, r, 0.0);
    /* Bullet: posError = 0 unless a relative position target is set; erp scales the velocity error */
    real velerr = -rel * M->params[PRB_P_GEAR_ERP];
    r->rhs = velerr * r->invD; r->cfm = 0; r->lo = -M->params[PRB_P_GEAR_MAX_IMPULSE]; r->hi = M->params[PRB_P_GEAR_MAX_IMPULSE];
  }
  int n_noncontact = nr;
  /* contact rows: normals first, then per contact spin row (optional) and the two friction rows */
  int normal_of[MAXCONTACT], spin_of[MAXCONTACT], fric_of[MAXCONTACT];
  for (int k = 0; k < nc; k++) {
    Contact* c = &C[k];
    Row* r = &rows[nr]; memset(r, 0, sizeof(*r));
    add_point_jac(&X, c->ca, c->pa, c->n, 1.0, 0, r->J);
    add_point_jac(&X, c->cb, c->pb, c->n, -1.0, 0, r->J);
    real cfm = 0, e = erp2;
    real sa = M->col_stiffness[c->ca], sb = M->col_stiffness[c->cb];
    if (sa >= 0 || sb >= 0) { /* BT_CONTACT_FLAG_CONTACT_STIFFNESS_DAMPING */
      real ka = sa >= 0 ? sa : 1e18, kb = sb >= 0 ? sb : 1e18;
      real da = sa >= 0 ? M->col_damping[c->ca] : 0.1, db = sb >= 0 ? M->col_damping[c->cb] : 0.1;
      real kk = 1.0 / (1.0 / ka + 1.0 / kb), dd = da + db;
      real denom = dt * kk + dd; if (denom < 1.1920929e-7) denom = 1.1920929e-7;
      cfm = 1.0 / denom; e = dt * kk / denom;
    }
    cfm /= dt;
    real rel = row_finish(&X, r, cfm);
    real pen = c->dist + M->params[PRB_P_LINEAR_SLOP];
    real poserr = 0, velerr = -rel;
    if (pen > 0) velerr -= pen / dt; else poserr = -pen * e / dt;
    r->rhs = (poserr + velerr) * r->invD; r->cfm = cfm * r->invD; r->lo = 0; r->hi = 1e10;
    normal_of[k] = nr++;
  }
  for (int k = 0; k < nc; k++) {
    Contact* c = &C[k];
    real spin = M->col_spin[c->ca] * M->col_friction[c->ca] + M->col_spin[c->cb] * M->col_friction[c->cb];
    spin_of[k] = -1;
    if (spin > 0) {
      Row* r = &rows[nr]; memset(r, 0, sizeof(*r));
      add_point_jac(&X, c->ca, c->pa, c->n, 1.0, 1, r->J);
      add_point_jac(&X, c->cb, c->pb, c->n, -1.0, 1, r->J);
      real rel = row_finish(&X, r, 0.0);
      r->rhs = -rel * r->invD; r->cfm = 0; r->mu = spin; r->normal_row = normal_of[k]; r->lo = 0; r->hi = 0;
      spin_of[k] = nr++;
    }
  }
  for (int k = 0; k < nc; k++) {
    Contact* c = &C[k];
    real mu = M->col_friction[c->ca] * M->col_friction[c->cb];
    if (mu > 10) mu = 10; if (mu < -10) mu = -10;
    v3 t1, t2; btPlaneSpace1(c->n, &t1, &t2);
    fric_of[k] = nr;
    for (int d = 0; d < 2; d++) {
      Row* r = &rows[nr]; memset(r, 0, sizeof(*r));
      v3 t = d == 0 ? t1 : t2;
      add_point_jac(&X, c->ca, c->pa, t, 1.0, 0, r->J);
      add_point_jac(&X, c->cb, c->pb, t, -1.0, 0, r->J);
      real rel = row_finish(&X, r, 0.0);
      r->rhs = -rel * r->invD; r->cfm = 0; r->mu = mu; r->normal_row = normal_of[k]; r->lo = 0; r->hi = 0;
      nr++;
    }
  }
  g_last_contacts = nc; g_last_rows = nr;
  for (int k = 0; k < nc; k++) {
    g_last_pairs[k][0] = C[k].ca; g_last_pairs[k][1] = C[k].cb;
    real* o = g_last_cdata[k];
    o[0] = C[k].pa.x; o[1] = C[k].pa.y; o[2] = C[k].pa.z; o[3] = C[k].pb.x; o[4] = C[k].pb.y; o[5] = C[k].pb.z;
    o[6] = C[k].n.x; o[7] = C[k].n.y; o[8] = C[k].n.z; o[9] = C[k].dist;
  }
  /* ---- PGS, btMultiBodyConstraintSolver::solveSingleIteration order */
  real dv[MAXV]; for (int i = 0; i < nv; i++) dv[i] = 0;
  for (int it = 0; it < M->solver_iters; it++) {
    for (int j = 0; j < n_noncontact; j++) {
      int idx = (it & 1) ? j : n_noncontact - 1 - j;
      resolve_row(&rows[idx], dv, nv);
    }
    for (int k = 0; k < nc; k++) resolve_row(&rows[normal_of[k]], dv, nv);
    for (int k = 0; k < nc; k++) if (spin_of[k] >= 0) {
      Row* r = &rows[spin_of[k]]; real tot = rows[r->normal_row].lambda;
      if (tot > 0) { r->lo = -r->mu * tot; r->hi = r->mu * tot; resolve_row(r, dv, nv); }
    }
    for (int k = 0; k < nc; k++) {
      Row *ra = &rows[fric_of[k]], *rb = &rows[fric_of[k] + 1];
      real tot = rows[ra->normal_row].lambda;
      ra->lo = -ra->mu * tot; ra->hi = ra->mu * tot; rb->lo = -rb->mu * tot; rb->hi = rb->mu * tot;
      resolve_cone(ra, rb, dv, nv);
    }
  }
  /* ---- apply + integrate (stepPositionsMultiDof) */
  for (int i = 0; i < nv; i++) { real v = X.v[i] + dv[i]; if (v > vmax) v = vmax; if (v < -vmax) v = -vmax; X.v[i] = v; }
  for (int i = 0; i < nd; i++) { S->qd[i] = X.v[i]; S->q[i] += dt * S->qd[i]; }
  for (int b = 0; b < M->n_free; b++) {
    int o = nd + 6 * b;
    for (int k = 0; k < 3; k++) { S->fvel[b][k] = X.v[o + k]; S->fang[b][k] = X.v[o + 3 + k]; S->fpos[b][k] += dt * S->fvel[b][k]; }
    v3 w = vloadr(S->fang[b]);
    real ang = vnorm(w);
    if (ang * dt > 0.7853981633974483) ang = 0.5 * 1.5707963267948966 / dt;   /* ANGULAR_MOTION_THRESHOLD */
    v3 ax;
    if (ang < 0.001) ax = vscale(w, 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang);
    else ax = vscale(w, sin(0.5 * ang * dt) / ang);
    real dq[4] = {ax.x, ax.y, ax.z, cos(ang * dt * 0.5)}, nq[4];
    quat_mul(dq, S->fquat[b], nq);
    real n = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
    for (int k = 0; k < 4; k++) S->fquat[b][k] = nq[k] / n;
  }
  for (int s = 0; s < M->n_slide; s++) {
    int o = nd + 6 * M->n_free + s;
    S->sqd[s] = X.v[o]; S->sq[s] += dt * S->sqd[s];
  }
}
/* Rows of the most recent substep, for the independent LCP check (tests/test_cpu_oracle_independent.py):
 * J and B = M^-1 J^T (nv each) and {rhs, cfm, invD, lo, hi, lambda, mu, normal_row} of row i. */
void orc_last_row(const prb_model* M, int i, real* J, real* B, real* sc) {
  const Row* r = &g_rows[i];
  int nv = nv_total(M);
  for (int k = 0; k < nv; k++) { J[k] = r->J[k]; B[k] = r->B[k]; }
  sc[0] = r->rhs; sc[1] = r->cfm; sc[2] = r->invD; sc[3] = r->lo; sc[4] = r->hi; sc[5] = r->lambda; sc[6] = r->mu; sc[7] = (real)r->normal_row;
}
int orc_nv(const prb_model* M) { return nv_total(M); }

void orc_substeps(const prb_model* M, real* state, int n) {
  State S; state_unpack(M, state, &S);
  for (int i = 0; i < n; i++) substep(M, &S);
  state_pack(M, &S, state);
}

/* ---------------------------------------------------------------- observation (environments.py:720-894) */
typedef struct {
  real obs_quat[24], achieved_goal[16], desired_goal[16], cag[4], fps[24], joints[8], velocity[6],
      observation[24], proprio, reward, success, target_poses[8];
} Out;
static int out_dim(const prb_model* M) { return M->obs_dim + 2 * M->goal_dim + 4 + M->fps_dim + 8 + 6 + M->observation_dim + 3 + M->n_ik; }
int orc_out_dim(const prb_model* M) { return out_dim(M); }
static void out_pack(const prb_model* M, const Out* O, real* o) {
  int k = 0;
  for (int i = 0; i < M->obs_dim; i++) o[k++] = O->obs_quat[i];
  for (int i = 0; i < M->goal_dim; i++) o[k++] = O->achieved_goal[i];
  for (int i = 0; i < M->goal_dim; i++) o[k++] = O->desired_goal[i];
  for (int i = 0; i < 4; i++) o[k++] = O->cag[i];
  for (int i = 0; i < M->fps_dim; i++) o[k++] = O->fps[i];
  for (int i = 0; i < 8; i++) o[k++] = O->joints[i];
  for (int i = 0; i < 6; i++) o[k++] = O->velocity[i];
  for (int i = 0; i < M->observation_dim; i++) o[k++] = O->observation[i];
  o[k++] = O->proprio; o[k++] = O->reward; o[k++] = O->success;
  for (int i = 0; i < M->n_ik; i++) o[k++] = O->target_poses[i];
}
static real py_mod(real a, real b) { real r = fmod(a, b); if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }
static real dial_to_0_1_range(real q) { return (py_mod(q, 2.0) * PI) / (2.2 * PI); }   /* scenes.py:342-343, precedence kept */

/* rayTest(point_one, point_two)[0] against every box collider: nearest hit */
static int ray_boxes(const prb_model* M, const Poses* P, v3 from, v3 to, real* frac_out) {
  real best = 1.0; int hit = -1;
  v3 d = vsub(to, from);
  for (int c = 0; c < M->n_col; c++) {
    v3 o = mtmulv(&P->cR[c], vsub(from, P->cp[c])), dl = mtmulv(&P->cR[c], d), h = vload(M->col_half + 3 * c);
    real t0 = 0, t1 = 1; int ok = 1;
    for (int k = 0; k < 3 && ok; k++) {
      real ok_ = vget(o, k), dk = vget(dl, k), hk = vget(h, k);
      if (fabs(dk) < 1e-12) { if (ok_ < -hk || ok_ > hk) ok = 0; }
      else {
        real ta = (-hk - ok_) / dk, tb = (hk - ok_) / dk;
        if (ta > tb) { real t = ta; ta = tb; tb = t; }
        if (ta > t0) t0 = ta; if (tb < t1) t1 = tb;
        if (t0 > t1) ok = 0;
      }
    }
    if (ok && t0 < best) { best = t0; hit = c; }
  }
  *frac_out = best;
  return hit;
}
static real compute_reward(const prb_model* M, const real* ag, const real* dg) {
  if (M->play) { /* playRewardFunc.py:66-77 */
    for (int k = 0; k < 3; k++) if (fabs(dg[k] - ag[k]) > 0.05) return -1;
    real eg[3], ea[3]; orc_euler_from_quat(dg + 3, eg); orc_euler_from_quat(ag + 3, ea);
    for (int k = 0; k < 3; k++) if (fabs(eg[k] - ea[k]) > PI / 4) return -1;
    if (fabs(dg[7] - ag[7]) > 0.025) return -1;
    if (fabs(dg[8] - ag[8]) > 0.04) return -1;     /* compare_door ignores its limit argument */
    if (fabs(dg[9] - ag[9]) > 0.01) return -1;
    if (fabs(dg[10] - ag[10]) > 0.3) return -1;
    return 0;
  }
  /* environments.py:283-304, num_goals = 1 */
  real d = sqrt((ag[0] - dg[0]) * (ag[0] - dg[0]) + (ag[1] - dg[1]) * (ag[1] - dg[1]) + (ag[2] - dg[2]) * (ag[2] - dg[2]));
  return d > M->params[PRB_P_SPARSE_THRESH] ? -1.0 : -d;
}
void orc_compute_reward(const prb_model* M, const real* ag, const real* dg, int64_t B, real* out) {
  for (int64_t i = 0; i < B; i++) out[i] = compute_reward(M, ag + i * M->goal_dim, dg + i * M->goal_dim);
}
static real f32(real x) { return (real)(float)x; }
static void calc_state(const prb_model* M, State* S, Out* O) {
  static Poses P; compute_poses(M, S, &P);
  int nd = M->nd;
  /* calc_actor_state :746-764 */
  v3 ep; m3 eR; site_pose(M, &P.K, 0, &ep, &eR);
  real eq[4]; mat_to_quat(&eR, eq);
  /* link velocity of the EE frame origin */
  v3 w = V(0, 0, 0), vl = V(0, 0, 0);
  int ee_link = M->site_link[0];
  for (int j = 0; j < nd; j++) if (is_ancestor(M, j, ee_link)) {
    if (M->arm_jtype[j] == 0) { w = vadd(w, vscale(P.K.a[j], S->qd[j])); vl = vadd(vl, vscale(vcross(P.K.a[j], vsub(ep, P.K.p[j])), S->qd[j])); }
    else vl = vadd(vl, vscale(P.K.a[j], S->qd[j]));
  }
  real grip = M->arm_kind == 0 ? S->q[M->grip_obs_dof] * 23.0 : S->q[M->grip_obs_dof];
  for (int j = 0; j < 8; j++) O->joints[j] = M->joints_obs_dof[j] >= 0 ? S->q[M->joints_obs_dof[j]] : 0.0;
  /* gripper_proprioception :720-743 */
  if (M->arm_kind == 0) {
    v3 g1, g2, wr; m3 t;
    site_pose(M, &P.K, 2, &g1, &t); site_pose(M, &P.K, 3, &g2, &t); site_pose(M, &P.K, 1, &wr, &t);
    v3 avg = vscale(vadd(g1, g2), 0.5), ew = vsub(ep, wr);
    v3 p1 = vsub(ep, vscale(ew, 0.5)), p2 = vadd(avg, vscale(ew, 0.2));
    real frac; int hit = ray_boxes(M, &P, p1, p2, &frac);
    int li = hit >= 0 ? M->col_urdf_link[hit] : -1;
    O->proprio = (hit < 0 || frac == 1.0 || li == 18 || li == 20) ? 0 : 1;
  } else O->proprio = -1;
  /* state vector :804-839 */
  real st[32]; int n = 0;
  st[n++] = ep.x; st[n++] = ep.y; st[n++] = ep.z;
  if (M->return_velocity) { st[n++] = vl.x; st[n++] = vl.y; st[n++] = vl.z; }
  if (M->use_orientation) { for (int k = 0; k < 4; k++) st[n++] = eq[k]; }
  st[n++] = grip;
  real ag[16]; int na = 0;
  if (M->n_free > 0) {
    int nobj = 1;
    for (int b = 0; b < nobj; b++) {
      for (int k = 0; k < 3; k++) st[n++] = S->fpos[b][k];
      if (M->use_orientation) for (int k = 0; k < 4; k++) st[n++] = S->fquat[b][k];
      if (M->return_velocity) for (int k = 0; k < 3; k++) st[n++] = S->fvel[b][k];
      for (int k = 0; k < 3; k++) ag[na++] = S->fpos[b][k];
      if (M->use_orientation) for (int k = 0; k < 4; k++) ag[na++] = S->fquat[b][k];
    }
    if (M->play) { /* :781-791 drawer y, door, button, dial */
      real e[4] = {S->fpos[1][1], S->sq[0], S->sq[1], dial_to_0_1_range(S->sq[2])};
      for (int k = 0; k < 4; k++) { st[n++] = e[k]; ag[na++] = e[k]; }
    }
  } else { ag[0] = ep.x; ag[1] = ep.y; ag[2] = ep.z; na = 3; }
  /* quaternion_safe_the_obs :868-894 (play only) */
  if (M->play) {
    if (S->last_valid > 0.5) {
      int flip_e = 1, flip_o = 1;
      for (int k = 0; k < 4; k++) {
        real a = st[3 + k], l = S->lastq[k]; int sa = (a > 0) - (a < 0), sl = (l > 0) - (l < 0);
        if (sa != -sl) flip_e = 0;
        a = st[11 + k]; l = S->lastq[4 + k]; sa = (a > 0) - (a < 0); sl = (l > 0) - (l < 0);
        if (sa != -sl) flip_o = 0;
      }
      if (flip_e) for (int k = 0; k < 4; k++) st[3 + k] = -st[3 + k];
      if (flip_o) for (int k = 0; k < 4; k++) { st[11 + k] = -st[11 + k]; ag[3 + k] = -ag[3 + k]; }
    }
    for (int k = 0; k < 4; k++) { S->lastq[k] = st[3 + k]; S->lastq[4 + k] = st[11 + k]; }
    S->last_valid = 1;
  }
  for (int i = 0; i < M->obs_dim; i++) O->obs_quat[i] = f32(st[i]);
  for (int i = 0; i < M->goal_dim; i++) { O->achieved_goal[i] = f32(ag[i]); O->desired_goal[i] = f32(S->goal[i]); }
  O->cag[0] = f32(ep.x); O->cag[1] = f32(ep.y); O->cag[2] = f32(ep.z); O->cag[3] = f32(grip);
  { int k = 0; O->fps[k++] = ep.x; O->fps[k++] = ep.y; O->fps[k++] = ep.z;
    if (M->use_orientation) for (int j = 0; j < 4; j++) O->fps[k++] = st[3 + j];
    O->fps[k++] = grip;
    if (M->n_free > 0) for (int j = 0; j < na; j++) O->fps[k++] = ag[j];
    for (int j = 0; j < k; j++) O->fps[j] = f32(O->fps[j]); }
  O->velocity[0] = vl.x; O->velocity[1] = vl.y; O->velocity[2] = vl.z; O->velocity[3] = w.x; O->velocity[4] = w.y; O->velocity[5] = w.z;
  /* 'observation' :859 = state[0:3] + euler(state[3:7]) + state[7:]  (6/12/18-dim quirk kept) */
  { real e[3]; orc_euler_from_quat(st + 3, e); int k = 0;
    O->observation[k++] = st[0]; O->observation[k++] = st[1]; O->observation[k++] = st[2];
    O->observation[k++] = e[0]; O->observation[k++] = e[1]; O->observation[k++] = e[2];
    for (int j = 7; j < n; j++) O->observation[k++] = st[j]; }
  /* reward is computed by the gym layer on the float32 dict entries (environments.py:211) */
  O->reward = compute_reward(M, O->achieved_goal, O->desired_goal);
  O->success = O->reward < 0 ? 0 : 1;
}

/* ---------------------------------------------------------------- step (environments.py:206-214) */
static real clampr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }
void orc_step(const prb_model* M, real* state, const real* action, real* out) {
  State S; state_unpack(M, state, &S);
  Out O; memset(&O, 0, sizeof(O));
  /* perform_action :915-934: 0 absolute_rpy, 1 relative_rpy, 2 absolute_quat, 3 relative_quat, 4 absolute_joints,
   * 5 relative_joints; clip to the action space first (:207, bounds :88-112) */
  const int atype = (int)(M->params[PRB_P_ACTION_TYPE] + 0.5), adim = (atype == 2 || atype == 3) ? 8 : (atype >= 4 ? M->n_ik + 1 : 7);
  real a[8] = {0};
  for (int k = 0; k < adim - 1; k++) a[k] = clampr(action[k], -M->params[PRB_P_ACTION_HIGH_XYZ], M->params[PRB_P_ACTION_HIGH_XYZ]);
  const real grip = clampr(action[adim - 1], -M->params[PRB_P_ACTION_HIGH_GRIP], M->params[PRB_P_ACTION_HIGH_GRIP]);
  real jp[MAXD];
  if (atype >= 4) {                                  /* relative / absolute_joint_step :973-981 (6-joint arms) */
    for (int i = 0; i < M->n_ik; i++) jp[i] = (atype == 5 ? S.q[i] : 0) + a[i];
  } else {
    real tp[3] = {a[0], a[1], a[2]}, tq[4];
    if (atype == 0) orc_quat_from_euler(a + 3, tq);  /* absolute_rpy_step :955-961 */
    else if (atype == 2) { for (int k = 0; k < 4; k++) tq[k] = a[3 + k]; }   /* absolute_quat_step :936-943 */
    else {                                           /* relative to getLinkState(arm, ee) :945-953, 962-970 */
      real sites[28];
      orc_fk_sites(M, S.q, sites);
      for (int k = 0; k < 3; k++) tp[k] += sites[k];
      if (atype == 1) {
        real rpy[3]; orc_euler_from_quat(sites + 3, rpy);
        for (int k = 0; k < 3; k++) rpy[k] += a[3 + k];
        orc_quat_from_euler(rpy, tq);
      } else for (int k = 0; k < 4; k++) tq[k] = sites[3 + k] + a[3 + k];
    }
    if (atype >= 2) {                                /* commanded quaternions are used normalised */
      real nrm = sqrt(tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2] + tq[3] * tq[3]);
      if (nrm > 1e-12) for (int k = 0; k < 4; k++) tq[k] /= nrm; else { tq[0] = tq[1] = tq[2] = 0; tq[3] = 1; }
    }
    /* goto :984-1007 */
    if (M->arm_kind == 0) orc_calc_angles(M, S.q, tp, tq, jp);
    else orc_ik(M, S.q, tp, tq, M->ik_iters, jp);     /* Panda: one call on the live arm, 200 iterations (:995-997) */
  }
  /* goto_joint_poses :1010-1034 */
  real dt = M->params[PRB_P_DT];
  for (int i = 0; i < M->n_ik; i++) {
    real t = clampr(jp[i], M->ctrl_ll[i], M->ctrl_ul[i]);
    t = clampr(t, S.q[i] - M->ctrl_inc[i], S.q[i] + M->ctrl_inc[i]);
    S.mtarget[i] = t; S.mkp[i] = M->params[PRB_P_MOTOR_KP]; S.mmaximp[i] = M->params[PRB_P_ARM_FORCE] * dt;
    O.target_poses[i] = t;
  }
  /* close_gripper :1037-1073 (mimic entries read the CURRENT position of their source joint) */
  for (int k = 0; k < M->n_grip; k++) {
    int d = M->grip_dof[k];
    real t = M->grip_mimic[k] >= 0 ? S.q[M->grip_mimic[k]] : M->grip_scale[k] * grip + M->grip_offset[k];
    S.mtarget[d] = t; S.mkp[d] = M->params[PRB_P_MOTOR_KP]; S.mmaximp[d] = M->grip_force[k] * dt;
  }
  for (int i = 0; i < M->n_substeps; i++) substep(M, &S);   /* runSimulation :485-490 */
  calc_state(M, &S, &O);
  state_pack(M, &S, state);
  out_pack(M, &O, out);
}
void orc_calc_state(const prb_model* M, real* state, real* out) {
  State S; state_unpack(M, state, &S); Out O; memset(&O, 0, sizeof(O));
  calc_state(M, &S, &O); state_pack(M, &S, state); out_pack(M, &O, out);
}

/* ---------------------------------------------------------------- reset (environments.py:173-187, 492-603) */
void orc_init_state(const prb_model* M, real* state) {
  State S; memset(&S, 0, sizeof(S));
  for (int i = 0; i < M->nd; i++) { S.mkp[i] = 0; S.mtarget[i] = 0; S.mmaximp[i] = M->params[PRB_P_DEFAULT_MOTOR_IMPULSE]; }
  for (int b = 0; b < M->n_free; b++) {
    for (int k = 0; k < 3; k++) S.fpos[b][k] = M->free_pos0[3 * b + k];
    for (int k = 0; k < 4; k++) S.fquat[b][k] = M->free_quat0[4 * b + k];
  }
  state_pack(M, &S, state);
}
void orc_reset(const prb_model* M, real* state, uint64_t seed, uint32_t env_id, real* out) {
  State S; state_unpack(M, state, &S);
  Out O; memset(&O, 0, sizeof(O));
  real r = 0;
  int guard = 0;
  while (r > -1 && guard++ < 16) {
    uint32_t attempt = (uint32_t)S.reset_count;
    S.reset_count += 1;
    real u[4];
    /* reset_object_pos :519-556 */
    int nobj = M->n_free > 0 ? 1 : 0;
    for (int t = 0; t < 4; t++) {
      if (M->play) {
        for (int k = 0; k < 3; k++) { S.fpos[1][k] = M->free_pos0[3 + k]; S.fvel[1][k] = 0; S.fang[1][k] = 0; }
        for (int k = 0; k < 4; k++) S.fquat[1][k] = M->free_quat0[4 + k];
        for (int s = 0; s < M->n_slide; s++) { S.sq[s] = 0; S.sqd[s] = 0; }
      }
      if (nobj) {
        rng4(seed, env_id, attempt, (uint32_t)t, u);
        for (int k = 0; k < 3; k++) { S.fpos[0][k] = M->obj_lo[k] + (M->obj_hi[k] - M->obj_lo[k]) * u[k]; S.fvel[0][k] = 0; S.fang[0][k] = 0; }
        S.fpos[0][2] += M->params[PRB_P_OBJ_RESET_DZ];
        S.fquat[0][0] = 0; S.fquat[0][1] = 0; S.fquat[0][2] = 0.7071; S.fquat[0][3] = 0.7071;
      }
      for (int i = 0; i < M->settle_steps; i++) substep(M, &S);
      int oob = 0;
      if (nobj) for (int k = 0; k < 3; k++) if (S.fpos[0][k] > M->env_hi[k]) oob = 1;
      if (!oob) break;
    }
    /* reset_arm :575-596 */
    rng4(seed, env_id, attempt, 4, u);
    real np_[3];
    for (int k = 0; k < 3; k++) np_[k] = M->goal_lo[k] + (M->goal_hi[k] - M->goal_lo[k]) * u[k];
    np_[2] += M->params[PRB_P_RESET_Z_OFFSET];
    /* reset_arm_joints(restJointPositions): UR5 sets joints 0..5; Panda sets 0..6 and finger 9 (poses[7]=0) */
    for (int i = 0; i < M->n_ik; i++) { S.q[i] = M->arm_rest[i]; S.qd[i] = 0; }
    if (M->arm_kind == 1) { S.q[M->n_ik] = 0; S.qd[M->n_ik] = 0; }
    real jp[MAXD];
    real dorn[4] = {(real)M->default_orn[0], (real)M->default_orn[1], (real)M->default_orn[2], (real)M->default_orn[3]};
    orc_ik(M, S.q, np_, dorn, M->ik_reset_iters, jp);
    for (int i = 0; i < 6; i++) { S.q[i] = jp[i]; S.qd[i] = 0; }      /* [0:6] for both arms (:593) */
    /* reset_goal_pos :492-516 */
    rng4(seed, env_id, attempt, 5, u);
    if (!M->play) for (int k = 0; k < 3; k++) S.goal[k] = M->goal_lo[k] + (M->goal_hi[k] - M->goal_lo[k]) * u[k];
    else {
      calc_state(M, &S, &O);                 /* self.calc_state()['achieved_goal'] (also updates the quat history) */
      int idx = (int)(u[0] * M->goal_dim); if (idx >= M->goal_dim) idx = M->goal_dim - 1;
      for (int k = 0; k < M->goal_dim; k++) S.goal[k] = O.achieved_goal[k];
      S.goal[idx] = f32(S.goal[idx] + u[1]);
    }
    calc_state(M, &S, &O);
    r = O.reward;
  }
  state_pack(M, &S, state);
  out_pack(M, &O, out);
}

/* playEnv.reset(o) (environments.py:173-187 with o given): objects re-seated from the observation without settle steps
 * (:541-556), arm from the rest pose through one IK call to the observed end-effector pose (:582-596), new goal (:492-516),
 * again while the state already satisfies it.  Deviations kept identical to the kernels: the object is read
 * at its real offset in the obs_quat layout (the reference's 11 / 10 indexing is wrong for the 19-D play layout), and
 * restore_env also restores drawer y / door / button / dial from the observation (the reference leaves them at defaults). */
void orc_reset_to(const prb_model* M, real* state, const real* obs, int restore_env, uint64_t seed, uint32_t env_id, real* out) {
  State S; state_unpack(M, state, &S);
  Out O; memset(&O, 0, sizeof(O));
  const int rv = M->return_velocity, uo = M->use_orientation;
  const int o_obj = 3 + (rv ? 3 : 0) + (uo ? 4 : 0) + 1;
  if (M->play) {
    for (int k = 0; k < 3; k++) { S.fpos[1][k] = M->free_pos0[3 + k]; S.fvel[1][k] = 0; S.fang[1][k] = 0; }
    for (int k = 0; k < 4; k++) S.fquat[1][k] = M->free_quat0[4 + k];
    for (int s = 0; s < M->n_slide; s++) { S.sq[s] = 0; S.sqd[s] = 0; }
  }
  if (M->n_free > 0) {
    for (int k = 0; k < 3; k++) { S.fpos[0][k] = obs[o_obj + k]; S.fvel[0][k] = 0; S.fang[0][k] = 0; }
    if (uo) for (int k = 0; k < 4; k++) S.fquat[0][k] = obs[o_obj + 3 + k];
    else { S.fquat[0][0] = 0; S.fquat[0][1] = 0; S.fquat[0][2] = 0; S.fquat[0][3] = 1; }
  }
  if (M->play && restore_env) {
    const int o_env = o_obj + 7;
    S.fpos[1][1] = obs[o_env];
    S.sq[0] = obs[o_env + 1]; S.sq[1] = obs[o_env + 2]; S.sq[2] = obs[o_env + 3] * 2.2;
  }
  real tp[3] = {obs[0], obs[1], obs[2]}, tq[4];
  for (int k = 0; k < 4; k++) tq[k] = uo ? obs[(rv ? 6 : 3) + k] : (real)M->default_orn[k];
  for (int i = 0; i < M->n_ik; i++) { S.q[i] = M->arm_rest[i]; S.qd[i] = 0; }
  if (M->arm_kind == 1) { S.q[M->n_ik] = 0; S.qd[M->n_ik] = 0; }
  real jp[MAXD];
  orc_ik(M, S.q, tp, tq, M->ik_reset_iters, jp);
  for (int i = 0; i < 6; i++) { S.q[i] = jp[i]; S.qd[i] = 0; }
  real r = 0;
  int guard = 0;
  while (r > -1 && guard++ < 16) {
    uint32_t attempt = (uint32_t)S.reset_count;
    S.reset_count += 1;
    real u[4];
    rng4(seed, env_id, attempt, 5, u);
    if (!M->play) for (int k = 0; k < 3; k++) S.goal[k] = M->goal_lo[k] + (M->goal_hi[k] - M->goal_lo[k]) * u[k];
    else {
      calc_state(M, &S, &O);
      int idx = (int)(u[0] * M->goal_dim); if (idx >= M->goal_dim) idx = M->goal_dim - 1;
      for (int k = 0; k < M->goal_dim; k++) S.goal[k] = O.achieved_goal[k];
      S.goal[idx] = f32(S.goal[idx] + u[1]);
    }
    calc_state(M, &S, &O);
    r = O.reward;
  }
  state_pack(M, &S, state);
  out_pack(M, &O, out);
}

