// prb_stream.cuh — the split ("stream") step pipeline.
//
// One stepSimulation() substep = two launches:
//   prb_setup_kernel   one WARP per env: integrate the previous substep's solution, then kinematics,
//                      collision detection, mass matrix and its inverse, unconstrained velocities and
//                      the constraint rows of this substep.  The rows (Jacobian segments J and
//                      M^-1 J^T, right-hand sides, limits) are written to a per-env record stream in
//                      HBM instead of shared memory.  The last launch of an env step also runs the
//                      fused observation / reward write.
//   prb_pgs_kernel     one THREAD per env: the 50 projected-Gauss-Seidel sweeps, in velocity space,
//                      in btMultiBodyConstraintSolver::solveSingleIteration order.  A thread walks
//                      its env's record stream sequentially (vectorised loads, no inter-thread
//                      communication); the velocity change dv (<= 27 floats) lives in shared memory
//                      (lane-interleaved, conflict-free), accumulated impulses in thread-local memory.
//
// Why: in the fused warp-per-env kernel (prb_step_kernel) 76 % of all warp instructions were the
// PGS sweeps, where a whole warp serves ONE serial chain of row updates (profiles/r1_v7_ncu.md).
// With a thread per env the same chain costs ~1/16 of the warp instructions per env and the row
// data streams from L2/HBM, so this part of the path is bandwidth-shaped rather than issue-bound.
//
// Reference path: environments.py:485-490 (12 x stepSimulation), Bullet btMultiBodyDynamicsWorld.
#pragma once
#include "prb_kernels.cuh"

// ---- record stream.  Envs are grouped by 32 (one group = the 32 envs a solver warp serves) and the
// group's records are interleaved at 32-byte granularity: float4 number q of env (g, l) lives at
// float4 index  g * 32 * SB_Q + (q >> 1) * 64 + 2 * l + (q & 1).  Solver lanes that walk their envs'
// streams in lockstep therefore read 1 KB contiguous per pair of float4 loads (fully coalesced), and
// the setup kernel writes whole 32-byte sectors.  All offsets below are in float4 units ("q").
#define Q_HDR 0          // {n_jrow, n_contact, n_spin, overflow flag} as ints
#define Q_VSTAR 2        // 8 q: unconstrained velocities v* of the substep (word i = DoF i)
#define Q_DV 10          // 8 q: solver output M^-1 J^T lambda
#define Q_MINV 18        // 12 rows x 3 q: arm inverse mass matrix (row d at Q_MINV + 3 d, zero padded)
#define Q_JROW 54        // SB_MAXJROW x 2 q: {d | d2 << 8, sign, rhs, invD} {lo, hi, -, -}
#define SB_MAXJROW 40
#define Q_CN 134         // SB_MAXCONTACT x 1 q, normal pass:   {packed, cfm * invD, rhs0, invD0}
#define SB_MAXCONTACT 32
#define Q_CF 166         // SB_MAXCONTACT x 2 q, friction pass: {packed, mu, rhs2, rhs3} {invD2, invD3, -, -}
#define Q_CS 230         // SB_MAXCONTACT x 1 q, spin pass (compact list of the contacts that have a
                         // torsional row): {packed, spin coefficient, rhs1, invD1}
#define Q_POOL 262       // SB_MAXCONTACT x 48 q: contact c, row k (normal, spin, friction 1, 2), body x (A, B)
                         // at Q_POOL + 48 c + 6 (2 k + x): J in the first nq float4, B = M^-1 J^T in the
                         // next nq, nq = ceil(n / 4), zero padded (n = 12 | 9 arm, 6 free body, 1 slide body)
#define SB_Q 1798        // float4 per env (even)
// packed contact word: offA | nA << 5 | offB << 9 | nB << 14 | c << 18  (offX: first dv index of body X's
// segment, nX: its length, 0 when the body is static; c: contact index)

struct SV {              // one env's column of its group
  float4* b;
  PRB_D float4& q(int i) const { return b[((i >> 1) << 6) + (i & 1)]; }
  PRB_D float& w(int i) const { return reinterpret_cast<float*>(&b[(((i >> 2) >> 1) << 6) + ((i >> 2) & 1)])[i & 3]; }
};
PRB_D SV sv_of(float* sbuf, int e) {
  SV s;
  s.b = reinterpret_cast<float4*>(sbuf) + (size_t)(e >> 5) * (SB_Q * 32) + (e & 31) * 2;
  return s;
}
PRB_D int pack_contact(int offA, int nA, int offB, int nB, int c) { return offA | (nA << 5) | (offB << 9) | (nB << 14) | (c << 18); }

struct SetupCfg {
  static constexpr int MAXJROW = SB_MAXJROW;
  static constexpr int MAXCONTACT = SB_MAXCONTACT;
  static constexpr int MAXOVL = 32;
  static constexpr int MAXCAND = 128;
  static constexpr int WPB = 4;
};

// shared memory of one env in the setup kernel: state + the substep's kinematics / collision scratch
template <class CFG>
struct SetupMemT {
  typedef CFG Cfg;
  float q[PRB_MAXD], qd[PRB_MAXD], mtarget[PRB_MAXD], mkp[PRB_MAXD], mmaximp[PRB_MAXD];
  float fpos[PRB_MAXFREE][3], fquat[PRB_MAXFREE][4], fvel[PRB_MAXFREE][3], fang[PRB_MAXFREE][3];
  float sq[PRB_MAXSLIDE], sqd[PRB_MAXSLIDE];
  float goal[12], lastq[8], last_valid, reset_count;
  float lp[PRB_MAXD][3], la[PRB_MAXD][3], lc[PRB_MAXD][3], lw[PRB_MAXD][3], lv[PRB_MAXD][3];
  float fR[PRB_MAXFREE][9], fIinv[PRB_MAXFREE][6];
  float sp[PRB_MAXSLIDE][3], sR[PRB_MAXSLIDE][9];
  float Minv[PRB_MAXD][PRB_MAXD + 1], Q[PRB_MAXD];
  float vs[32];
  unsigned short ovl[CFG::MAXOVL];
  int n_ovl, n_contact, n_jrow, pool_used, overflow;
  int dbg_a, dbg_c, dbg_p, dbg_u;
  Contact ct[CFG::MAXCONTACT];
  float lR[PRB_MAXD][9], lIw[PRB_MAXD][6], lf[PRB_MAXD][3], ln[PRB_MAXD][3];
  float Mm[PRB_MAXD][PRB_MAXD + 1];
  float aabb[PRB_MAXCOL][6];
  Contact cand[CFG::MAXCAND];
};

PRB_D int nq_of(int n) { return (n + 3) >> 2; }

// J (unit force `dir` at world point pt, or unit torque when angular, on the body of collider col)
// and B = M^-1 J^T, written to the 6-float4 region q0 of the record stream; returns J.B and
// accumulates J.v*
template <int ND, class WM>
PRB_D float stream_segment(const DevModel& M, const WM& W, int col, v3 pt, v3 dir, float sign, bool angular,
                           const SV& S, int q0, float* rel) {
  const int body = M.col_body[col];
  float d = 0.f;
  if (body == 0) {
    const int link = M.col_link[col];
    float J[12], B[12];
    const unsigned anc = M.anc_mask[link];
#pragma unroll
    for (int j = 0; j < 12; j++) {
      float g = 0.f;
      if (j < ND && ((anc >> j) & 1u)) {
        v3 aj = ld3(W.la[j]);
        if (M.jtype[j] == 0) g = angular ? dot(aj, dir) : dot(aj, cross(pt - ld3(W.lp[j]), dir));
        else g = angular ? 0.f : dot(aj, dir);
      }
      J[j] = sign * g;
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      float s = 0.f;
      if (i < ND) {
#pragma unroll
        for (int j = 0; j < ND; j++) s = fmaf(W.Minv[i][j], J[j], s);
        d = fmaf(J[i], s, d); r = fmaf(J[i], W.vs[i], r);
      }
      B[i] = s;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      S.q(q0 + k) = make_float4(J[4 * k], J[4 * k + 1], J[4 * k + 2], J[4 * k + 3]);
      S.q(q0 + 3 + k) = make_float4(B[4 * k], B[4 * k + 1], B[4 * k + 2], B[4 * k + 3]);
    }
    *rel += r;
  } else if (body <= M.n_free) {
    const int b = body - 1, o = M.nd + 6 * b;
    v3 t = angular ? dir : cross(pt - ld3(W.fpos[b]), dir);
    v3 jl = angular ? V3(0, 0, 0) : dir * sign, ja = t * sign;
    float im = 1.0f / M.free_mass[b];
    v3 bl = jl * im, ba = symmul(W.fIinv[b], ja);
    float J[6] = {jl.x, jl.y, jl.z, ja.x, ja.y, ja.z}, B[6] = {bl.x, bl.y, bl.z, ba.x, ba.y, ba.z};
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) { d = fmaf(J[k], B[k], d); r = fmaf(J[k], W.vs[o + k], r); }
    S.q(q0) = make_float4(J[0], J[1], J[2], J[3]); S.q(q0 + 1) = make_float4(J[4], J[5], 0.f, 0.f);
    S.q(q0 + 2) = make_float4(B[0], B[1], B[2], B[3]); S.q(q0 + 3) = make_float4(B[4], B[5], 0.f, 0.f);
    *rel += r;
  } else {
    const int s = body - 1 - M.n_free, o = M.nd + 6 * M.n_free + s;
    v3 a = ld3(M.slide_axis_w[s]);
    float g;
    if (M.slide_jtype[s] == 0) g = angular ? dot(a, dir) : dot(a, cross(pt - ld3(W.sp[s]), dir));
    else g = angular ? 0.f : dot(a, dir);
    const float j = sign * g, bb = j * M.slide_minv[s];
    S.q(q0) = make_float4(j, 0.f, 0.f, 0.f); S.q(q0 + 1) = make_float4(bb, 0.f, 0.f, 0.f);
    d = j * bb; *rel += j * W.vs[o];
  }
  return d;
}

// constraint rows of the substep -> record stream
template <int ND, class WM>
PRB_D void phase_rows_stream(const DevModel& M, WM& W, int lane, const SV& S) {
  const float dt = M.params[P_DT], erp = M.params[P_ERP_JOINT], erp2 = M.params[P_ERP_CONTACT];
  const int nd = M.nd;
  // ---- joint rows (lane 0, serial: <= 40 rows of a few flops each): limits, motors, gear
  if (lane == 0) {
    int nr = 0;
#define PRB_PUT_JROW(d_, d2_, sg_, rhs_, invD_, lo_, hi_)                                                  \
    do {                                                                                                  \
      if (nr >= SB_MAXJROW) { W.overflow = 1; }                                                            \
      else {                                                                                              \
        S.q(Q_JROW + 2 * nr) = make_float4(__int_as_float((int)(d_) | (((int)(d2_) & 0xff) << 8)), sg_, rhs_, invD_); \
        S.q(Q_JROW + 2 * nr + 1) = make_float4(lo_, hi_, 0.f, 0.f);                                         \
        nr++;                                                                                             \
      }                                                                                                   \
    } while (0)
    for (int i = 0; i < nd; i++) {
      if (M.lo[i] > M.hi[i]) continue;
      for (int side = 0; side < 2; side++) {
        float pen = side == 0 ? W.q[i] - M.lo[i] : M.hi[i] - W.q[i];
        if (pen > 0.f) continue;
        float sg = side == 0 ? 1.0f : -1.0f;
        float invD = 1.0f / W.Minv[i][i];
        float rel = sg * W.vs[i];
        float e = pen > -0.04f ? erp : erp2;
        PRB_PUT_JROW(i, 0xff, sg, (-pen * e / dt - rel) * invD, invD, 0.f, M.params[P_LIMIT_MAX_IMPULSE]);
      }
    }
    for (int i = 0; i < nd; i++) {
      if (W.mmaximp[i] <= 0.f) continue;
      float invD = 1.0f / W.Minv[i][i];
      float v = W.vs[i];
      float target_v = W.mkp[i] * (W.mtarget[i] - W.q[i]) / dt + v + M.params[P_MOTOR_KD] * (0.f - v);
      PRB_PUT_JROW(i, 0xff, 1.0f, (target_v - v) * invD, invD, -W.mmaximp[i], W.mmaximp[i]);
    }
    for (int s = 0; s < M.n_slide; s++) {
      int o = nd + 6 * M.n_free + s;
      float maximp = M.slide_motor[s][3] < 0 ? M.params[P_DEFAULT_MOTOR_IMPULSE] : M.slide_motor[s][3];
      if (maximp <= 0.f) continue;
      float invD = 1.0f / M.slide_minv[s];
      float v = W.vs[o];
      float target_v = M.slide_motor[s][1] * (M.slide_motor[s][0] - W.sq[s]) / dt + v + M.slide_motor[s][2] * (0.f - v);
      PRB_PUT_JROW(o, 0xff, 1.0f, (target_v - v) * invD, invD, -maximp, maximp);
    }
    if (M.gear_a >= 0) {
      int a = M.gear_a, b = M.gear_b;
      float r = M.params[P_GEAR_RATIO];
      float D = W.Minv[a][a] + 2.f * r * W.Minv[a][b] + r * r * W.Minv[b][b];
      float invD = 1.0f / D;
      float rel = W.vs[a] + r * W.vs[b];
      PRB_PUT_JROW(a, b, 1.0f, (-rel * M.params[P_GEAR_ERP]) * invD, invD, -M.params[P_GEAR_MAX_IMPULSE], M.params[P_GEAR_MAX_IMPULSE]);
    }
#undef PRB_PUT_JROW
    W.n_jrow = nr;
  }
  // ---- arm inverse mass matrix (lane = row), zero padded to 12 columns
  if (lane < ND) {
    float r[12];
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = j < ND ? W.Minv[lane][j] : 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) S.q(Q_MINV + 3 * lane + k) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
  // ---- contact rows: lane = contact, fixed 48-float4 slot per contact
  const int nc = W.n_contact;
  bool has_spin = false;
  float spin = 0.f, rhs[4] = {0.f, 0.f, 0.f, 0.f}, invDs[4] = {0.f, 0.f, 0.f, 0.f};
  int packed = 0;
  if (lane < nc) {
    const Contact c = W.ct[lane];
    const int ca = c.cols & 0xff, cb = (c.cols >> 8) & 0xff;
    const int bodyA = col_dyn_body(M, ca), bodyB = col_dyn_body(M, cb);
    const int nA = bodyA >= 0 ? body_size(M, bodyA) : 0, nB = bodyB >= 0 ? body_size(M, bodyB) : 0;
    v3 n = V3(c.nx, c.ny, c.nz), pb = V3(c.pbx, c.pby, c.pbz), pa = pb + n * c.dist;
    float cfm = 0.f, e = erp2;
    float sa = M.col_stiff[ca], sb = M.col_stiff[cb];
    if (sa >= 0.f || sb >= 0.f) {       // URDF <contact> stiffness / damping on the gripper links
      float ka = sa >= 0.f ? sa : 1e18f, kb = sb >= 0.f ? sb : 1e18f;
      float da = sa >= 0.f ? M.col_damp[ca] : 0.1f, db = sb >= 0.f ? M.col_damp[cb] : 0.1f;
      float kk = 1.0f / (1.0f / ka + 1.0f / kb), dd = da + db;
      float denom = fmaxf(dt * kk + dd, 1.1920929e-7f);
      cfm = 1.0f / denom; e = dt * kk / denom;
    }
    cfm /= dt;
    spin = M.col_spin[ca] * M.col_fric[ca] + M.col_spin[cb] * M.col_fric[cb];
    has_spin = spin > 0.f;
    const float mu = clampf(M.col_fric[ca] * M.col_fric[cb], -10.f, 10.f);
    v3 t1, t2;
    plane_space(n, t1, t2);
    float cfms = 0.f;
    const int qc = Q_POOL + 48 * lane;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      v3 dir = k == 0 ? n : (k == 1 ? n : (k == 2 ? t1 : t2));
      const bool ang = (k == 1);
      if (k == 1 && !has_spin) continue;      // no torsional row: never visited by the solver
      float rel = 0.f, D = 0.f;
      if (bodyA >= 0) D += stream_segment<ND>(M, W, ca, pa, dir, 1.0f, ang, S, qc + 12 * k, &rel);
      if (bodyB >= 0) D += stream_segment<ND>(M, W, cb, pb, dir, -1.0f, ang, S, qc + 12 * k + 6, &rel);
      if (k == 0) D += cfm;
      const float invD = D > 1.1920929e-7f ? 1.0f / D : 0.f;
      if (k == 0) {
        float pen = c.dist + M.params[P_LINEAR_SLOP];
        float poserr = 0.f, velerr = -rel;
        if (pen > 0.f) velerr -= pen / dt; else poserr = -pen * e / dt;
        rhs[0] = (poserr + velerr) * invD;
        cfms = cfm * invD;
      } else rhs[k] = -rel * invD;
      invDs[k] = invD;
    }
    const int offA = bodyA >= 0 ? body_dof0(M, bodyA) : 0, offB = bodyB >= 0 ? body_dof0(M, bodyB) : 0;
    packed = pack_contact(offA, nA, offB, nB, lane);
    S.q(Q_CN + lane) = make_float4(__int_as_float(packed), cfms, rhs[0], invDs[0]);
    S.q(Q_CF + 2 * lane) = make_float4(__int_as_float(packed), mu, rhs[2], rhs[3]);
    S.q(Q_CF + 2 * lane + 1) = make_float4(invDs[2], invDs[3], 0.f, 0.f);
  }
  const unsigned spinmask = __ballot_sync(FULL, has_spin);
  if (has_spin) S.q(Q_CS + __popc(spinmask & ((1u << lane) - 1u))) = make_float4(__int_as_float(packed), spin, rhs[1], invDs[1]);
  __syncwarp();
  if (lane == 0) {
    S.q(Q_HDR) = make_float4(__int_as_float(W.n_jrow), __int_as_float(nc), __int_as_float(__popc(spinmask)), __int_as_float(W.overflow));
    if (nc > W.dbg_c) W.dbg_c = nc;
  }
  if (lane < M.nv) S.w(4 * Q_VSTAR + lane) = W.vs[lane];
}

// ============================================================================ setup kernel (warp per env)
enum { SETUP_INTEGRATE = 1, SETUP_BUILD = 2, SETUP_OBSERVE = 4 };

#ifdef PRB_EMU
static char g_emu_smem2[8 * sizeof(SetupMemT<SetupCfg>) + 256];
#define PRB_SMEM_DECL2 WM* wm = (WM*)g_emu_smem2
#else
#define PRB_SMEM_DECL2 extern __shared__ __align__(16) unsigned char prb_dyn_smem2[]; WM* wm = (WM*)prb_dyn_smem2
#endif

template <int ND>
__global__ void __launch_bounds__(32 * SetupCfg::WPB) prb_setup_kernel(const DevModel* __restrict__ Mp, float* __restrict__ state,
                                                                         float* __restrict__ sbuf, DevOut O, int N, int flags) {
  typedef SetupMemT<SetupCfg> WM;
  PRB_SMEM_DECL2;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * SetupCfg::WPB + wib;
  if (e >= N) return;
  const DevModel& M = *Mp;
  WM& W = wm[wib];
  float* st = state + (size_t)e * M.state_stride;
  const SV S = sv_of(sbuf, e);
  load_state(M, W, st, lane);
  if (lane == 0) { W.overflow = 0; W.dbg_a = 0; W.dbg_c = 0; W.dbg_p = 0; W.dbg_u = 0; }
  __syncwarp();
  if (flags & SETUP_INTEGRATE) {
    float vstar = 0.f, dv = 0.f;
    if (lane < M.nv) { vstar = S.w(4 * Q_VSTAR + lane); dv = S.w(4 * Q_DV + lane); }
    phase_integrate(M, W, lane, vstar, dv);
  }
  if (flags & SETUP_BUILD) {
    phase_fk(M, W, lane, true);
    phase_collide(M, W, lane);
    phase_crba(M, W, lane);
    phase_minv<ND>(W, lane);
    phase_vstar(M, W, lane);
    phase_rows_stream<ND>(M, W, lane, S);
    __syncwarp();
    if (lane == 0 && W.overflow && O.overflow) atomicAdd(O.overflow, 1ull);
    if (lane == 0 && O.dbg) { O.dbg[4 * e] = 0; O.dbg[4 * e + 1] = W.dbg_c; O.dbg[4 * e + 2] = 0; O.dbg[4 * e + 3] = W.n_jrow; }
  }
  if (flags & SETUP_OBSERVE) phase_observe(M, W, lane, O, (size_t)e, true);
  __syncwarp();
  if (flags & (SETUP_INTEGRATE | SETUP_OBSERVE)) store_state(M, W, st, lane);
}

// ============================================================================ PGS kernel (thread per env)
#ifndef PGS_BLOCK
#define PGS_BLOCK 64
#endif
#define PGS_MAXLAM (SB_MAXJROW + 4 * SB_MAXCONTACT)

// One body segment of a constraint row: up to 3 float4 of J and of B, zero padded, so the arithmetic
// runs on whole float4 (dv rows past the segment are multiplied by 0).  All loads of a row are issued
// before the first use: one memory round trip per row.
struct Seg { float4 j[3], b[3]; };
template <bool WITH_B>
PRB_D void seg_load(const SV& S, int q0, int n, Seg& g) {
  const int nq = nq_of(n);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    g.j[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WITH_B) g.b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < nq) { g.j[k] = S.q(q0 + k); if (WITH_B) g.b[k] = S.q(q0 + nq + k); }
  }
}
PRB_D void seg_load_b(const SV& S, int q0, int n, Seg& g) {
  const int nq = nq_of(n);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    g.b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < nq) g.b[k] = S.q(q0 + nq + k);
  }
}
PRB_D float seg_dot(const Seg& g, const float* dv, int n) {
  const int nq = nq_of(n);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (k < nq) {
      s0 = fmaf(g.j[k].x, dv[(4 * k) * PGS_BLOCK], s0);
      s1 = fmaf(g.j[k].y, dv[(4 * k + 1) * PGS_BLOCK], s1);
      s0 = fmaf(g.j[k].z, dv[(4 * k + 2) * PGS_BLOCK], s0);
      s1 = fmaf(g.j[k].w, dv[(4 * k + 3) * PGS_BLOCK], s1);
    }
  return s0 + s1;
}
PRB_D void seg_axpy(const Seg& g, float* dv, int n, float dl) {
  const int nq = nq_of(n);
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (k < nq) {
      dv[(4 * k) * PGS_BLOCK] = fmaf(g.b[k].x, dl, dv[(4 * k) * PGS_BLOCK]);
      dv[(4 * k + 1) * PGS_BLOCK] = fmaf(g.b[k].y, dl, dv[(4 * k + 1) * PGS_BLOCK]);
      dv[(4 * k + 2) * PGS_BLOCK] = fmaf(g.b[k].z, dl, dv[(4 * k + 2) * PGS_BLOCK]);
      dv[(4 * k + 3) * PGS_BLOCK] = fmaf(g.b[k].w, dl, dv[(4 * k + 3) * PGS_BLOCK]);
    }
}
struct CPk { int offA, nA, offB, nB, c; };
PRB_D CPk unpack_contact(float f) {
  const int pk = __float_as_int(f);
  CPk r;
  r.offA = pk & 31; r.nA = (pk >> 5) & 15; r.offB = (pk >> 9) & 31; r.nB = (pk >> 14) & 15; r.c = (pk >> 18) & 31;
  return r;
}

template <int ND>
__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N) {
  // 36 rows: 27 velocity DoF + slack for the zero-padded tail of the last segment
  __shared__ float dvs[36 * PGS_BLOCK];
  const int e = blockIdx.x * PGS_BLOCK + threadIdx.x;
  if (e >= N) return;
  const DevModel& M = *Mp;
  const SV S = sv_of(sbuf, e);
  float* dv = dvs + threadIdx.x;
  const int nv = M.nv, nd = M.nd;
  const float4 hdr = S.q(Q_HDR);
  const int njr = __float_as_int(hdr.x), nc = __float_as_int(hdr.y), ns = __float_as_int(hdr.z);
  const float ratio = M.params[P_GEAR_RATIO];
  float lam[PGS_MAXLAM];       // thread-local (lane-interleaved by the hardware): joint rows, then 4 per contact
#pragma unroll 1
  for (int i = 0; i < 36; i++) dv[i * PGS_BLOCK] = 0.f;
#pragma unroll 1
  for (int i = 0; i < njr + 4 * nc; i++) lam[i] = 0.f;
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    // ---- non-contact rows, sweep direction alternating per iteration
#pragma unroll 1
    for (int i = 0; i < njr; i++) {
      const int j = (it & 1) ? i : njr - 1 - i;
      const float4 r0 = S.q(Q_JROW + 2 * j), r1 = S.q(Q_JROW + 2 * j + 1);
      const int pk = __float_as_int(r0.x);
      const int d = pk & 0xff, d2 = (pk >> 8) & 0xff;
      const bool arm = d < nd;
      // the M^-1 row is needed only if the impulse changes; issue its loads now anyway (same round trip)
      float4 m0[3], m1[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        m0[k] = make_float4(0.f, 0.f, 0.f, 0.f); m1[k] = m0[k];
        if (arm) m0[k] = S.q(Q_MINV + 3 * d + k);
        if (d2 != 0xff) m1[k] = S.q(Q_MINV + 3 * d2 + k);
      }
      float u = dv[d * PGS_BLOCK];
      if (d2 != 0xff) u = fmaf(ratio, dv[d2 * PGS_BLOCK], u);
      u *= r0.y;
      const float l0 = lam[j];
      const float nl = clampf(l0 + (r0.z - u * r0.w), r1.x, r1.y);
      const float dl = (nl - l0) * r0.y;
      lam[j] = nl;
      if (dl != 0.f) {
        if (arm) {
          const float dl2 = dl * ratio;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            if (4 * k < ND) {
              dv[(4 * k) * PGS_BLOCK] = fmaf(m1[k].x, dl2, fmaf(m0[k].x, dl, dv[(4 * k) * PGS_BLOCK]));
              dv[(4 * k + 1) * PGS_BLOCK] = fmaf(m1[k].y, dl2, fmaf(m0[k].y, dl, dv[(4 * k + 1) * PGS_BLOCK]));
              dv[(4 * k + 2) * PGS_BLOCK] = fmaf(m1[k].z, dl2, fmaf(m0[k].z, dl, dv[(4 * k + 2) * PGS_BLOCK]));
              dv[(4 * k + 3) * PGS_BLOCK] = fmaf(m1[k].w, dl2, fmaf(m0[k].w, dl, dv[(4 * k + 3) * PGS_BLOCK]));
            }
          }
        } else {
          dv[d * PGS_BLOCK] = fmaf(M.slide_minv[d - nd - 6 * M.n_free], dl, dv[d * PGS_BLOCK]);
        }
      }
    }
    // ---- contact normals
    {
      float4 h = S.q(Q_CN);
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const float4 hn = S.q(Q_CN + c + 1);          // next header (slot SB_MAXCONTACT is readable padding)
        const CPk p = unpack_contact(h.x);
        Seg A, B;
        seg_load<true>(S, Q_POOL + 48 * c, p.nA, A);
        seg_load<true>(S, Q_POOL + 48 * c + 6, p.nB, B);
        const float u = seg_dot(A, dv + p.offA * PGS_BLOCK, p.nA) + seg_dot(B, dv + p.offB * PGS_BLOCK, p.nB);
        const float l0 = lam[njr + 4 * c];
        const float nl = fmaxf(l0 + (h.z - l0 * h.y - u * h.w), 0.f);
        const float dl = nl - l0;
        lam[njr + 4 * c] = nl;
        if (dl != 0.f) {
          seg_axpy(A, dv + p.offA * PGS_BLOCK, p.nA, dl);
          seg_axpy(B, dv + p.offB * PGS_BLOCK, p.nB, dl);
        }
        h = hn;
      }
    }
    // ---- spinning friction (Bullet skips the row while the normal impulse is 0)
#pragma unroll 1
    for (int i = 0; i < ns; i++) {
      const float4 h = S.q(Q_CS + i);
      const CPk p = unpack_contact(h.x);
      const float tot = lam[njr + 4 * p.c];
      if (!(tot > 0.f)) continue;
      Seg A, B;
      seg_load<true>(S, Q_POOL + 48 * p.c + 12, p.nA, A);
      seg_load<true>(S, Q_POOL + 48 * p.c + 18, p.nB, B);
      const float u = seg_dot(A, dv + p.offA * PGS_BLOCK, p.nA) + seg_dot(B, dv + p.offB * PGS_BLOCK, p.nB);
      const float lim = h.y * tot;
      const float l0 = lam[njr + 4 * p.c + 1];
      const float nl = clampf(l0 + (h.z - u * h.w), -lim, lim);
      const float dl = nl - l0;
      lam[njr + 4 * p.c + 1] = nl;
      if (dl != 0.f) {
        seg_axpy(A, dv + p.offA * PGS_BLOCK, p.nA, dl);
        seg_axpy(B, dv + p.offB * PGS_BLOCK, p.nB, dl);
      }
    }
    // ---- lateral friction: the two rows of a contact are solved together (implicit cone)
    {
      float4 h0 = S.q(Q_CF), h1 = S.q(Q_CF + 1);
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const float4 hn0 = S.q(Q_CF + 2 * c + 2), hn1 = S.q(Q_CF + 2 * c + 3);
        const CPk p = unpack_contact(h0.x);
        Seg A1, B1, A2, B2;
        seg_load<true>(S, Q_POOL + 48 * c + 24, p.nA, A1);
        seg_load<true>(S, Q_POOL + 48 * c + 30, p.nB, B1);
        seg_load<true>(S, Q_POOL + 48 * c + 36, p.nA, A2);
        seg_load<true>(S, Q_POOL + 48 * c + 42, p.nB, B2);
        const float ua = seg_dot(A1, dv + p.offA * PGS_BLOCK, p.nA) + seg_dot(B1, dv + p.offB * PGS_BLOCK, p.nB);
        const float ub = seg_dot(A2, dv + p.offA * PGS_BLOCK, p.nA) + seg_dot(B2, dv + p.offB * PGS_BLOCK, p.nB);
        const float lim = h0.y * lam[njr + 4 * c];
        const float la = lam[njr + 4 * c + 2], lb = lam[njr + 4 * c + 3];
        const float sumA = la + (h0.z - ua * h1.x);
        const float sumB = lb + (h0.w - ub * h1.y);
        float na = sumA, nb = sumB;
        if (sumA < -lim || sumA > lim || sumB < -lim || sumB > lim) {
          const float ss = sumA * sumA + sumB * sumB;
          const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
          const float ca_ = fabsf(lim * sumA) * inv, cb_ = ss > 0.f ? fabsf(lim * sumB) * inv : fabsf(lim);
          na = clampf(sumA, -ca_, ca_); nb = clampf(sumB, -cb_, cb_);
        }
        const float d1 = na - la, d2 = nb - lb;
        lam[njr + 4 * c + 2] = na; lam[njr + 4 * c + 3] = nb;
        if (d1 != 0.f) {
          seg_axpy(A1, dv + p.offA * PGS_BLOCK, p.nA, d1);
          seg_axpy(B1, dv + p.offB * PGS_BLOCK, p.nB, d1);
        }
        if (d2 != 0.f) {
          seg_axpy(A2, dv + p.offA * PGS_BLOCK, p.nA, d2);
          seg_axpy(B2, dv + p.offB * PGS_BLOCK, p.nB, d2);
        }
        h0 = hn0; h1 = hn1;
      }
    }
  }
#pragma unroll 1
  for (int i = 0; i < nv; i++) S.w(4 * Q_DV + i) = dv[i * PGS_BLOCK];
}
