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
#define Q_JROW 18        // SB_MAXJROW x 8 q: {d | d2 << 8, sign, rhs, invD} {lo, hi, 1/m of a slide DoF, LAMBDA}
                         //   {row d of the arm's M^-1: 3 q} {row d2 (gear rows only): 3 q}
#define SB_MAXJROW 40
#define Q_CN 338         // SB_MAXCONTACT x 2 q, normal pass:   {packed, cfm * invD, rhs0, invD0} {LAMBDA normal, LAMBDA spin, -, -}
#define SB_MAXCONTACT 32
#define Q_CF 402         // SB_MAXCONTACT x 2 q, friction pass: {packed, mu, rhs2, rhs3} {invD2, invD3, LAMBDA 2, LAMBDA 3}
#define Q_CS 466         // SB_MAXCONTACT x 1 q, spin pass (compact list of the contacts that have a
                         // torsional row): {packed, spin coefficient, rhs1, invD1}
#define Q_POOL 498       // SB_MAXCONTACT x 48 q: contact c, row k (normal, spin, friction 1, 2), body x (A, B)
                         // at Q_POOL + 48 c + 6 (2 k + x): J in float4 0..2, B = M^-1 J^T in float4 3..5, zero
                         // padded (n = 12 | 9 arm, 1 slide body); a free body (n = 6) stores J only: the
                         // solver rebuilds B from Q_BODY
#define Q_BODY 2034      // PRB_MAXFREE x 2 q: free body b: {1/m, Iinv xx, xy, xz} {Iinv yy, yz, zz, -} (world frame)
#define SB_Q 2038        // float4 per env (even)
// The accumulated impulses (LAMBDA words) live in the rows' own records: the setup kernel zeroes them,
// the solver reads them with the row (same sector as the row header) and writes them back.
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
  int jr_pk[CFG::MAXJROW];
  float jr_sign[CFG::MAXJROW], jr_rhs[CFG::MAXJROW], jr_invD[CFG::MAXJROW], jr_lo[CFG::MAXJROW], jr_hi[CFG::MAXJROW];
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
    *rel += r;
  } else {
    const int s = body - 1 - M.n_free, o = M.nd + 6 * M.n_free + s;
    v3 a = ld3(M.slide_axis_w[s]);
    float g;
    if (M.slide_jtype[s] == 0) g = angular ? dot(a, dir) : dot(a, cross(pt - ld3(W.sp[s]), dir));
    else g = angular ? 0.f : dot(a, dir);
    const float j = sign * g, bb = j * M.slide_minv[s];
    S.q(q0) = make_float4(j, 0.f, 0.f, 0.f); S.q(q0 + 3) = make_float4(bb, 0.f, 0.f, 0.f);
    d = j * bb; *rel += j * W.vs[o];
  }
  return d;
}

// constraint rows of the substep -> record stream
template <int ND, class WM>
PRB_D void phase_rows_stream(const DevModel& M, WM& W, int lane, const SV& S) {
  const float dt = M.params[P_DT], erp = M.params[P_ERP_JOINT], erp2 = M.params[P_ERP_CONTACT];
  const int nd = M.nd;
  // ---- joint rows: lane 0 lists them (serial: <= 40 rows of a few flops each): limits, motors, gear;
  //      then lane = row writes the 8-float4 record, including the M^-1 row(s) the row's update needs
  if (lane == 0) {
    int nr = 0;
#define PRB_PUT_JROW(d_, d2_, sg_, rhs_, invD_, lo_, hi_)                                   \
    do {                                                                                   \
      if (nr >= SB_MAXJROW) { W.overflow = 1; }                                             \
      else {                                                                               \
        W.jr_pk[nr] = (int)(d_) | (((int)(d2_) & 0xff) << 8); W.jr_sign[nr] = sg_;          \
        W.jr_rhs[nr] = rhs_; W.jr_invD[nr] = invD_; W.jr_lo[nr] = lo_; W.jr_hi[nr] = hi_;   \
        nr++;                                                                              \
      }                                                                                    \
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
  __syncwarp();
  for (int j = lane; j < W.n_jrow; j += 32) {
    const int pk = W.jr_pk[j], d = pk & 0xff, d2 = (pk >> 8) & 0xff;
    const int qj = Q_JROW + 8 * j;
    float sminv = 0.f;
    if (d >= nd) sminv = M.slide_minv[d - nd - 6 * M.n_free];
    S.q(qj) = make_float4(__int_as_float(pk), W.jr_sign[j], W.jr_rhs[j], W.jr_invD[j]);
    S.q(qj + 1) = make_float4(W.jr_lo[j], W.jr_hi[j], sminv, 0.f);
    if (d < nd) {
      float r[12];
#pragma unroll
      for (int k = 0; k < 12; k++) r[k] = k < ND ? W.Minv[d][k] : 0.f;       // M^-1 is symmetric: column d = row d
#pragma unroll
      for (int k = 0; k < 3; k++) S.q(qj + 2 + k) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    }
    if (d2 != 0xff) {
      float r[12];
#pragma unroll
      for (int k = 0; k < 12; k++) r[k] = k < ND ? W.Minv[d2][k] : 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) S.q(qj + 5 + k) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    }
  }
  // ---- free-body inverse inertia (the solver rebuilds B = M^-1 J^T of free-body segments from it)
  if (lane < M.n_free) {
    const float* I = W.fIinv[lane];
    S.q(Q_BODY + 2 * lane) = make_float4(1.0f / M.free_mass[lane], I[0], I[1], I[2]);
    S.q(Q_BODY + 2 * lane + 1) = make_float4(I[3], I[4], I[5], 0.f);
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
    S.q(Q_CN + 2 * lane) = make_float4(__int_as_float(packed), cfms, rhs[0], invDs[0]);
    S.q(Q_CN + 2 * lane + 1) = make_float4(0.f, 0.f, 0.f, 0.f);
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
// The solver thread walks its env's rows in sweep order: per iteration the joint rows (direction
// alternating), the contact normals, [the rare torsional rows,] the friction pairs.  Every row's
// record is self-contained and at an address known without reading anything else, so all loads of a
// row are issued together, and the NEXT row's record (joint rows) or header + first-body Jacobian
// (contacts; the common free-body-vs-static case needs nothing more) is loaded into registers
// while the current row is being solved: global-memory latency is off the dependent chain.
#define PGS_BLOCK 128

// float4 `off` of the stream region that starts at the EVEN float4 index the pointer was built from
PRB_D const float4* region_ptr(const SV& S, int q_even) { return S.b + ((q_even >> 1) << 6); }
PRB_D const float4& at(const float4* p, int off) { return p[((off >> 1) << 6) + (off & 1)]; }

struct CPk { int offA, nA, offB, nB, c; };
PRB_D CPk unpack_contact(float f) {
  const int pk = __float_as_int(f);
  CPk r;
  r.offA = pk & 31; r.nA = (pk >> 5) & 15; r.offB = (pk >> 9) & 31; r.nB = (pk >> 14) & 15; r.c = (pk >> 18) & 31;
  return r;
}

PRB_D float dot4(float4 j, const float* dv, float s) {
  s = fmaf(j.x, dv[0], s); s = fmaf(j.y, dv[PGS_BLOCK], s);
  s = fmaf(j.z, dv[2 * PGS_BLOCK], s); return fmaf(j.w, dv[3 * PGS_BLOCK], s);
}
PRB_D void axpy4(float4 b, float* dv, float dl) {
  dv[0] = fmaf(b.x, dl, dv[0]); dv[PGS_BLOCK] = fmaf(b.y, dl, dv[PGS_BLOCK]);
  dv[2 * PGS_BLOCK] = fmaf(b.z, dl, dv[2 * PGS_BLOCK]); dv[3 * PGS_BLOCK] = fmaf(b.w, dl, dv[3 * PGS_BLOCK]);
}
// J . dv of one body segment whose first two float4 (j0, j1) are already in registers; a third (arm
// segments) is read from the segment's stream region
PRB_D float seg_dot(float4 j0, float4 j1, const float4* reg, const float* dv, int n) {
  if (n == 0) return 0.f;
  float s = dot4(j0, dv, 0.f);
  if (n > 4) s = dot4(j1, dv + 4 * PGS_BLOCK, s);
  if (n > 8) s = dot4(at(reg, 2), dv + 8 * PGS_BLOCK, s);
  return s;
}
// dv += B dl.  Free bodies (n == 6): B = (J_lin / m, I^-1 J_ang) rebuilt from J; otherwise B is float4
// 3..5 of the segment's stream region (touched only when the impulse changes).
PRB_D void seg_axpy(float4 j0, float4 j1, const float4* reg, int n, bool first, const float* fi, float* dv, float dl) {
  if (n == 6) {
    const float* f = fi + (first ? 0 : 7 * PGS_BLOCK);
    const float im = f[0] * dl;
    const float xx = f[PGS_BLOCK], xy = f[2 * PGS_BLOCK], xz = f[3 * PGS_BLOCK];
    const float yy = f[4 * PGS_BLOCK], yz = f[5 * PGS_BLOCK], zz = f[6 * PGS_BLOCK];
    const float ax = j0.w * dl, ay = j1.x * dl, az = j1.y * dl;
    dv[0] = fmaf(j0.x, im, dv[0]);
    dv[PGS_BLOCK] = fmaf(j0.y, im, dv[PGS_BLOCK]);
    dv[2 * PGS_BLOCK] = fmaf(j0.z, im, dv[2 * PGS_BLOCK]);
    dv[3 * PGS_BLOCK] += xx * ax + xy * ay + xz * az;
    dv[4 * PGS_BLOCK] += xy * ax + yy * ay + yz * az;
    dv[5 * PGS_BLOCK] += xz * ax + yz * ay + zz * az;
  } else if (n > 0) {
    const float4 b0 = at(reg, 3);
    float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f), b2 = b1;
    if (n > 4) b1 = at(reg, 4);
    if (n > 8) b2 = at(reg, 5);
    axpy4(b0, dv, dl);
    if (n > 4) axpy4(b1, dv + 4 * PGS_BLOCK, dl);
    if (n > 8) axpy4(b2, dv + 8 * PGS_BLOCK, dl);
  }
}
// second-body segment of a row (absent in the common case): nothing prefetched
PRB_D float seg_dot_g(const float4* reg, const float* dv, int n) {
  if (n == 0) return 0.f;
  return seg_dot(at(reg, 0), n > 4 ? at(reg, 1) : make_float4(0.f, 0.f, 0.f, 0.f), reg, dv, n);
}
PRB_D void seg_axpy_g(const float4* reg, int n, bool first, const float* fi, float* dv, float dl) {
  if (n == 0) return;
  seg_axpy(at(reg, 0), n > 4 ? at(reg, 1) : make_float4(0.f, 0.f, 0.f, 0.f), reg, n, first, fi, dv, dl);
}

// friction cone of Bullet's implicit pair solve (see pgs_candidate<U_PAIR> of the fused kernel)
PRB_D void cone_clamp(float lim, float sumA, float sumB, float& na, float& nb) {
  na = sumA; nb = sumB;
  if (sumA < -lim || sumA > lim || sumB < -lim || sumB > lim) {
    const float ss = sumA * sumA + sumB * sumB;
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    const float ca_ = fabsf(lim * sumA) * inv, cb_ = ss > 0.f ? fabsf(lim * sumB) * inv : fabsf(lim);
    na = clampf(sumA, -ca_, ca_); nb = clampf(sumB, -cb_, cb_);
  }
}

struct JRow { float4 r0, r1, m0, m1, m2; };
PRB_D JRow load_jrow(const SV& S, int j) {
  const float4* p = region_ptr(S, Q_JROW + 8 * j);
  JRow r;
  r.r0 = at(p, 0); r.r1 = at(p, 1); r.m0 = at(p, 2); r.m1 = at(p, 3); r.m2 = at(p, 4);
  return r;
}

// address of word w (0..3) of a float4 of the stream (impulse write-back)
PRB_D float* word_of(const float4* p, int off, int w) { return const_cast<float*>(reinterpret_cast<const float*>(&at(p, off))) + w; }

template <int ND, bool GEAR>
__global__ void __launch_bounds__(PGS_BLOCK, 4) prb_pgs_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N) {
  __shared__ float dvs[32 * PGS_BLOCK];       // velocity change, lane-interleaved (rows >= nv stay 0: padding of the last segment)
  __shared__ float fis[14 * PGS_BLOCK];       // free-body inverse mass / inertia: 7 floats per body
  const int e = blockIdx.x * PGS_BLOCK + threadIdx.x;
  if (e >= N) return;
  const DevModel& M = *Mp;
  const SV S = sv_of(sbuf, e);
  float* dv = dvs + threadIdx.x;
  float* fi = fis + threadIdx.x;
  const int nv = M.nv, nd = M.nd;
  const float4 hdr = S.q(Q_HDR);
  const int njr = __float_as_int(hdr.x), nc = __float_as_int(hdr.y), ns = __float_as_int(hdr.z);
  const float ratio = M.params[P_GEAR_RATIO];
  {
    const float4 a0 = S.q(Q_BODY), a1 = S.q(Q_BODY + 1), b0 = S.q(Q_BODY + 2), b1 = S.q(Q_BODY + 3);
    fi[0] = a0.x; fi[PGS_BLOCK] = a0.y; fi[2 * PGS_BLOCK] = a0.z; fi[3 * PGS_BLOCK] = a0.w;
    fi[4 * PGS_BLOCK] = a1.x; fi[5 * PGS_BLOCK] = a1.y; fi[6 * PGS_BLOCK] = a1.z;
    fi[7 * PGS_BLOCK] = b0.x; fi[8 * PGS_BLOCK] = b0.y; fi[9 * PGS_BLOCK] = b0.z; fi[10 * PGS_BLOCK] = b0.w;
    fi[11 * PGS_BLOCK] = b1.x; fi[12 * PGS_BLOCK] = b1.y; fi[13 * PGS_BLOCK] = b1.z;
  }
#pragma unroll 1
  for (int i = 0; i < 32; i++) dv[i * PGS_BLOCK] = 0.f;
  const int iters = M.solver_iters;
  const float4* pool = region_ptr(S, Q_POOL);
  const float4* cn = region_ptr(S, Q_CN);
  const float4* cf = region_ptr(S, Q_CF);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    // ---- non-contact rows (limit, motor, gear), sweep direction alternating per iteration
    if (njr > 0) {
      const int dir = (it & 1) ? 1 : -1;
      int j = (it & 1) ? 0 : njr - 1;
      JRow cur = load_jrow(S, j);
#pragma unroll 1
      for (int i = 0; i < njr; i++) {
        const int jn = (i + 1 < njr) ? j + dir : j;
        const JRow nxt = load_jrow(S, jn);                 // next row's record: in flight while this row is solved
        const float l0 = cur.r1.w;
        const int w = __float_as_int(cur.r0.x);
        const int d = w & 0xff, d2 = (w >> 8) & 0xff;
        float u = dv[d * PGS_BLOCK];
        if (GEAR && d2 != 0xff) u = fmaf(ratio, dv[d2 * PGS_BLOCK], u);
        u *= cur.r0.y;
        const float nl = clampf(l0 + (cur.r0.z - u * cur.r0.w), cur.r1.x, cur.r1.y);
        const float dl = (nl - l0) * cur.r0.y;
        if (dl != 0.f) {
          const float4* p = region_ptr(S, Q_JROW + 8 * j);
          *word_of(p, 1, 3) = nl;
          if (d < nd) {
            axpy4(cur.m0, dv, dl); axpy4(cur.m1, dv + 4 * PGS_BLOCK, dl);
            if (ND > 8) axpy4(cur.m2, dv + 8 * PGS_BLOCK, dl);
            if (GEAR && d2 != 0xff) {
              const float dl2 = dl * ratio;
              axpy4(at(p, 5), dv, dl2); axpy4(at(p, 6), dv + 4 * PGS_BLOCK, dl2);
              if (ND > 8) axpy4(at(p, 7), dv + 8 * PGS_BLOCK, dl2);
            }
          } else {
            dv[d * PGS_BLOCK] = fmaf(cur.r1.z, dl, dv[d * PGS_BLOCK]);
          }
        }
        cur = nxt; j = jn;
      }
    }
    // ---- contact normals
    if (nc > 0) {
      float4 h = at(cn, 0), hl = at(cn, 1), j0 = at(pool, 0), j1 = at(pool, 1);
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const float4* reg = pool + 24 * 64 * c;                      // region_ptr(S, Q_POOL + 48 c)
        const int cnx = min(c + 1, SB_MAXCONTACT - 1);
        const float4* nreg = pool + 24 * 64 * cnx;
        const float4 hn = at(cn, 2 * cnx), hln = at(cn, 2 * cnx + 1), jn0 = at(nreg, 0), jn1 = at(nreg, 1);   // next contact
        const float l0 = hl.x;
        const CPk p = unpack_contact(h.x);
        float* dvA = dv + p.offA * PGS_BLOCK;
        float* dvB = dv + p.offB * PGS_BLOCK;
        const float u = seg_dot(j0, j1, reg, dvA, p.nA) + seg_dot_g(reg + 3 * 64, dvB, p.nB);
        const float nl = fmaxf(l0 + (h.z - l0 * h.y - u * h.w), 0.f);
        const float dl = nl - l0;
        if (dl != 0.f) {
          *word_of(cn, 2 * c + 1, 0) = nl;
          seg_axpy(j0, j1, reg, p.nA, p.offA == nd, fi, dvA, dl);
          seg_axpy_g(reg + 3 * 64, p.nB, p.offB == nd, fi, dvB, dl);
        }
        h = hn; hl = hln; j0 = jn0; j1 = jn1;
      }
    }
    // ---- torsional friction rows (rare: not prefetched); Bullet skips the row while its normal impulse is 0
#pragma unroll 1
    for (int i = 0; i < ns; i++) {
      const float4 h = S.q(Q_CS + i);
      const CPk p = unpack_contact(h.x);
      const float4 hl = at(cn, 2 * p.c + 1);
      const float tot = hl.x;
      if (!(tot > 0.f)) continue;
      const float4* rA = pool + 24 * 64 * p.c + 6 * 64;              // row 1, body A: float4 12 of the contact's slot
      const float4* rB = rA + 3 * 64;
      float* dvA = dv + p.offA * PGS_BLOCK;
      float* dvB = dv + p.offB * PGS_BLOCK;
      const float l0 = hl.y;
      const float u = seg_dot_g(rA, dvA, p.nA) + seg_dot_g(rB, dvB, p.nB);
      const float lim = h.y * tot;
      const float nl = clampf(l0 + (h.z - u * h.w), -lim, lim);
      const float dl = nl - l0;
      if (dl != 0.f) {
        *word_of(cn, 2 * p.c + 1, 1) = nl;
        seg_axpy_g(rA, p.nA, p.offA == nd, fi, dvA, dl);
        seg_axpy_g(rB, p.nB, p.offB == nd, fi, dvB, dl);
      }
    }
    // ---- lateral friction: the two rows of a contact are solved together (implicit cone)
    if (nc > 0) {
      float4 h0 = at(cf, 0), h1 = at(cf, 1);
      float4 a0 = at(pool, 24), a1 = at(pool, 25), b0 = at(pool, 36), b1 = at(pool, 37);
      float ln = at(cn, 1).x;
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const float4* reg = pool + 24 * 64 * c;
        const int cnx = min(c + 1, SB_MAXCONTACT - 1);
        const float4* nreg = pool + 24 * 64 * cnx;
        const float4 hn0 = at(cf, 2 * cnx), hn1 = at(cf, 2 * cnx + 1);
        const float4 an0 = at(nreg, 24), an1 = at(nreg, 25), bn0 = at(nreg, 36), bn1 = at(nreg, 37);
        const float lnn = at(cn, 2 * cnx + 1).x;
        const float la = h1.z, lb = h1.w;
        const CPk p = unpack_contact(h0.x);
        float* dvA = dv + p.offA * PGS_BLOCK;
        float* dvB = dv + p.offB * PGS_BLOCK;
        const float4* rA1 = reg + 12 * 64;       // row 2 body A (float4 24), body B (30); row 3 body A (36), body B (42)
        const float4* rB1 = reg + 15 * 64;
        const float4* rA2 = reg + 18 * 64;
        const float4* rB2 = reg + 21 * 64;
        const float ua = seg_dot(a0, a1, rA1, dvA, p.nA) + seg_dot_g(rB1, dvB, p.nB);
        const float ub = seg_dot(b0, b1, rA2, dvA, p.nA) + seg_dot_g(rB2, dvB, p.nB);
        float na, nb;
        cone_clamp(h0.y * ln, la + (h0.z - ua * h1.x), lb + (h0.w - ub * h1.y), na, nb);
        const float d1 = na - la, d2 = nb - lb;
        if (d1 != 0.f || d2 != 0.f) *reinterpret_cast<float2*>(word_of(cf, 2 * c + 1, 2)) = make_float2(na, nb);
        if (d1 != 0.f) {
          seg_axpy(a0, a1, rA1, p.nA, p.offA == nd, fi, dvA, d1);
          seg_axpy_g(rB1, p.nB, p.offB == nd, fi, dvB, d1);
        }
        if (d2 != 0.f) {
          seg_axpy(b0, b1, rA2, p.nA, p.offA == nd, fi, dvA, d2);
          seg_axpy_g(rB2, p.nB, p.offB == nd, fi, dvB, d2);
        }
        h0 = hn0; h1 = hn1; a0 = an0; a1 = an1; b0 = bn0; b1 = bn1; ln = lnn;
      }
    }
  }
#pragma unroll 1
  for (int i = 0; i < nv; i++) S.w(4 * Q_DV + i) = dv[i * PGS_BLOCK];
}
