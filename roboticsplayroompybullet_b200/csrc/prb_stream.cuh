// prb_stream.cuh — the split ("stream") step pipeline.
//
// One stepSimulation() substep = two launches:
//   prb_setup_kernel   one WARP per env: integrate the previous substep's solution, then kinematics,
//                      collision detection, mass matrix and its inverse, unconstrained velocities and
//                      the constraint rows of this substep.  The rows are written as COMPACT records
//                      to a per-env record stream in HBM (~2.3 kB per env-substep).  The last launch
//                      of an env step also runs the fused observation / reward write.
//   prb_pgs_kernel     one THREAD per env, one warp per block: the block stages its 32 envs' records
//                      into shared memory ONCE (coalesced 512-byte rows), then runs the 50
//                      projected-Gauss-Seidel sweeps, in velocity space, in
//                      btMultiBodyConstraintSolver::solveSingleIteration order, entirely out of
//                      shared memory: records, accumulated impulses and the velocity change dv are
//                      lane-interleaved float4 columns (conflict-free).  Records past the per-env
//                      stage capacity are read from the stream in place (L2), so capacity only costs
//                      speed, never contacts.
//
// Why: the first thread-per-env solver re-streamed 4.5 kB of explicit Jacobian rows per env per
// iteration from HBM (14.7 GB of DRAM reads per launch at 65536 envs, profiles/r1_v9_ncu.md) and was
// latency-bound on dependent global loads.  95 % of the contacts in the playroom are a free body
// (block, drawer) against static geometry: their three rows are fully described by the contact
// frame (n, t1), the lever arm r and the body's inverse inertia, so the record is 6 float4 instead
// of 15 and the Jacobians are rebuilt in registers.  Arm-side rows keep explicit J and M^-1 J^T.
//
// Reference path: environments.py:485-490 (12 x stepSimulation), Bullet btMultiBodyDynamicsWorld.
#pragma once
#include "prb_kernels.cuh"

// ---- record stream.  Envs are grouped by 32 (one group = the 32 envs a solver warp serves); float4
// number q of env (g, l) lives at float4 index  g * 32 * SB_Q + q * 32 + l.  All offsets are in
// float4 units ("q").
//
// Constraint islands.  Rows that share no dynamic body never exchange data in Gauss-Seidel, so the
// sweep order BETWEEN islands is immaterial (bit-identical results) and islands can be solved
// concurrently.  Bodies are grouped as {arm + door/button/dial}, {free body 0}, {free body 1}; a
// contact between two groups merges them.  Each island is solved by the "slot" of its lowest group:
//   slot 0: joint rows (limits, motors, gear) + every contact of the island that contains the arm
//   slot 1, 2: contacts of a free body (block, drawer) that touches only static geometry (or the
//              other free body) — the common case; these records need no arm data at all.
// Within a slot the contacts keep Bullet's order.
#define Q_HDR 0          // ints {n_jrow, n_contact[0], n_spin[0], t_spin[0]}
                         //      {n_contact[1] | n_spin[1] << 8 | slot of free body 0 << 16 | slot of free body 1 << 18, start[1], t_spin[1], -}
                         //      {n_contact[2] | n_spin[2] << 8, start[2], t_spin[2], -}
#define Q_VSTAR 3        // 8 q: unconstrained velocities v* of the substep (word i = DoF i)
#define Q_DV 11          // 8 q: solver output M^-1 J^T lambda (word i = DoF i)
#define Q_ST 19          // start of the slot regions; region 0 starts here, region s at Q_ST + start[s];
                         // "t" offsets are relative to the start of the slot's region
#define T_BODY 0         // (region 0) 2 q per free body: world inverse inertia {xx xy xz yy} {yz zz 1/m -}
#define T_MINV 4         // (region 0) 12 rows x 3 q: arm inverse mass matrix (row d at T_MINV + 3 d, zero padded)
#define T_JROW 40        // (region 0) n_jrow x 1 q: {packed, rhs, invD, hi}
                         //   packed: pd | pd2 << 8 | a << 16 | a2 << 20 | neg << 24 | sym << 25
                         //   pd/pd2: solver dv word of the DoF (pd2 = 0xff: none); a/a2: arm row (15: not arm)
                         //   neg: J = -e_d; sym: lo = -hi (else lo = 0)
                         // then ceil(n_jrow / 4) q of accumulated impulses, then the slot's contacts and spin list
#define SB_MAXJROW 40
#define SB_MAXCONTACT 32
// contact record (>= 6 q), P = primary side, S = secondary side (see kind_rank):
//   +0 {packed, cfm * invD0, rhs0, invD0}     +1 {n.xyz, lambda0}        +2 {rP.xyz, mu}
//   +3 {t1.xyz, lambda1 (spin)}               +4 {rhs2, rhs3, invD2, invD3}   +5 {lambda2, lambda3, -, -}
//   then the geometry of P (free: none, its lever arm is rP; slide: 1 q {j0 j1 j2 j3}; arm: 4 rows x
//   {3 q J, 3 q B = M^-1 J^T}, rows = normal, spin, friction 1, friction 2) and of S (free: 1 q {r.xyz, -};
//   slide, arm as for P).
//   packed: kP | kS << 2 | iP << 4 | iS << 7 | neg << 10 | spin << 11 | stride << 12
//   (kX: 0 static, 1 free, 2 slide, 3 arm; iX: free / slide index; neg: sign of P is -1; stride: q to the
//   slot's next record — a record never straddles the slot's stage boundary)
// spin list entry (1 q): {t of the contact record, spin coefficient, rhs1, invD1}
#define CT_BASE_Q 6
#define CT_ARM_Q 24
#define SB_Q (Q_ST + T_JROW + SB_MAXJROW + SB_MAXJROW / 4 + SB_MAXCONTACT * (CT_BASE_Q + 2 * CT_ARM_Q) + SB_MAXCONTACT + 3 * 64)   // + stage-boundary gaps
#define SB_PAD_Q 48      // readable slack after the last group (the solvers prefetch one record ahead)
// stage capacities (q of shared memory per env) of the three solver kernels
#ifndef PGS_STAGE_J
#define PGS_STAGE_J 68   // joint-row kernel: region 0 up to the impulses of <= 22 joint rows; 6 blocks per SM
#endif
#ifndef PGS_STAGE_F
#define PGS_STAGE_F 64   // free-body kernel: 10 contact records; 6 blocks per SM
#endif
#ifndef PGS_STAGE_G
#define PGS_STAGE_G 200  // general kernel (islands that contain the arm): 2 blocks per SM
#endif
#define PGS_G_LW 2       // general kernel: an env owns 4 shared-memory columns (8 envs per warp)
#define PGS_ROWS_GA 104  // class A: 416 q per env, 4 blocks per SM
#define PGS_ROWS_GB 216  // class B: 864 q per env, 2 blocks per SM
#define PGS_CLASS_A_MAXQ (PGS_ROWS_GA << PGS_G_LW)
#define PGS_MAXJROW_J ((PGS_STAGE_J - T_JROW) * 4 / 5)     // njr + ceil(njr / 4) <= PGS_STAGE_J - T_JROW
#define PGS_DVQ 8        // general solver dv: arm q 0..2, free body b q 3+2b..4+2b (6 words used), slides q 7
#define DVW_FREE(b) (12 + 8 * (b))
#define DVW_SLIDE(s) (12 + 8 * PRB_MAXFREE + (s))
enum { K_STATIC = 0, K_FREE = 1, K_SLIDE = 2, K_ARM = 3 };

struct SV {              // one env's column of its group
  float4* b;
  PRB_D float4& q(int i) const { return b[i * 32]; }
  PRB_D float& w(int i) const { return reinterpret_cast<float*>(&b[(i >> 2) * 32])[i & 3]; }
};
PRB_D SV sv_of(float* sbuf, int e) {
  SV s;
  s.b = reinterpret_cast<float4*>(sbuf) + (size_t)(e >> 5) * (SB_Q * 32) + (e & 31);
  return s;
}
PRB_HD size_t sbuf_bytes(int64_t N) { return ((size_t)((N + 31) / 32) * 32 * SB_Q + 32 * SB_PAD_Q) * sizeof(float4); }

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

PRB_D int body_kind(const DevModel& M, int body) { return body < 0 ? K_STATIC : (body == 0 ? K_ARM : (body <= M.n_free ? K_FREE : K_SLIDE)); }
PRB_D int kind_rank(int k) { return k == K_ARM ? 3 : (k == K_SLIDE ? 2 : (k == K_FREE ? 1 : 0)); }
PRB_D int geom_q(int kind, bool primary) { return kind == K_ARM ? CT_ARM_Q : (kind == K_SLIDE ? 1 : ((kind == K_FREE && !primary) ? 1 : 0)); }

// One side of one constraint row: J = unit force `dir` at world point pt (or unit torque when angular)
// on the body of collider col, times sign; B = M^-1 J^T.  Returns J.B and accumulates J.v*.  Arm sides
// store J and B (3 q each, stride 32) at garm; slide sides return the scalar J in *jslide; free-body
// sides store nothing (the solver rebuilds them from the contact frame and the lever arm).
template <int ND, class WM>
PRB_D float side_row(const DevModel& M, const WM& W, int col, v3 pt, v3 dir, float sign, bool angular,
                     float4* garm, float* jslide, float* rel) {
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
      garm[k * 32] = make_float4(J[4 * k], J[4 * k + 1], J[4 * k + 2], J[4 * k + 3]);
      garm[(3 + k) * 32] = make_float4(B[4 * k], B[4 * k + 1], B[4 * k + 2], B[4 * k + 3]);
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
    *rel += r;
  } else {
    const int s = body - 1 - M.n_free, o = M.nd + 6 * M.n_free + s;
    v3 a = ld3(M.slide_axis_w[s]);
    float g;
    if (M.slide_jtype[s] == 0) g = angular ? dot(a, dir) : dot(a, cross(pt - ld3(W.sp[s]), dir));
    else g = angular ? 0.f : dot(a, dir);
    const float j = sign * g, bb = j * M.slide_minv[s];
    *jslide = j;
    d = j * bb; *rel += j * W.vs[o];
  }
  return d;
}

PRB_D int dvw_of(const DevModel& M, int d) {     // velocity DoF -> solver dv word
  if (d < M.nd) return d;
  int r = d - M.nd;
  if (r < 6 * M.n_free) return DVW_FREE(r / 6) + r % 6;
  return DVW_SLIDE(r - 6 * M.n_free);
}

// constraint rows of the substep -> record stream
template <int ND, class WM>
PRB_D void phase_rows_stream(const DevModel& M, WM& W, int lane, const SV& S) {
  const float dt = M.params[P_DT], erp = M.params[P_ERP_JOINT], erp2 = M.params[P_ERP_CONTACT];
  const int nd = M.nd;
  // ---- joint rows (lane 0, serial: <= 40 rows of a few flops each): limits, motors, gear
  if (lane == 0) {
    int nr = 0;
#define PRB_PUT_JROW(d_, d2_, neg_, sym_, rhs_, invD_, hi_)                                                \
    do {                                                                                                  \
      if (nr >= SB_MAXJROW) { W.overflow = 1; }                                                            \
      else {                                                                                              \
        const int d__ = (d_), d2__ = (d2_);                                                               \
        const int pk__ = dvw_of(M, d__) | ((d2__ < 0 ? 0xff : dvw_of(M, d2__)) << 8) | ((d__ < nd ? d__ : 15) << 16) | \
                         (((d2__ >= 0 && d2__ < nd) ? d2__ : 15) << 20) | ((neg_) << 24) | ((sym_) << 25); \
        S.q(Q_ST + T_JROW + nr) = make_float4(__int_as_float(pk__), rhs_, invD_, hi_);                    \
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
        PRB_PUT_JROW(i, -1, side, 0, (-pen * e / dt - rel) * invD, invD, M.params[P_LIMIT_MAX_IMPULSE]);
      }
    }
    for (int i = 0; i < nd; i++) {
      if (W.mmaximp[i] <= 0.f) continue;
      float invD = 1.0f / W.Minv[i][i];
      float v = W.vs[i];
      float target_v = W.mkp[i] * (W.mtarget[i] - W.q[i]) / dt + v + M.params[P_MOTOR_KD] * (0.f - v);
      PRB_PUT_JROW(i, -1, 0, 1, (target_v - v) * invD, invD, W.mmaximp[i]);
    }
    for (int s = 0; s < M.n_slide; s++) {
      int o = nd + 6 * M.n_free + s;
      float maximp = M.slide_motor[s][3] < 0 ? M.params[P_DEFAULT_MOTOR_IMPULSE] : M.slide_motor[s][3];
      if (maximp <= 0.f) continue;
      float invD = 1.0f / M.slide_minv[s];
      float v = W.vs[o];
      float target_v = M.slide_motor[s][1] * (M.slide_motor[s][0] - W.sq[s]) / dt + v + M.slide_motor[s][2] * (0.f - v);
      PRB_PUT_JROW(o, -1, 0, 1, (target_v - v) * invD, invD, maximp);
    }
    if (M.gear_a >= 0) {
      int a = M.gear_a, b = M.gear_b;
      float r = M.params[P_GEAR_RATIO];
      float D = W.Minv[a][a] + 2.f * r * W.Minv[a][b] + r * r * W.Minv[b][b];
      float invD = 1.0f / D;
      float rel = W.vs[a] + r * W.vs[b];
      PRB_PUT_JROW(a, b, 0, 1, (-rel * M.params[P_GEAR_ERP]) * invD, invD, M.params[P_GEAR_MAX_IMPULSE]);
    }
#undef PRB_PUT_JROW
    W.n_jrow = nr;
  }
  // ---- arm inverse mass matrix (lane = row), zero padded to 12 columns; free-body table
  if (lane < ND) {
    float r[12];
#pragma unroll
    for (int j = 0; j < 12; j++) r[j] = j < ND ? W.Minv[lane][j] : 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) S.q(Q_ST + T_MINV + 3 * lane + k) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
  if (lane >= 16 && lane - 16 < M.n_free) {
    const int b = lane - 16;
    S.q(Q_ST + T_BODY + 2 * b) = make_float4(W.fIinv[b][0], W.fIinv[b][1], W.fIinv[b][2], W.fIinv[b][3]);
    S.q(Q_ST + T_BODY + 2 * b + 1) = make_float4(W.fIinv[b][4], W.fIinv[b][5], 1.0f / M.free_mass[b], 0.f);
  }
  __syncwarp();
  // ---- contact records: lane = contact
  const int nc = W.n_contact, njr = W.n_jrow;
  int colP = 0, colS = 0, kP = K_STATIC, kS = K_STATIC, size = 0, grpP = 0, grpS = -1;
  bool swapped = false;
  Contact c;
  if (lane < nc) {
    c = W.ct[lane];
    const int ca = c.cols & 0xff, cb = (c.cols >> 8) & 0xff;
    const int kA = body_kind(M, col_dyn_body(M, ca)), kB = body_kind(M, col_dyn_body(M, cb));
    swapped = kind_rank(kB) > kind_rank(kA);
    colP = swapped ? cb : ca; colS = swapped ? ca : cb;
    kP = swapped ? kB : kA; kS = swapped ? kA : kB;
    size = CT_BASE_Q + geom_q(kP, true) + geom_q(kS, false);
    grpP = kP == K_FREE ? M.col_body[colP] : 0;                 // 0: arm + slide bodies, 1 + b: free body b
    grpS = kS == K_STATIC ? -1 : (kS == K_FREE ? M.col_body[colS] : 0);
  }
  // islands over the three groups -> slot of each free body and of each contact
  int slotf[PRB_MAXFREE];
  {
    const bool two = lane < nc && grpS >= 0 && grpS != grpP;
    const int lo_ = min(grpP, grpS), hi_ = max(grpP, grpS);
    const bool m01 = __any_sync(FULL, two && lo_ == 0 && hi_ == 1);
    const bool m02 = __any_sync(FULL, two && lo_ == 0 && hi_ == 2);
    const bool m12 = __any_sync(FULL, two && lo_ == 1 && hi_ == 2);
    const bool c01 = m01 || (m12 && m02), c02 = m02 || (m12 && m01), c12 = m12 || (m01 && m02);
    slotf[0] = c01 ? 0 : 1;
    slotf[1] = c02 ? 0 : (c12 ? 1 : 2);
  }
  const int slot = lane < nc ? (grpP == 0 ? 0 : slotf[grpP - 1]) : -1;
  // per-slot placement: region 0 = fixed part + slot-0 contacts + spin list, then regions 1 and 2
  const int t_ct0 = T_JROW + njr + ((njr + 3) >> 2);
  bool has_spin = false;
  float spin = 0.f;
  if (lane < nc) {
    const int ca = c.cols & 0xff, cb = (c.cols >> 8) & 0xff;
    spin = M.col_spin[ca] * M.col_fric[ca] + M.col_spin[cb] * M.col_fric[cb];
    has_spin = spin > 0.f;
  }
  int t = 0, stride = size, t_spin_mine = 0, spin_rank = 0, region_mine = 0;
  int ncs[3], nss[3], tsp[3], start[3];
  {
    int region = 0;                                   // start of the slot's region relative to Q_ST
#pragma unroll
    for (int sidx = 0; sidx < 3; sidx++) {
      const int stage = sidx == 0 ? PGS_STAGE_G : PGS_STAGE_F;
      const bool mine = slot == sidx;
      const unsigned mask = __ballot_sync(FULL, mine);
      const unsigned smask = __ballot_sync(FULL, mine && has_spin);
      int total;
      int ts = (sidx == 0 ? t_ct0 : 0) + warp_excl_scan(mine ? size : 0, lane, &total);
      const bool straddle = mine && ts < stage && ts + size > stage;
      const unsigned sm = __ballot_sync(FULL, straddle);
      int shift = 0;
      if (sm) {
        const int sl_ = __ffs((int)sm) - 1;
        shift = stage - __shfl_sync(FULL, ts, sl_);
        if (lane >= sl_) ts += shift;
      }
      const unsigned later = mask & ~((2u << lane) - 1u);          // lanes of this slot after me (lane 31: none)
      const int nl = later ? __ffs((int)later) - 1 : lane;
      const int tnext = __shfl_sync(FULL, ts, nl);
      const int tend = (sidx == 0 ? t_ct0 : 0) + total + shift;
      if (mine) {
        t = ts; stride = (lane < 31 && later) ? tnext - ts : size;
        t_spin_mine = tend; spin_rank = __popc(smask & ((1u << lane) - 1u)); region_mine = region;
      }
      ncs[sidx] = __popc(mask); nss[sidx] = __popc(smask); tsp[sidx] = tend; start[sidx] = region;
      region += tend + nss[sidx];
    }
    if (lane == 0) W.dbg_p = region;
  }
  float rhs[4] = {0.f, 0.f, 0.f, 0.f}, invDs[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < nc) {
    const int ca = c.cols & 0xff, cb = (c.cols >> 8) & 0xff;
    v3 n = V3(c.nx, c.ny, c.nz), pb = V3(c.pbx, c.pby, c.pbz), pa = pb + n * c.dist;
    const v3 pP = swapped ? pb : pa, pS = swapped ? pa : pb;
    const float sP = swapped ? -1.0f : 1.0f;
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
    const float mu = clampf(M.col_fric[ca] * M.col_fric[cb], -10.f, 10.f);
    v3 t1, t2;
    plane_space(n, t1, t2);
    float cfms = 0.f;
    float4* rec = &S.q(Q_ST + region_mine + t);
    float4* gP = rec + CT_BASE_Q * 32;
    float4* gS = gP + geom_q(kP, true) * 32;
    float jP[4] = {0.f, 0.f, 0.f, 0.f}, jS[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      v3 dir = k == 0 ? n : (k == 1 ? n : (k == 2 ? t1 : t2));
      const bool ang = (k == 1);
      if (k == 1 && !has_spin) continue;      // no torsional row: never visited by the solver
      float rel = 0.f, D = 0.f;
      D += side_row<ND>(M, W, colP, pP, dir, sP, ang, gP + 6 * k * 32, &jP[k], &rel);
      if (kS != K_STATIC) D += side_row<ND>(M, W, colS, pS, dir, -sP, ang, gS + 6 * k * 32, &jS[k], &rel);
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
    const int bP = M.col_body[colP], bS = M.col_body[colS];
    const int iP = kP == K_FREE ? bP - 1 : (kP == K_SLIDE ? bP - 1 - M.n_free : 0);
    const int iS = kS == K_FREE ? bS - 1 : (kS == K_SLIDE ? bS - 1 - M.n_free : 0);
    v3 rP = V3(0, 0, 0);
    if (kP == K_FREE) rP = pP - ld3(W.fpos[iP]);
    if (kP == K_SLIDE) gP[0] = make_float4(jP[0], jP[1], jP[2], jP[3]);
    if (kS == K_FREE) { v3 rS = pS - ld3(W.fpos[iS]); gS[0] = make_float4(rS.x, rS.y, rS.z, 0.f); }
    if (kS == K_SLIDE) gS[0] = make_float4(jS[0], jS[1], jS[2], jS[3]);
    const int packed = kP | (kS << 2) | (iP << 4) | (iS << 7) | ((swapped ? 1 : 0) << 10) | ((has_spin ? 1 : 0) << 11) | (stride << 12);
    rec[0] = make_float4(__int_as_float(packed), cfms, rhs[0], invDs[0]);
    rec[32] = make_float4(n.x, n.y, n.z, 0.f);
    rec[64] = make_float4(rP.x, rP.y, rP.z, mu);
    rec[96] = make_float4(t1.x, t1.y, t1.z, 0.f);
    rec[128] = make_float4(rhs[2], rhs[3], invDs[2], invDs[3]);
    rec[160] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_spin) S.q(Q_ST + region_mine + t_spin_mine + spin_rank) = make_float4(__int_as_float(t), spin, rhs[1], invDs[1]);
  }
  __syncwarp();
  if (lane == 0) {
    S.q(Q_HDR) = make_float4(__int_as_float(njr), __int_as_float(ncs[0]), __int_as_float(nss[0]), __int_as_float(tsp[0]));
    S.q(Q_HDR + 1) = make_float4(__int_as_float(ncs[1] | (nss[1] << 8) | (slotf[0] << 16) | (slotf[1] << 18)), __int_as_float(start[1]),
                                 __int_as_float(tsp[1]), 0.f);
    S.q(Q_HDR + 2) = make_float4(__int_as_float(ncs[2] | (nss[2] << 8)), __int_as_float(start[2]), __int_as_float(tsp[2]), 0.f);
    if (nc > W.dbg_c) W.dbg_c = nc;
    // island of the arm needs the general solver: size class by the q count of region 0
    W.dbg_u = (ncs[0] > 0 || njr > PGS_MAXJROW_J) ? ((tsp[0] + nss[0] <= PGS_CLASS_A_MAXQ) ? 1 : 2) : 0;
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
                                                                         float* __restrict__ sbuf, DevOut O, int N, int flags,
                                                                         int* __restrict__ heavy_list, int* __restrict__ heavy_cnt) {
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
    if (lane == 0 && W.dbg_u) {                      // order within a list is immaterial: envs are independent
      const int cls = W.dbg_u - 1;                   // 0: region 0 fits class A's stage, 1: larger
      heavy_list[(size_t)cls * N + atomicAdd(heavy_cnt + cls, 1)] = e;
    }
    if (lane == 0 && O.dbg) { O.dbg[4 * e] = 0; O.dbg[4 * e + 1] = W.dbg_c; O.dbg[4 * e + 2] = W.dbg_p; O.dbg[4 * e + 3] = W.n_jrow; }
  }
  if (flags & SETUP_OBSERVE) phase_observe(M, W, lane, O, (size_t)e, true);
  __syncwarp();
  if (flags & (SETUP_INTEGRATE | SETUP_OBSERVE)) store_state(M, W, st, lane);
}

// ============================================================================ PGS kernels (thread per env-island)
// Three kernels, all one warp per block, records staged once into shared memory and swept 50 times there:
//   prb_pgs_joint_kernel   slot 0 of the envs whose arm island has no contacts: joint rows only (32 envs / warp)
//   prb_pgs_free_kernel    slots 1 and 2 (blockIdx.y): free-body islands, compact records only (32 envs / warp)
//   prb_pgs_kernel         slot 0 of the envs on a "heavy" list: joint rows + every record kind.  The arm
//                          island's records are large (explicit 12-wide J and M^-1 J^T per row), so an
//                          env owns FOUR adjacent shared-memory columns here (8 envs / warp): q number t
//                          sits in row t >> 2, column t & 3.  Two size classes, two launches.
#define PGS_BLOCK 32
#define PGS_J_DVQ 4      // arm q 0..2, slides q 3
#define PGS_F_TAILQ 8    // free-body kernel: dv 4 q (body b at 2b, 2b+1) + body table 4 q
#define PGS_G_EPW (32 >> PGS_G_LW)                      // envs per warp
#define PGS_SMEM_J ((PGS_STAGE_J + PGS_J_DVQ) * 32 * 16)
#define PGS_SMEM_F ((PGS_STAGE_F + PGS_F_TAILQ) * 32 * 16)
#define PGS_SMEM_G(rows) (((rows) + (PGS_DVQ >> PGS_G_LW)) * 32 * 16)

#ifdef PRB_EMU
static float4 g_emu_pgs_smem[(PGS_ROWS_GB + 2) * 32 + (PGS_STAGE_J + PGS_STAGE_F + 16) * 32];
#define PRB_PGS_SMEM_DECL float4* sm = g_emu_pgs_smem
#else
#define PRB_PGS_SMEM_DECL extern __shared__ __align__(16) float4 prb_pgs_smem[]; float4* sm = prb_pgs_smem
#endif

PRB_D float f4comp(const float4& a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? a.z : a.w)); }
PRB_D float dot4(const float4& a, const float4& b, float s) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s)))); }
PRB_D void axpy4(float4& y, const float4& b, float a) { y.x = fmaf(b.x, a, y.x); y.y = fmaf(b.y, a, y.y); y.z = fmaf(b.z, a, y.z); y.w = fmaf(b.w, a, y.w); }

// Where an env-island's q number t lives: LW = 0: one column (row t); LW = 2: four columns.  q numbers
// at or past `cap` were not staged and are read from the stream in place.
template <int LW>
struct PgsMem {
  float4* sl;              // first column of this env in shared memory
  float4* Gr;              // this env's column of the slot's region in the stream
  int cap;
  PRB_D float4& s(int t) const { return sl[(t >> LW) * 32 + (t & ((1 << LW) - 1))]; }       // always staged
  PRB_D float4& g(int t) const { return *(t < cap ? &s(t) : Gr + t * 32); }
  PRB_D float& sw(int t0, int w) const { return reinterpret_cast<float*>(&s(t0 + (w >> 2)))[w & 3]; }   // word w of the q array at t0
};

// one side of a contact inside the solver
struct PSide { int kind, idx, gt; float sgn; v3 r; };       // gt: t of the side's geometry
// the dv of that side's body, loaded once per row visit
struct PVel { v3 v, w, pv; float4 a0, a1, a2; float sl; };

// The contact sweeps.  LIGHT: only free-body sides exist (slots 1, 2); dv q array at t = dvt: general
// layout arm 0..2, free body b at 3 + 2 b, slides 7; LIGHT layout free body b at 2 b.
template <bool LIGHT, int LW>
struct PgsSweep {
  static constexpr int FQ0 = LIGHT ? 0 : 3;
  static constexpr int SLQ = 3 + 2 * PRB_MAXFREE;
  PgsMem<LW> m;
  int dvt;                 // t of the dv array (past the records, always staged)
  int bodyt;               // t of the free-body table (always staged)
  const DevModel* M;

  PRB_D void load(const PSide& s, PVel& V) const {
    if (LIGHT || s.kind == K_FREE) {
      const float4 x = m.s(dvt + FQ0 + 2 * s.idx), y = m.s(dvt + FQ0 + 1 + 2 * s.idx);
      V.v = V3(x.x, x.y, x.z); V.w = V3(x.w, y.x, y.y);
      V.pv = V.v + cross(V.w, s.r);
    } else if (s.kind == K_ARM) {
      V.a0 = m.s(dvt); V.a1 = m.s(dvt + 1); V.a2 = m.s(dvt + 2);
    } else if (s.kind == K_SLIDE) {
      V.sl = m.sw(dvt + SLQ, s.idx);
    }
  }
  // J_k . dv of one side (row k: 0 normal, 1 spin, 2 / 3 friction; d: the row's direction)
  PRB_D float jdot(const PSide& s, const PVel& V, v3 d, int k, bool ang) const {
    if (LIGHT || s.kind == K_FREE) return s.sgn * (ang ? dot(d, V.w) : dot(d, V.pv));
    if (s.kind == K_ARM) {
      const float4 j0 = m.g(s.gt + 6 * k), j1 = m.g(s.gt + 6 * k + 1), j2 = m.g(s.gt + 6 * k + 2);
      return dot4(j0, V.a0, 0.f) + dot4(j1, V.a1, dot4(j2, V.a2, 0.f));
    }
    if (s.kind == K_SLIDE) return f4comp(m.g(s.gt), k) * V.sl;
    return 0.f;
  }
  // dv += B_k dl (+ B_k2 dl2): P = sum of direction * impulse (linear rows) or the angular impulse (spin)
  template <bool PAIR>
  PRB_D void apply(const PSide& s, PVel& V, v3 P, bool ang, int k, float dl, int k2, float dl2) const {
    if (LIGHT || s.kind == K_FREE) {
      const float4 i0 = m.s(bodyt + 2 * s.idx), i1 = m.s(bodyt + 2 * s.idx + 1);
      const float I[6] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y};
      const v3 Ps = P * s.sgn;
      v3 v = V.v, w = V.w;
      if (ang) w = w + symmul(I, Ps);
      else { v = v + Ps * i1.z; w = w + symmul(I, cross(s.r, Ps)); }
      m.s(dvt + FQ0 + 2 * s.idx) = make_float4(v.x, v.y, v.z, w.x);
      m.s(dvt + FQ0 + 1 + 2 * s.idx) = make_float4(w.y, w.z, 0.f, 0.f);
    } else if (s.kind == K_ARM) {
      float4 a0 = V.a0, a1 = V.a1, a2 = V.a2;
      axpy4(a0, m.g(s.gt + 6 * k + 3), dl); axpy4(a1, m.g(s.gt + 6 * k + 4), dl); axpy4(a2, m.g(s.gt + 6 * k + 5), dl);
      if (PAIR) { axpy4(a0, m.g(s.gt + 6 * k2 + 3), dl2); axpy4(a1, m.g(s.gt + 6 * k2 + 4), dl2); axpy4(a2, m.g(s.gt + 6 * k2 + 5), dl2); }
      m.s(dvt) = a0; m.s(dvt + 1) = a1; m.s(dvt + 2) = a2;
    } else if (s.kind == K_SLIDE) {
      const float4 j = m.g(s.gt);
      float tt = f4comp(j, k) * dl;
      if (PAIR) tt = fmaf(f4comp(j, k2), dl2, tt);
      m.sw(dvt + SLQ, s.idx) = fmaf(tt, M->slide_minv[s.idx], V.sl);
    }
  }
  PRB_D void sides_of(int pk, int t, const float4& q2, PSide& P, PSide& Sd) const {
    P.kind = pk & 3; Sd.kind = (pk >> 2) & 3;
    P.idx = (pk >> 4) & 7; Sd.idx = (pk >> 7) & 7;
    P.sgn = ((pk >> 10) & 1) ? -1.0f : 1.0f; Sd.sgn = -P.sgn;
    P.r = V3(q2.x, q2.y, q2.z);
    P.gt = t + CT_BASE_Q;
    Sd.gt = LIGHT ? P.gt : P.gt + geom_q(P.kind, true);
    Sd.r = V3(0, 0, 0);
    if (Sd.kind == K_FREE) { const float4 g = m.g(Sd.gt); Sd.r = V3(g.x, g.y, g.z); }
  }
  PRB_D bool same_body(const PSide& P, const PSide& Sd) const { return !LIGHT && Sd.kind == K_ARM && P.kind == K_ARM; }

  // ---- contact normals
  PRB_D void normals(int t_ct, int nc) const {
    int t = t_ct;
    float4 q0 = m.g(t), q1 = m.g(t + 1), q2 = m.g(t + 2);
#pragma unroll 1
    for (int c = 0; c < nc; c++) {
      const int pk = __float_as_int(q0.x);
      const int tn = t + ((pk >> 12) & 255);
      const float4 n0 = m.g(tn), n1 = m.g(tn + 1), n2 = m.g(tn + 2);     // next record (readable slack after the last)
      PSide P, Sd;
      sides_of(pk, t, q2, P, Sd);
      const v3 n = V3(q1.x, q1.y, q1.z);
      PVel VP, VS;
      load(P, VP);
      float u = jdot(P, VP, n, 0, false);
      if (Sd.kind != K_STATIC) { load(Sd, VS); u += jdot(Sd, VS, n, 0, false); }
      const float l0 = q1.w;
      const float nl = fmaxf(l0 + (q0.z - l0 * q0.y - u * q0.w), 0.f);
      const float dl = nl - l0;
      if (dl != 0.f) {
        m.g(t + 1).w = nl;
        apply<false>(P, VP, n * dl, false, 0, dl, 0, 0.f);
        if (Sd.kind != K_STATIC) {
          if (same_body(P, Sd)) load(Sd, VS);       // same body: see P's update
          apply<false>(Sd, VS, n * dl, false, 0, dl, 0, 0.f);
        }
      }
      t = tn; q0 = n0; q1 = n1; q2 = n2;
    }
  }
  // ---- spinning friction (Bullet skips the row while the normal impulse is 0)
  PRB_D void spins(int t_spin, int ns) const {
    float4 h = m.g(t_spin);
#pragma unroll 1
    for (int i = 0; i < ns; i++) {
      const float4 hc = h;
      h = m.g(t_spin + i + 1);                                      // next entry (readable slack after the last)
      const int t = __float_as_int(hc.x);
      const float4 q0 = m.g(t), q1 = m.g(t + 1), q2 = m.g(t + 2);
      const float tot = q1.w;
      if (!(tot > 0.f)) continue;
      PSide P, Sd;
      sides_of(__float_as_int(q0.x), t, q2, P, Sd);
      const v3 n = V3(q1.x, q1.y, q1.z);
      PVel VP, VS;
      load(P, VP);
      float u = jdot(P, VP, n, 1, true);
      if (Sd.kind != K_STATIC) { load(Sd, VS); u += jdot(Sd, VS, n, 1, true); }
      const float lim = hc.y * tot;
      float& lam1 = m.g(t + 3).w;
      const float l0 = lam1;
      const float nl = clampf(l0 + (hc.z - u * hc.w), -lim, lim);
      const float dl = nl - l0;
      if (dl != 0.f) {
        lam1 = nl;
        apply<false>(P, VP, n * dl, true, 1, dl, 1, 0.f);
        if (Sd.kind != K_STATIC) {
          if (same_body(P, Sd)) load(Sd, VS);
          apply<false>(Sd, VS, n * dl, true, 1, dl, 1, 0.f);
        }
      }
    }
  }
  // ---- lateral friction: the two rows of a contact are solved together (implicit cone)
  PRB_D void frictions(int t_ct, int nc) const {
    int t = t_ct;
    float4 q0 = m.g(t), q1 = m.g(t + 1), q2 = m.g(t + 2), q3 = m.g(t + 3), q4 = m.g(t + 4), q5 = m.g(t + 5);
#pragma unroll 1
    for (int c = 0; c < nc; c++) {
      const int pk = __float_as_int(q0.x);
      const int tn = t + ((pk >> 12) & 255);
      const float4 n0 = m.g(tn), n1 = m.g(tn + 1), n2 = m.g(tn + 2), n3 = m.g(tn + 3), n4 = m.g(tn + 4), n5 = m.g(tn + 5);
      PSide P, Sd;
      sides_of(pk, t, q2, P, Sd);
      const v3 n = V3(q1.x, q1.y, q1.z), t1 = V3(q3.x, q3.y, q3.z), t2 = cross(n, t1);
      PVel VP, VS;
      load(P, VP);
      float ua = jdot(P, VP, t1, 2, false), ub = jdot(P, VP, t2, 3, false);
      if (Sd.kind != K_STATIC) {
        load(Sd, VS);
        ua += jdot(Sd, VS, t1, 2, false); ub += jdot(Sd, VS, t2, 3, false);
      }
      const float lim = q2.w * q1.w;
      const float la = q5.x, lb = q5.y;
      const float sumA = la + (q4.x - ua * q4.z);
      const float sumB = lb + (q4.y - ub * q4.w);
      float na = sumA, nb = sumB;
      if (sumA < -lim || sumA > lim || sumB < -lim || sumB > lim) {
        const float ss = sumA * sumA + sumB * sumB;
        const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
        const float ca_ = fabsf(lim * sumA) * inv, cb_ = ss > 0.f ? fabsf(lim * sumB) * inv : fabsf(lim);
        na = clampf(sumA, -ca_, ca_); nb = clampf(sumB, -cb_, cb_);
      }
      const float d1 = na - la, d2 = nb - lb;
      if (d1 != 0.f || d2 != 0.f) {
        m.g(t + 5) = make_float4(na, nb, 0.f, 0.f);
        const v3 Pv = t1 * d1 + t2 * d2;
        apply<true>(P, VP, Pv, false, 2, d1, 3, d2);
        if (Sd.kind != K_STATIC) {
          if (same_body(P, Sd)) load(Sd, VS);
          apply<true>(Sd, VS, Pv, false, 2, d1, 3, d2);
        }
      }
      t = tn; q0 = n0; q1 = n1; q2 = n2; q3 = n3; q4 = n4; q5 = n5;
    }
  }
};

// ---- non-contact rows (limits, motors, gear), sweep direction alternating per iteration.
// region 0 staged in m; dv q array at t = dvt: arm at q 0..2, slide DoFs at q SLQ
template <int ND, int SLQ, int LW>
PRB_D void pgs_joint_rows(const DevModel& M, const PgsMem<LW>& m, int dvt, int njr, int it, float ratio) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t_jlam = T_JROW + njr;
#pragma unroll 1
  for (int i = 0; i < njr; i++) {
    const int j = (it & 1) ? i : njr - 1 - i;
    const float4 r = m.s(T_JROW + j);
    const int pk = __float_as_int(r.x);
    int pd = pk & 0xff, pd2 = (pk >> 8) & 0xff;
    const int a = (pk >> 16) & 15, a2 = (pk >> 20) & 15;
    const int sidx = pd - DVW_SLIDE(0);                              // slide index when a == 15
    if (a == 15) pd = 4 * SLQ + sidx;
    const float sg = ((pk >> 24) & 1) ? -1.0f : 1.0f;
    // the M^-1 rows are needed only if the impulse changes; issue their loads now anyway
    float4 m0[3], m1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      m0[k] = z4; m1[k] = z4;
      if (a != 15) m0[k] = m.s(T_MINV + 3 * a + k);
      if (a2 != 15) m1[k] = m.s(T_MINV + 3 * a2 + k);
    }
    float& ru = m.sw(dvt, pd);
    float u = ru;
    if (pd2 != 0xff) u = fmaf(ratio, m.sw(dvt, pd2), u);
    u *= sg;
    float& rl = m.sw(t_jlam, j);
    const float l0 = rl;
    const float hi = r.w, lo = ((pk >> 25) & 1) ? -hi : 0.f;
    const float nl = clampf(l0 + (r.y - u * r.z), lo, hi);
    const float dl = (nl - l0) * sg;
    if (dl != 0.f) {
      rl = nl;
      if (a != 15) {
        const float dl2 = dl * ratio;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (4 * k < ND) {
            float4 y = m.s(dvt + k);
            axpy4(y, m0[k], dl);
            axpy4(y, m1[k], dl2);
            m.s(dvt + k) = y;
          }
        }
      } else {
        ru = fmaf(M.slide_minv[sidx], dl, ru);
      }
    }
  }
}

// solver dv words of the arm and the slide bodies -> stream (linear DoF order)
template <int LW>
PRB_D void pgs_store_arm(const DevModel& M, float4* G, const PgsMem<LW>& m, int dvt, int slq) {
  float* gd = reinterpret_cast<float*>(G + Q_DV * 32);
  const int nd = M.nd, o = nd + 6 * M.n_free;
#pragma unroll 1
  for (int i = 0; i < nd; i++) gd[(i >> 2) * 128 + (i & 3)] = m.sw(dvt, i);
#pragma unroll 1
  for (int s = 0; s < M.n_slide; s++) gd[((o + s) >> 2) * 128 + ((o + s) & 3)] = m.sw(dvt + slq, s);
}
template <int LW>
PRB_D void pgs_store_free(const DevModel& M, float4* G, const PgsMem<LW>& m, int t0, int b) {
  float* gd = reinterpret_cast<float*>(G + Q_DV * 32);
  const float4 x = m.s(t0 + 2 * b), y = m.s(t0 + 2 * b + 1);
  const float v[6] = {x.x, x.y, x.z, x.w, y.x, y.y};
  const int o = M.nd + 6 * b;
#pragma unroll
  for (int k = 0; k < 6; k++) gd[((o + k) >> 2) * 128 + ((o + k) & 3)] = v[k];
}
PRB_D float4* stream_col(float* sbuf, int e) { return reinterpret_cast<float4*>(sbuf) + (size_t)(e >> 5) * (SB_Q * 32) + (e & 31); }

// ---- slot 0, arm island without contacts: joint rows only
template <int ND>
__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_joint_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N) {
  PRB_PGS_SMEM_DECL;
  const int lane = threadIdx.x;
  const int e = blockIdx.x * PGS_BLOCK + lane;
  if (e >= N) return;
  const DevModel& M = *Mp;
  float4* G = stream_col(sbuf, e);
  const float4 hdr = G[Q_HDR * 32];
  const int njr = __float_as_int(hdr.x), nc0 = __float_as_int(hdr.y);
  if (nc0 > 0 || njr > PGS_MAXJROW_J) return;            // on a heavy list: prb_pgs_kernel solves it
  PgsMem<0> m;
  m.sl = sm + lane; m.Gr = G + Q_ST * 32; m.cap = PGS_STAGE_J;
  const int dvt = PGS_STAGE_J;
  {
    const int tq = T_JROW + njr;
#pragma unroll 8
    for (int q = T_MINV; q < tq; q++) m.s(q) = m.Gr[q * 32];
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < PGS_J_DVQ; i++) m.s(dvt + i) = z4;
  for (int i = 0; i < ((njr + 3) >> 2); i++) m.s(T_JROW + njr + i) = z4;
  const float ratio = M.params[P_GEAR_RATIO];
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) pgs_joint_rows<ND, 3, 0>(M, m, dvt, njr, it, ratio);
  pgs_store_arm(M, G, m, dvt, 3);
}

// ---- slots 1 and 2 (blockIdx.y + 1): islands of free bodies against static geometry / each other
__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_free_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N) {
  PRB_PGS_SMEM_DECL;
  const int lane = threadIdx.x;
  const int e = blockIdx.x * PGS_BLOCK + lane;
  const int slot = blockIdx.y + 1;
  if (e >= N) return;
  const DevModel& M = *Mp;
  float4* G = stream_col(sbuf, e);
  const float4 h1 = G[(Q_HDR + 1) * 32], hs = G[(Q_HDR + slot) * 32];
  const int info = __float_as_int(h1.x);
  const int cnt = __float_as_int(hs.x);
  const int nc = cnt & 0xff, ns = (cnt >> 8) & 0xff, start = __float_as_int(hs.y), t_spin = __float_as_int(hs.z);
  bool owner = false;
  for (int b = 0; b < M.n_free; b++) owner = owner || ((info >> (16 + 2 * b)) & 3) == slot;
  if (!owner) return;                                    // merged into another island
  PgsSweep<true, 0> sw;
  sw.m.sl = sm + lane; sw.m.Gr = G + (Q_ST + start) * 32; sw.m.cap = PGS_STAGE_F;
  sw.dvt = PGS_STAGE_F; sw.bodyt = PGS_STAGE_F + 4; sw.M = Mp;
  {
    const int tq = min(t_spin + ns, PGS_STAGE_F);
#pragma unroll 8
    for (int q = 0; q < tq; q++) sw.m.s(q) = sw.m.Gr[q * 32];
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; i++) { sw.m.s(sw.dvt + i) = z4; sw.m.s(sw.bodyt + i) = G[(Q_ST + T_BODY + i) * 32]; }
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    sw.normals(0, nc);
    sw.spins(t_spin, ns);
    sw.frictions(0, nc);
  }
  for (int b = 0; b < M.n_free; b++)
    if (((info >> (16 + 2 * b)) & 3) == slot) pgs_store_free(M, G, sw.m, sw.dvt, b);
}

// ---- slot 0 of the envs whose arm island has contacts (a heavy list): joint rows + all record kinds.
// 8 envs per warp: lanes 4k..4k+3 stage the four columns of env k, lane 4k solves it.
template <int ND>
__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf,
                                                            const int* __restrict__ heavy_list, const int* __restrict__ heavy_cnt, int rows) {
  PRB_PGS_SMEM_DECL;
  const int lane = threadIdx.x, col = lane & ((1 << PGS_G_LW) - 1);
  const int i = blockIdx.x * PGS_G_EPW + (lane >> PGS_G_LW);
  const int cnt = *heavy_cnt;
  if (blockIdx.x * PGS_G_EPW >= cnt) return;             // whole warp
  const bool have = i < cnt;
  const int e = have ? heavy_list[i] : 0;
  const DevModel& M = *Mp;
  float4* G = stream_col(sbuf, e);
  int njr = 0, nc = 0, ns = 0, t_spin = 0, info = 0;
  if (have) {
    const float4 hdr = G[Q_HDR * 32], h1 = G[(Q_HDR + 1) * 32];
    njr = __float_as_int(hdr.x); nc = __float_as_int(hdr.y); ns = __float_as_int(hdr.z); t_spin = __float_as_int(hdr.w);
    info = __float_as_int(h1.x);
  }
  PgsSweep<false, PGS_G_LW> sw;
  sw.m.sl = sm + (lane - col); sw.m.Gr = G + Q_ST * 32; sw.m.cap = rows << PGS_G_LW;
  sw.dvt = rows << PGS_G_LW; sw.bodyt = T_BODY; sw.M = Mp;
  {
    const int tq = min(t_spin + ns, sw.m.cap);
#pragma unroll 4
    for (int q = col; q < tq; q += (1 << PGS_G_LW)) sw.m.s(q) = sw.m.Gr[q * 32];
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col == 0) {
#pragma unroll
    for (int k = 0; k < PGS_DVQ; k++) sw.m.s(sw.dvt + k) = z4;
  }
  __syncwarp();
  if (col != 0 || !have) return;
  for (int k = 0; k < ((njr + 3) >> 2); k++) sw.m.s(T_JROW + njr + k) = z4;
  const int t_ct = T_JROW + njr + ((njr + 3) >> 2);
  const float ratio = M.params[P_GEAR_RATIO];
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    pgs_joint_rows<ND, 3 + 2 * PRB_MAXFREE, PGS_G_LW>(M, sw.m, sw.dvt, njr, it, ratio);
    sw.normals(t_ct, nc);
    sw.spins(t_spin, ns);
    sw.frictions(t_ct, nc);
  }
  pgs_store_arm(M, G, sw.m, sw.dvt, 3 + 2 * PRB_MAXFREE);
  for (int b = 0; b < M.n_free; b++)
    if (((info >> (16 + 2 * b)) & 3) == 0) pgs_store_free(M, G, sw.m, sw.dvt + 3, b);
}
