// prb_stream.cuh — the split ("stream") step pipeline.
//
// One stepSimulation() substep = one setup launch + the constraint solve:
//   prb_setup_kernel   one WARP per env: integrate the previous substep's solution, then kinematics,
//                      collision detection, mass matrix and its inverse, unconstrained velocities and
//                      the constraint rows of this substep, written as records to a per-env record
//                      stream in HBM (~2-3 kB per env-substep).  The last launch of an env step also
//                      runs the fused observation / reward write.
//   solver kernels     the 50 projected-Gauss-Seidel sweeps, in velocity space, in
//                      btMultiBodyConstraintSolver::solveSingleIteration order.  Every kernel stages its
//                      envs' records ONCE into shared memory (one warp per block, lane-interleaved
//                      float4 columns, conflict-free) and sweeps them there; the constraint islands of
//                      an env are solved by different kernels concurrently (see "Constraint islands").
//
// Why (profiles/r1_v9_ncu.md): the first thread-per-env solver re-streamed 4.5 kB of explicit Jacobian
// rows per env per iteration from HBM (14.7 GB of DRAM reads per launch at 65536 envs) and was
// latency-bound on dependent global loads.  95 % of the contacts in the playroom are a free body
// (block, drawer) against static geometry: their three rows are fully described by the contact frame
// (n, t1), the lever arm r and the body's inverse inertia, so the record is 6 float4 instead of 15 and
// the Jacobians are rebuilt in registers.  Rows that involve the arm keep explicit J and M^-1 J^T and are
// solved by four lanes per env.
//
// Reference path: environments.py:485-490 (12 x stepSimulation), Bullet btMultiBodyDynamicsWorld.
#pragma once
#include "prb_kernels.cuh"

// ---- record stream.  Envs are grouped by 32; float4 number q of env (g, l) lives at float4 index
// g * 32 * SB_Q + q * 32 + l.  All offsets are in float4 units ("q").
//
// Constraint islands.  Rows that share no dynamic body never exchange data in Gauss-Seidel, so the
// sweep order BETWEEN islands is immaterial (bit-identical results) and islands can be solved
// concurrently.  Bodies are grouped as {arm + door/button/dial}, {free body 0}, {free body 1}; a
// contact between two groups merges them.  Each island is solved by the "slot" of its lowest group:
//   slot 0: joint rows (limits, motors, gear) + every contact of the island that contains the arm
//   slot 1, 2: contacts of a free body (block, drawer) that touches only static geometry (or the
//              other free body) — the common case; these records need no arm data at all.
// Within a slot the contacts keep Bullet's order.
#define Q_HDR 0          // ints {n_jrow | n_contact[0] << 8 | n_spin[0] << 16, t of slot 0's spin rows, t of its friction rows, t end of region 0}
                         //      {n_contact[1] | n_spin[1] << 8 | slot of free body 0 << 16 | slot of free body 1 << 18, start[1], t_spin[1], -}
                         //      {n_contact[2] | n_spin[2] << 8, start[2], t_spin[2], -}
#define Q_VSTAR 3        // 8 q: unconstrained velocities v* of the substep (word i = DoF i)
#define Q_DV 11          // 8 q: solver output M^-1 J^T lambda (word i = DoF i)
#define Q_ST 19          // start of the slot regions; region 0 starts here, region s at Q_ST + start[s];
                         // "t" offsets are relative to the start of the slot's region
// ---- region 0
#define T_BODY 0         // 2 q per free body: world inverse inertia {xx xy xz yy} {yz zz 1/m -}
#define T_MINV 4         // 12 rows x 3 q: arm inverse mass matrix (row d at T_MINV + 3 d, zero padded)
#define T_JROW 40        // n_jrow x 1 q: {packed, rhs, invD, hi}
                         //   packed: pd | pd2 << 8 | a << 16 | a2 << 20 | neg << 24 | sym << 25
                         //   pd/pd2: dv word of the DoF (arm: d, slide s: DVW_SLIDE(s); pd2 = 0xff: none);
                         //   a/a2: arm row (15: not arm); neg: J = -e_d; sym: lo = -hi (else lo = 0)
                         // then ceil(n_jrow / 4) q of accumulated impulses, then (4-q aligned) slot 0's contact rows
#define SB_MAXJROW 40
#define SB_MAXCONTACT 32
// slot 0 contacts.  The arm part of a row is EXPLICIT (12-wide J and B = M^-1 J^T); a free-body side is
// rebuilt from the contact frame like in slots 1, 2; a slide-body side is one scalar.  Items are multiples of
// 4 q, grouped by pass (all normal items in contact order, all spin items, all friction items); the first
// header of an item is its q 3:  H1 = {flags, rhs, invD, lambda},
//   flags: bits 0-1 item type (0 row, 1 normal row + geometry, 2 compact record, 3 pointer), bit 2 slide side,
//          bits 3-4 slide index, bit 5 free-body side, bit 6 free body index, bit 7 sign of the free side is -1
//   row (8 q):   {JA[0..3], JA[4..7], JA[8..11], H1} {BA[0..3], BA[4..7], BA[8..11], H2}
//                H2 = {cfm * invD (normal) | spin coefficient | mu (friction 1), t of the contact's normal item,
//                      J of the slide side, B of the slide side}
//   normal row + geometry (12 q): the row, then {n.xyz,-} {t1.xyz,-} {r.xyz,-} {-} of the free-body side
//                (spin and friction rows of the contact read it through H2.y)
//   friction item (16 q): the two friction rows of a contact
//   compact record (8 q; both sides free body / static, as in slots 1, 2): q 0..2 = record +0..+2, q 3 = {flags},
//                q 4..7 = record +3..+6
//   pointer (4 q; spin / friction item of a compact record): {t of the record, spin coefficient, rhs1, invD1} - - {flags}
// The solver kernel gives an env four lanes: lane c holds arm words 4c..4c+3 in registers and column c of the
// items; free-body and slide velocities are replicated in the registers of all four lanes.
#define XROW_Q 8
// ---- slots 1, 2: compact contact record (6 q, +1 q when the second side is the other free body):
//   +0 {packed, cfm * invD0, rhs0, invD0}     +1 {n.xyz, lambda0}        +2 {rP.xyz, mu}
//   +3 {t1.xyz, lambda1 (spin)}               +4 {rhs2, rhs3, invD2, invD3}   +5 {lambda2, lambda3, -, -}
//   +6 {rS.xyz, -} when side S is a free body
//   packed: kS << 2 | iP << 4 | iS << 7 | neg << 10 | spin << 11 | stride << 12
//   (kS: 0 static, 1 free; iX: free body index; neg: sign of P is -1; stride: q to the slot's next
//   record — a record never straddles the stage boundary PGS_STAGE_F)
// spin list entry (1 q): {t of the contact record, spin coefficient, rhs1, invD1}
#define CT_BASE_Q 6
#define SB_Q (Q_ST + 96 + SB_MAXCONTACT * (12 + 8 + 16) + 3 * 4 + SB_MAXCONTACT * 8 + 64 + 8)
#define SB_PAD_Q 48      // readable slack after the last group (the solvers prefetch one record ahead)
// stage capacities (q of shared memory per env) of the solver kernels
#ifndef PGS_STAGE_J
#define PGS_STAGE_J 68   // joint-row kernel: region 0 up to the impulses of <= 22 joint rows; 6 blocks per SM
#endif
#ifndef PGS_STAGE_F
#define PGS_STAGE_F 64   // free-body kernel: 10 contact records; 6 blocks per SM
#endif
#define PGS_G_LW 2       // arm-island kernel: an env owns 4 shared-memory columns (8 envs per warp)
#define PGS_NCLASS 4      // size classes of the arm-island kernel (by the q count of region 0): rows of 4 q per env
#ifndef PGS_ROWS_G0       // (overridable: the CPU tests build a variant with tiny stages to exercise the read-in-place path)
#define PGS_ROWS_G0 56   // class 0: 224 q per env, 7 blocks per SM (56 envs)
#define PGS_ROWS_G1 80   // class 1: 320 q per env, 5 blocks per SM (40 envs)
#define PGS_ROWS_G2 104  // class 2: 416 q per env, 4 blocks per SM (32 envs)
#define PGS_ROWS_G3 216  // class 3: 864 q per env (larger islands read the rest in place), 2 blocks per SM (16 envs)
#endif
#define PGS_ROWS_GMAX PGS_ROWS_G3
PRB_HD int pgs_class_rows(int cls) { return cls == 0 ? PGS_ROWS_G0 : (cls == 1 ? PGS_ROWS_G1 : (cls == 2 ? PGS_ROWS_G2 : PGS_ROWS_G3)); }
#define PGS_MAXJROW_J ((PGS_STAGE_J - T_JROW) * 4 / 5)     // njr + ceil(njr / 4) <= PGS_STAGE_J - T_JROW
#define DVW_SLIDE(s) (28 + (s))
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

// One side of one constraint row: J = unit force `dir` at world point pt (or unit torque when angular)
// on the body of collider col, times sign; B = M^-1 J^T.  Returns J.B and accumulates J.v*.  When gJ is
// given, an arm side writes (accum: adds to) its explicit J and B, 3 q each at stride 32; a slide side
// returns its scalar J and B in *js, *bs; free-body sides store nothing (the solvers rebuild them from the
// contact geometry).
template <int ND, class WM>
PRB_D float side_row(const DevModel& M, const WM& W, int col, v3 pt, v3 dir, float sign, bool angular,
                     float4* gJ, float4* gB, bool accum, float* js, float* bs, float* rel) {
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
    if (gJ) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float4 j4 = make_float4(J[4 * k], J[4 * k + 1], J[4 * k + 2], J[4 * k + 3]);
        float4 b4 = make_float4(B[4 * k], B[4 * k + 1], B[4 * k + 2], B[4 * k + 3]);
        if (accum) {                           // second arm side of an arm-arm contact
          const float4 pj = gJ[k * 32], pb = gB[k * 32];
          j4 = make_float4(j4.x + pj.x, j4.y + pj.y, j4.z + pj.z, j4.w + pj.w);
          b4 = make_float4(b4.x + pb.x, b4.y + pb.y, b4.z + pb.z, b4.w + pb.w);
        }
        gJ[k * 32] = j4; gB[k * 32] = b4;
      }
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
    *js = j; *bs = bb;
    d = j * bb; *rel += j * W.vs[o];
  }
  return d;
}

PRB_D int dvw_of(const DevModel& M, int d) {     // velocity DoF -> dv word of a joint row
  if (d < M.nd) return d;
  return DVW_SLIDE(d - M.nd - 6 * M.n_free);
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
  // ---- contacts: lane = contact
  const int nc = W.n_contact, njr = W.n_jrow;
  int colP = 0, colS = 0, kP = K_STATIC, kS = K_STATIC, grpP = 0, grpS = -1;
  bool swapped = false, has_spin = false;
  float spin = 0.f;
  Contact c;
  if (lane < nc) {
    c = W.ct[lane];
    const int ca = c.cols & 0xff, cb = (c.cols >> 8) & 0xff;
    const int kA = body_kind(M, col_dyn_body(M, ca)), kB = body_kind(M, col_dyn_body(M, cb));
    swapped = kind_rank(kB) > kind_rank(kA);
    colP = swapped ? cb : ca; colS = swapped ? ca : cb;
    kP = swapped ? kB : kA; kS = swapped ? kA : kB;
    grpP = kP == K_FREE ? M.col_body[colP] : 0;                 // 0: arm + slide bodies, 1 + b: free body b
    grpS = kS == K_STATIC ? -1 : (kS == K_FREE ? M.col_body[colS] : 0);
    spin = M.col_spin[ca] * M.col_fric[ca] + M.col_spin[cb] * M.col_fric[cb];
    has_spin = spin > 0.f;
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
  // ---- placement.  Slot 0: items grouped by pass
  const bool s0 = slot == 0;
  const bool cmp0 = s0 && kP == K_FREE;                              // both sides free body / static: compact record
  const bool freeS = s0 && !cmp0 && kS == K_FREE;                    // row items with a free-body side
  const int szN = s0 ? (freeS ? 12 : 8) : 0;
  const int szS = (s0 && has_spin) ? (cmp0 ? 4 : 8) : 0;
  const int szT = s0 ? (cmp0 ? 4 : 16) : 0;
  int totN, totS, totT;
  const int offN = warp_excl_scan(szN, lane, &totN);
  const int offS = warp_excl_scan(szS, lane, &totS);
  const int offT = warp_excl_scan(szT, lane, &totT);
  const int tN0 = (T_JROW + njr + ((njr + 3) >> 2) + 3) & ~3;
  const int tS0 = tN0 + totN, tT0 = tS0 + totS, tEnd0 = tT0 + totT;
  const int nc0 = __popc(__ballot_sync(FULL, s0)), ns0 = __popc(__ballot_sync(FULL, s0 && has_spin));
  // slots 1, 2: compact records + spin list per region
  const int size = (lane < nc && !s0) ? CT_BASE_Q + (kS == K_FREE ? 1 : 0) : 0;
  int t = 0, stride = size, t_spin_mine = 0, spin_rank = 0, region_mine = 0;
  int ncs[3] = {nc0, 0, 0}, nss[3] = {ns0, 0, 0}, tsp[3] = {tS0, 0, 0}, start[3] = {0, 0, 0};
  {
    int region = tEnd0;                               // start of the slot's region relative to Q_ST
#pragma unroll
    for (int sidx = 1; sidx < 3; sidx++) {
      const bool mine = slot == sidx;
      const unsigned mask = __ballot_sync(FULL, mine);
      const unsigned smask = __ballot_sync(FULL, mine && has_spin);
      int total;
      int ts = warp_excl_scan(mine ? size : 0, lane, &total);
      const bool straddle = mine && ts < PGS_STAGE_F && ts + size > PGS_STAGE_F;
      const unsigned sm = __ballot_sync(FULL, straddle);
      int shift = 0;
      if (sm) {
        const int sl_ = __ffs((int)sm) - 1;
        shift = PGS_STAGE_F - __shfl_sync(FULL, ts, sl_);
        if (lane >= sl_) ts += shift;
      }
      const unsigned later = mask & ~((2u << lane) - 1u);          // lanes of this slot after me (lane 31: none)
      const int nl = later ? __ffs((int)later) - 1 : lane;
      const int tnext = __shfl_sync(FULL, ts, nl);
      const int tend = total + shift;
      if (mine) {
        t = ts; stride = later ? tnext - ts : size;
        t_spin_mine = tend; spin_rank = __popc(smask & ((1u << lane) - 1u)); region_mine = region;
      }
      ncs[sidx] = __popc(mask); nss[sidx] = __popc(smask); tsp[sidx] = tend; start[sidx] = region;
      region += tend + nss[sidx];
    }
    if (lane == 0) { W.dbg_p = region; W.dbg_a = tEnd0; }
  }
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
    float rhs[4] = {0.f, 0.f, 0.f, 0.f}, invDs[4] = {0.f, 0.f, 0.f, 0.f};
    const int tN = tN0 + offN;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      v3 dir = k == 0 ? n : (k == 1 ? n : (k == 2 ? t1 : t2));
      const bool ang = (k == 1);
      if (k == 1 && !has_spin) continue;      // no torsional row: never visited by the solver
      const bool xrow = s0 && !cmp0;
      const int tr = k == 0 ? tN : (k == 1 ? tS0 + offS : tT0 + offT + (k == 3 ? 8 : 0));
      float4* row = &S.q(Q_ST + tr);
      float js = 0.f, bs = 0.f;
      if (xrow && kP != K_ARM) {               // no arm side: zero arm part
#pragma unroll
        for (int j = 0; j < 3; j++) { row[j * 32] = make_float4(0.f, 0.f, 0.f, 0.f); row[(4 + j) * 32] = make_float4(0.f, 0.f, 0.f, 0.f); }
      }
      float rel = 0.f, D = 0.f;
      D += side_row<ND>(M, W, colP, pP, dir, sP, ang, xrow ? row : nullptr, row + 4 * 32, false, &js, &bs, &rel);
      if (kS != K_STATIC) D += side_row<ND>(M, W, colS, pS, dir, -sP, ang, xrow ? row : nullptr, row + 4 * 32, true, &js, &bs, &rel);
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
      if (xrow) {                              // arm part written above; slide side as a scalar; free side through the geometry
        const int sld = kP == K_SLIDE ? M.col_body[colP] - 1 - M.n_free : (kS == K_SLIDE ? M.col_body[colS] - 1 - M.n_free : -1);
        if (kP == K_SLIDE && kS == K_SLIDE) W.overflow = 1;          // two slide bodies in one contact: not representable
        const int fb = freeS ? M.col_body[colS] - 1 : 0;
        const int flags = (k == 0 && freeS ? 1 : 0) | (sld >= 0 ? (4 | (sld << 3)) : 0) | (freeS ? (32 | (fb << 6) | (sP > 0.f ? 128 : 0)) : 0);
        row[3 * 32] = make_float4(__int_as_float(flags), rhs[k], invD, 0.f);
        row[7 * 32] = make_float4(k == 0 ? cfms : (k == 1 ? spin : (k == 2 ? mu : 0.f)), __int_as_float(tN), js, bs);
        if (k == 0 && freeS) {
          const v3 rS = pS - ld3(W.fpos[fb]);
          row[8 * 32] = make_float4(n.x, n.y, n.z, 0.f);
          row[9 * 32] = make_float4(t1.x, t1.y, t1.z, 0.f);
          row[10 * 32] = make_float4(rS.x, rS.y, rS.z, 0.f);
          row[11 * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    if (!s0 || cmp0) {                         // compact record (both sides are free bodies or static)
      const int iP = M.col_body[colP] - 1, iS = kS == K_FREE ? M.col_body[colS] - 1 : 0;
      const v3 rP = pP - ld3(W.fpos[iP]);
      const int packed = (kS << 2) | (iP << 4) | (iS << 7) | ((swapped ? 1 : 0) << 10) | ((has_spin ? 1 : 0) << 11) | (stride << 12);
      const float4 r0 = make_float4(__int_as_float(packed), cfms, rhs[0], invDs[0]), r1 = make_float4(n.x, n.y, n.z, 0.f);
      const float4 r2 = make_float4(rP.x, rP.y, rP.z, mu), r3 = make_float4(t1.x, t1.y, t1.z, 0.f);
      const float4 r4 = make_float4(rhs[2], rhs[3], invDs[2], invDs[3]), r5 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r6 = r5;
      if (kS == K_FREE) { const v3 rS = pS - ld3(W.fpos[iS]); r6 = make_float4(rS.x, rS.y, rS.z, 0.f); }
      if (!s0) {
        float4* rec = &S.q(Q_ST + region_mine + t);
        rec[0] = r0; rec[32] = r1; rec[64] = r2; rec[96] = r3; rec[128] = r4; rec[160] = r5;
        if (kS == K_FREE) rec[192] = r6;
        if (has_spin) S.q(Q_ST + region_mine + t_spin_mine + spin_rank) = make_float4(__int_as_float(t), spin, rhs[1], invDs[1]);
      } else {                                 // slot 0: compact item + pointer items in the spin / friction passes
        float4* it = &S.q(Q_ST + tN);
        it[0] = r0; it[32] = r1; it[64] = r2; it[96] = make_float4(__int_as_float(2), 0.f, 0.f, 0.f);
        it[128] = r3; it[160] = r4; it[192] = r5; it[224] = r6;
        const float4 ph = make_float4(__int_as_float(3), 0.f, 0.f, 0.f);
        if (has_spin) { float4* ps = &S.q(Q_ST + tS0 + offS); ps[0] = make_float4(__int_as_float(tN), spin, rhs[1], invDs[1]); ps[96] = ph; }
        float4* pt_ = &S.q(Q_ST + tT0 + offT);
        pt_[0] = make_float4(__int_as_float(tN), 0.f, 0.f, 0.f); pt_[96] = ph;
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    S.q(Q_HDR) = make_float4(__int_as_float(njr | (nc0 << 8) | (ns0 << 16)), __int_as_float(tS0), __int_as_float(tT0), __int_as_float(tEnd0));
    S.q(Q_HDR + 1) = make_float4(__int_as_float(ncs[1] | (nss[1] << 8) | (slotf[0] << 16) | (slotf[1] << 18)), __int_as_float(start[1]),
                                 __int_as_float(tsp[1]), 0.f);
    S.q(Q_HDR + 2) = make_float4(__int_as_float(ncs[2] | (nss[2] << 8)), __int_as_float(start[2]), __int_as_float(tsp[2]), 0.f);
    if (nc > W.dbg_c) W.dbg_c = nc;
    // island of the arm needs the arm-island solver: size class by the q count of region 0
    W.dbg_u = (nc0 > 0 || njr > PGS_MAXJROW_J) ? (tEnd0 <= 4 * PGS_ROWS_G0 ? 1 : (tEnd0 <= 4 * PGS_ROWS_G1 ? 2 : (tEnd0 <= 4 * PGS_ROWS_G2 ? 3 : 4))) : 0;
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
                                                                         int* __restrict__ heavy_list, int* __restrict__ heavy_cnt,
                                                                         const unsigned char* __restrict__ active) {
  typedef SetupMemT<SetupCfg> WM;
  PRB_SMEM_DECL2;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * SetupCfg::WPB + wib;
  if (e >= N) return;
  if (active != nullptr && active[e] == 0) return;       // masked stepping (reset: only the envs being reset settle)
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
    // a dropped contact marks the env; the mark is counted (once per env step) by the launch that observes
    if (lane == 0 && W.overflow) { if (O.ovf_env) O.ovf_env[e] = 1; else if (O.overflow) atomicAdd(O.overflow, 1ull); }
    if (lane == 0 && W.dbg_u) {                      // order within a list is immaterial: envs are independent
      const int cls = W.dbg_u - 1;                   // size class of region 0
      heavy_list[(size_t)cls * N + atomicAdd(heavy_cnt + 4 * cls, 1)] = e;    // heavy_cnt: {length, -, work counter, -} per class
    }
    if (lane == 0 && O.dbg) { O.dbg[4 * e] = W.dbg_u | (W.dbg_a << 8); O.dbg[4 * e + 1] = W.dbg_c; O.dbg[4 * e + 2] = W.dbg_p; O.dbg[4 * e + 3] = W.n_jrow; }
  }
  if (flags & SETUP_OBSERVE) {
    phase_observe(M, W, lane, O, (size_t)e, true);
    if (lane == 0 && O.ovf_env && O.ovf_env[e]) { O.ovf_env[e] = 0; if (O.overflow) atomicAdd(O.overflow, 1ull); }
  }
  __syncwarp();
  if (flags & (SETUP_INTEGRATE | SETUP_OBSERVE)) store_state(M, W, st, lane);
}

// ============================================================================ solver kernels
//   prb_pgs_joint_kernel   slot 0 of the envs whose arm island has no contacts: joint rows only (32 envs / warp)
//   prb_pgs_free_kernel    slots 1 and 2 (blockIdx.y): free-body islands, compact records (32 envs / warp)
//   prb_pgs_arm_kernel     slot 0 of the envs on a "heavy" list (arm island with contacts): joint rows +
//                          explicit rows; FOUR lanes per env (8 envs / warp), velocities in registers,
//                          one 2-step quad shuffle reduction per row.  Two size classes, two launches.
#define PGS_BLOCK 32
#define PGS_J_DVQ 4      // arm q 0..2, slides q 3
#define PGS_F_TAILQ 8    // free-body kernel: dv 4 q (body b at 2b, 2b+1) + body table 4 q
#define PGS_G_EPW (32 >> PGS_G_LW)                      // envs per warp of the arm-island kernel
#define PGS_SMEM_J ((PGS_STAGE_J + PGS_J_DVQ) * 32 * 16)
#define PGS_SMEM_F ((PGS_STAGE_F + PGS_F_TAILQ) * 32 * 16)
#define PGS_G_THREADS 32                                // threads per block of the arm-island kernel
#define PGS_SMEM_G(rows) ((rows) * PGS_G_THREADS * 16)

#ifdef PRB_EMU
static float4 g_emu_pgs_smem[(PGS_ROWS_GMAX + 2) * 32 + (PGS_STAGE_J + PGS_STAGE_F + 16) * 32];
#define PRB_PGS_SMEM_DECL float4* sm = g_emu_pgs_smem
#else
#define PRB_PGS_SMEM_DECL extern __shared__ __align__(16) float4 prb_pgs_smem[]; float4* sm = prb_pgs_smem
#endif

PRB_D float f4comp(const float4& a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? a.z : a.w)); }
PRB_D void f4add(float4& a, int k, float v) { a.x += k == 0 ? v : 0.f; a.y += k == 1 ? v : 0.f; a.z += k == 2 ? v : 0.f; a.w += k == 3 ? v : 0.f; }
PRB_D float dot4(const float4& a, const float4& b, float s) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s)))); }
PRB_D void axpy4(float4& y, const float4& b, float a) { y.x = fmaf(b.x, a, y.x); y.y = fmaf(b.y, a, y.y); y.z = fmaf(b.z, a, y.z); y.w = fmaf(b.w, a, y.w); }
PRB_D float4* stream_col(float* sbuf, int e) { return reinterpret_cast<float4*>(sbuf) + (size_t)(e >> 5) * (SB_Q * 32) + (e & 31); }

// ---- slot 0, arm island without contacts: joint rows only.  sl: this env's column (row t = q t of region 0)
template <int ND>
__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_joint_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N,
                                                                 const unsigned char* __restrict__ active) {
  PRB_PGS_SMEM_DECL;
  const int lane = threadIdx.x;
  const DevModel& M = *Mp;
  // persistent blocks: a block walks groups of 32 envs (launching one block per group costs more than
  // the solve: each block launch allocates its shared memory)
  for (int e = blockIdx.x * PGS_BLOCK + lane; e < N; e += gridDim.x * PGS_BLOCK) {
  if (active != nullptr && active[e] == 0) continue;
  float4* G = stream_col(sbuf, e);
  const int h0 = __float_as_int(G[Q_HDR * 32].x);
  const int njr = h0 & 0xff, nc0 = (h0 >> 8) & 0xff;
  if (nc0 > 0 || njr > PGS_MAXJROW_J) continue;          // on a heavy list: prb_pgs_arm_kernel solves it
  float4* sl = sm + lane;
  const float4* Gr = G + Q_ST * 32;
  float4* dvq = sl + PGS_STAGE_J * 32;                   // arm q 0..2, slides q 3
  {
    const int tq = T_JROW + njr;
#pragma unroll 8
    for (int q = T_MINV; q < tq; q++) sl[q * 32] = Gr[q * 32];
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < PGS_J_DVQ; i++) dvq[i * 32] = z4;
  for (int i = 0; i < ((njr + 3) >> 2); i++) sl[(T_JROW + njr + i) * 32] = z4;
  const float4* minv = sl + T_MINV * 32;
  float* jlam = reinterpret_cast<float*>(sl + (T_JROW + njr) * 32);  // word j at jlam[(j >> 2) * 128 + (j & 3)]
  float* dvw = reinterpret_cast<float*>(dvq);                        // word i at dvw[(i >> 2) * 128 + (i & 3)]
  const float ratio = M.params[P_GEAR_RATIO];
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    // A sweep that changes no impulse leaves the state untouched, so every later sweep repeats it
    // exactly: stopping there is bit-identical to running all the iterations.
    bool changed = false;
    // non-contact rows, sweep direction alternating per iteration
#pragma unroll 1
    for (int i = 0; i < njr; i++) {
      const int j = (it & 1) ? i : njr - 1 - i;
      const float4 r = sl[(T_JROW + j) * 32];
      const int pk = __float_as_int(r.x);
      int pd = pk & 0xff;
      const int pd2 = (pk >> 8) & 0xff, a = (pk >> 16) & 15, a2 = (pk >> 20) & 15;
      const int sidx = pd - DVW_SLIDE(0);                            // slide index when a == 15
      if (a == 15) pd = 12 + sidx;
      const float sg = ((pk >> 24) & 1) ? -1.0f : 1.0f;
      // the M^-1 rows are needed only if the impulse changes; issue their loads now anyway
      float4 m0[3], m1[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        m0[k] = z4; m1[k] = z4;
        if (a != 15) m0[k] = minv[(3 * a + k) * 32];
        if (a2 != 15) m1[k] = minv[(3 * a2 + k) * 32];
      }
      float* pu = &dvw[(pd >> 2) * 128 + (pd & 3)];
      float u = *pu;
      if (pd2 != 0xff) u = fmaf(ratio, dvw[(pd2 >> 2) * 128 + (pd2 & 3)], u);
      u *= sg;
      float* pl = &jlam[(j >> 2) * 128 + (j & 3)];
      const float l0 = *pl;
      const float hi = r.w, lo = ((pk >> 25) & 1) ? -hi : 0.f;
      const float nl = clampf(l0 + (r.y - u * r.z), lo, hi);
      const float dl = (nl - l0) * sg;
      if (dl != 0.f) {
        changed = true;
        *pl = nl;
        if (a != 15) {
          const float dl2 = dl * ratio;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            if (4 * k < ND) {
              float4 y = dvq[k * 32];
              axpy4(y, m0[k], dl);
              axpy4(y, m1[k], dl2);
              dvq[k * 32] = y;
            }
          }
        } else {
          *pu = fmaf(M.slide_minv[sidx], dl, *pu);
        }
      }
    }
    if (!changed) break;
  }
  float* gd = reinterpret_cast<float*>(G + Q_DV * 32);
  const int nd = M.nd, o = nd + 6 * M.n_free;
#pragma unroll 1
  for (int i = 0; i < nd; i++) gd[(i >> 2) * 128 + (i & 3)] = dvw[(i >> 2) * 128 + (i & 3)];
#pragma unroll 1
  for (int s = 0; s < M.n_slide; s++) gd[((o + s) >> 2) * 128 + ((o + s) & 3)] = dvw[3 * 128 + s];
  }
}

// ---- slots 1 and 2 (blockIdx.y + 1): islands of free bodies against static geometry / each other.
// dv of free body b at dvq q 2b, 2b+1 (6 words used)
struct FSide { int idx; float sgn; v3 r; };
struct FVel { v3 v, w, pv; };
PRB_D void fside_load(const FSide& s, const float4* dvq, FVel& V) {
  const float4 x = dvq[(2 * s.idx) * 32], y = dvq[(2 * s.idx + 1) * 32];
  V.v = V3(x.x, x.y, x.z); V.w = V3(x.w, y.x, y.y);
  V.pv = V.v + cross(V.w, s.r);
}
// dv += B P: P = sum of direction * impulse (linear rows) or the angular impulse (spin row)
PRB_D void fside_apply(const FSide& s, const FVel& V, float4* dvq, const float4* body, v3 P, bool ang) {
  const float4 i0 = body[(2 * s.idx) * 32], i1 = body[(2 * s.idx + 1) * 32];
  const float I[6] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y};
  const v3 Ps = P * s.sgn;
  v3 v = V.v, w = V.w;
  if (ang) w = w + symmul(I, Ps);
  else { v = v + Ps * i1.z; w = w + symmul(I, cross(s.r, Ps)); }
  dvq[(2 * s.idx) * 32] = make_float4(v.x, v.y, v.z, w.x);
  dvq[(2 * s.idx + 1) * 32] = make_float4(w.y, w.z, 0.f, 0.f);
}
PRB_D bool fsides_of(int pk, const float4* rec, const float4& q2, FSide& P, FSide& Sd) {
  P.idx = (pk >> 4) & 7; Sd.idx = (pk >> 7) & 7;
  P.sgn = ((pk >> 10) & 1) ? -1.0f : 1.0f; Sd.sgn = -P.sgn;
  P.r = V3(q2.x, q2.y, q2.z);
  Sd.r = V3(0, 0, 0);
  const bool two = ((pk >> 2) & 3) == K_FREE;
  if (two) { const float4 g = rec[CT_BASE_Q * 32]; Sd.r = V3(g.x, g.y, g.z); }
  return two;
}

__global__ void __launch_bounds__(PGS_BLOCK) prb_pgs_free_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf, int N,
                                                                const unsigned char* __restrict__ active) {
  PRB_PGS_SMEM_DECL;
  const int lane = threadIdx.x;
  const int slot = blockIdx.y + 1;
  const DevModel& M = *Mp;
  for (int e = blockIdx.x * PGS_BLOCK + lane; e < N; e += gridDim.x * PGS_BLOCK) {     // persistent blocks
  if (active != nullptr && active[e] == 0) continue;
  float4* G = stream_col(sbuf, e);
  const float4 h1 = G[(Q_HDR + 1) * 32], hs = G[(Q_HDR + slot) * 32];
  const int info = __float_as_int(h1.x);
  const int cnt = __float_as_int(hs.x);
  const int nc = cnt & 0xff, ns = (cnt >> 8) & 0xff, start = __float_as_int(hs.y), t_spin = __float_as_int(hs.z);
  bool owner = false;
  for (int b = 0; b < M.n_free; b++) owner = owner || ((info >> (16 + 2 * b)) & 3) == slot;
  if (!owner) continue;                                  // merged into another island
  float4* sl = sm + lane;
  float4* Gr = G + (Q_ST + start) * 32;
  float4* dvq = sl + PGS_STAGE_F * 32;
  float4* body = dvq + 4 * 32;
  {
    const int tq = min(t_spin + ns, PGS_STAGE_F);
#pragma unroll 8
    for (int q = 0; q < tq; q++) sl[q * 32] = Gr[q * 32];
  }
#define PGS_PTR(t_) ((t_) < PGS_STAGE_F ? sl + (t_) * 32 : Gr + (t_) * 32)
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; i++) { dvq[i * 32] = z4; body[i * 32] = G[(Q_ST + T_BODY + i) * 32]; }
  const int iters = M.solver_iters;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    bool changed = false;                                  // fixed point reached: later sweeps are exact repeats
    // ---- contact normals
    {
      int t = 0;
      float4* rec = PGS_PTR(t);
      float4 q0 = rec[0], q1 = rec[32], q2 = rec[64];
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const int pk = __float_as_int(q0.x);
        const int tn = t + ((pk >> 12) & 255);
        float4* recn = PGS_PTR(tn);
        const float4 n0 = recn[0], n1 = recn[32], n2 = recn[64];     // next record (readable slack after the last)
        FSide P, Sd;
        const bool two = fsides_of(pk, rec, q2, P, Sd);
        const v3 n = V3(q1.x, q1.y, q1.z);
        FVel VP, VS;
        fside_load(P, dvq, VP);
        float u = P.sgn * dot(n, VP.pv);
        if (two) { fside_load(Sd, dvq, VS); u += Sd.sgn * dot(n, VS.pv); }
        const float l0 = q1.w;
        const float nl = fmaxf(l0 + (q0.z - l0 * q0.y - u * q0.w), 0.f);
        const float dl = nl - l0;
        if (dl != 0.f) {
          changed = true;
          reinterpret_cast<float*>(rec + 32)[3] = nl;
          fside_apply(P, VP, dvq, body, n * dl, false);
          if (two) fside_apply(Sd, VS, dvq, body, n * dl, false);
        }
        t = tn; rec = recn; q0 = n0; q1 = n1; q2 = n2;
      }
    }
    // ---- spinning friction (Bullet skips the row while the normal impulse is 0)
#pragma unroll 1
    for (int i = 0; i < ns; i++) {
      const float4 h = *PGS_PTR(t_spin + i);
      float4* rec = PGS_PTR(__float_as_int(h.x));
      const float4 q0 = rec[0], q1 = rec[32], q2 = rec[64];
      const float tot = q1.w;
      if (!(tot > 0.f)) continue;
      FSide P, Sd;
      const bool two = fsides_of(__float_as_int(q0.x), rec, q2, P, Sd);
      const v3 n = V3(q1.x, q1.y, q1.z);
      FVel VP, VS;
      fside_load(P, dvq, VP);
      float u = P.sgn * dot(n, VP.w);
      if (two) { fside_load(Sd, dvq, VS); u += Sd.sgn * dot(n, VS.w); }
      const float lim = h.y * tot;
      float* pl = reinterpret_cast<float*>(rec + 96) + 3;
      const float l0 = *pl;
      const float nl = clampf(l0 + (h.z - u * h.w), -lim, lim);
      const float dl = nl - l0;
      if (dl != 0.f) {
        changed = true;
        *pl = nl;
        fside_apply(P, VP, dvq, body, n * dl, true);
        if (two) fside_apply(Sd, VS, dvq, body, n * dl, true);
      }
    }
    // ---- lateral friction: the two rows of a contact are solved together (implicit cone)
    {
      int t = 0;
      float4* rec = PGS_PTR(t);
      float4 q0 = rec[0], q1 = rec[32], q2 = rec[64], q3 = rec[96], q4 = rec[128], q5 = rec[160];
#pragma unroll 1
      for (int c = 0; c < nc; c++) {
        const int pk = __float_as_int(q0.x);
        const int tn = t + ((pk >> 12) & 255);
        float4* recn = PGS_PTR(tn);
        const float4 n0 = recn[0], n1 = recn[32], n2 = recn[64], n3 = recn[96], n4 = recn[128], n5 = recn[160];
        FSide P, Sd;
        const bool two = fsides_of(pk, rec, q2, P, Sd);
        const v3 n = V3(q1.x, q1.y, q1.z), t1 = V3(q3.x, q3.y, q3.z), t2 = cross(n, t1);
        FVel VP, VS;
        fside_load(P, dvq, VP);
        float ua = P.sgn * dot(t1, VP.pv), ub = P.sgn * dot(t2, VP.pv);
        if (two) {
          fside_load(Sd, dvq, VS);
          ua += Sd.sgn * dot(t1, VS.pv); ub += Sd.sgn * dot(t2, VS.pv);
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
          changed = true;
          rec[160] = make_float4(na, nb, 0.f, 0.f);
          const v3 Pv = t1 * d1 + t2 * d2;
          fside_apply(P, VP, dvq, body, Pv, false);
          if (two) fside_apply(Sd, VS, dvq, body, Pv, false);
        }
        t = tn; rec = recn; q0 = n0; q1 = n1; q2 = n2; q3 = n3; q4 = n4; q5 = n5;
      }
    }
    if (!changed) break;
  }
#undef PGS_PTR
  float* gd = reinterpret_cast<float*>(G + Q_DV * 32);
  for (int b = 0; b < M.n_free; b++)
    if (((info >> (16 + 2 * b)) & 3) == slot) {
      const float4 x = dvq[(2 * b) * 32], y = dvq[(2 * b + 1) * 32];
      const float v[6] = {x.x, x.y, x.z, x.w, y.x, y.y};
      const int o = M.nd + 6 * b;
#pragma unroll
      for (int k = 0; k < 6; k++) gd[((o + k) >> 2) * 128 + ((o + k) & 3)] = v[k];
    }
  }
}

// ---- slot 0 of the envs whose arm island has contacts (a heavy list): joint rows + explicit rows.
// Four lanes per env (quad): lane c of the quad keeps words 4c..4c+3 of the island's velocity change
// (A: arm, F: free bodies + slides) in registers and owns column c of the env's four shared-memory
// columns: q number t of region 0 sits in row t >> 2, column t & 3, so a row's J (B) slices are one
// conflict-free LDS.128 per lane and its header a quad broadcast.  J . dv = 4 FMAs per lane + a 2-step
// quad shuffle reduction.
struct QuadMem {
  float4* base;            // column 0 of this env in shared memory
  float4* Gr;              // this env's column of region 0 in the stream
  int cap;                 // staged q count; items that do not end below it are read from the stream in place
  static constexpr int rs = 32;   // row stride = columns of the block's stage = threads per block
  PRB_D float4* sp(int t) const { return base + (t >> 2) * rs + (t & 3); }     // always-staged q (fixed part of region 0)
  PRB_D bool staged(int t) const { return t + 16 <= cap; }                    // item starting at t (items are <= 16 q)
  PRB_D float4 ldq(int t) const { return t < cap ? *sp(t) : Gr[t * 32]; }     // immutable data (geometry)
};
// an item in shared memory (q k of the item at rp[(k >> 2) * 32 + (k & 3)]) or in the stream (gp[k * 32])
struct RowS {
  float4* rp;
  static constexpr int rs = 32;
  template <int K0> PRB_D float4 ld(int c) const { return rp[(K0 >> 2) * rs + c]; }
  template <int K> PRB_D float4* q() const { return rp + (K >> 2) * rs + (K & 3); }
};
struct RowG {
  float4* gp;
  template <int K0> PRB_D float4 ld(int c) const { return gp[(K0 + c) * 32]; }
  template <int K> PRB_D float4* q() const { return gp + K * 32; }
};
PRB_D float quad_sum(unsigned qmask, float v) {
  v += __shfl_xor_sync(qmask, v, 1);
  v += __shfl_xor_sync(qmask, v, 2);
  return v;
}
// velocities of the non-arm bodies of the island, replicated in the four lanes of the quad
struct QuadFree {
  v3 v[PRB_MAXFREE], w[PRB_MAXFREE];
  float4 sl;                                     // slide bodies
  float I[PRB_MAXFREE][6], invm[PRB_MAXFREE];    // world inverse inertia, 1 / mass
  PRB_D v3 vel(int b) const { return b ? v[PRB_MAXFREE - 1] : v[0]; }
  PRB_D v3 ang(int b) const { return b ? w[PRB_MAXFREE - 1] : w[0]; }
  // J . dv of a free-body side: unit force d at lever arm r (or unit torque d), times sgn
  PRB_D float jdot(int b, float sgn, v3 r, v3 d, bool angular) const {
    const v3 ww = ang(b);
    return sgn * (angular ? dot(d, ww) : dot(d, vel(b) + cross(ww, r)));
  }
  // dv += B P: P = sum of direction * impulse (linear rows) or the angular impulse (spin row)
  PRB_D void apply(int b, float sgn, v3 r, v3 P, bool angular) {
    const v3 Ps = P * sgn;
    // register selects (no runtime-indexed arrays: they would live in local memory)
    float Ib[6];
#pragma unroll
    for (int k = 0; k < 6; k++) Ib[k] = b ? I[PRB_MAXFREE - 1][k] : I[0][k];
    const float im = b ? invm[PRB_MAXFREE - 1] : invm[0];
    v3 dv = V3(0, 0, 0), dw;
    if (angular) dw = symmul(Ib, Ps);
    else { dv = Ps * im; dw = symmul(Ib, cross(r, Ps)); }
    if (b) { v[PRB_MAXFREE - 1] = v[PRB_MAXFREE - 1] + dv; w[PRB_MAXFREE - 1] = w[PRB_MAXFREE - 1] + dw; }
    else { v[0] = v[0] + dv; w[0] = w[0] + dw; }
  }
};
PRB_D float* quad_lam0(const QuadMem& m, int tN) {        // normal impulse of the contact whose normal item is at tN
  float4* h = m.staged(tN) ? m.sp(tN + 3) : m.Gr + (tN + 3) * 32;
  const int type = __float_as_int(h->x) & 3;
  if (type == 2) { float4* q1 = m.staged(tN) ? m.sp(tN + 1) : m.Gr + (tN + 1) * 32; return reinterpret_cast<float*>(q1) + 3; }
  return reinterpret_cast<float*>(h) + 3;
}
// one explicit row; KIND 0: contact normal (lambda >= 0, soft CFM), 1: spin (|lambda| <= coefficient * normal impulse)
template <int KIND, class Row>
PRB_D bool quad_xrow(const Row& r, const QuadMem& m, int c, unsigned qmask, float4& A, QuadFree& Fr, float tot) {
  const float4 H1 = *r.template q<3>(), H2 = *r.template q<7>();
  const int flags = __float_as_int(H1.x);
  const float4 J = r.template ld<0>(c), B = r.template ld<4>(c);
  float u = quad_sum(qmask, c < 3 ? dot4(J, A, 0.f) : 0.f);
  const int sidx = (flags >> 3) & 3, fb = (flags >> 6) & 1;
  const float fs = (flags & 128) ? -1.0f : 1.0f;
  v3 n = V3(0, 0, 0), rr = n;
  if (flags & 4) u = fmaf(H2.z, f4comp(Fr.sl, sidx), u);
  if (flags & 32) {
    const int tg = (KIND == 0 ? 0 : __float_as_int(H2.y)) + 8;      // geometry of the contact: after its normal row
    float4 g0, g2;
    if (KIND == 0) { g0 = *r.template q<8>(); g2 = *r.template q<10>(); } else { g0 = m.ldq(tg); g2 = m.ldq(tg + 2); }
    n = V3(g0.x, g0.y, g0.z); rr = V3(g2.x, g2.y, g2.z);
    u += Fr.jdot(fb, fs, rr, n, KIND == 1);
  }
  const float l0 = H1.w;
  float nl;
  if (KIND == 0) nl = fmaxf(l0 + (H1.y - l0 * H2.x - u * H1.z), 0.f);
  else { const float lim = H2.x * tot; nl = clampf(l0 + (H1.y - u * H1.z), -lim, lim); }
  const float dl = nl - l0;
  __syncwarp(qmask);                                     // every lane of the quad has read lambda
  if (dl != 0.f) {
    if (c == 3) reinterpret_cast<float*>(r.template q<3>())[3] = nl;
    if (c < 3) axpy4(A, B, dl);
    if (flags & 4) f4add(Fr.sl, sidx, H2.w * dl);
    if (flags & 32) Fr.apply(fb, fs, rr, n * dl, KIND == 1);
  }
  __syncwarp(qmask);                                     // the new impulse is visible to the quad
  return dl != 0.f;
}
// lateral friction of an explicit contact: the two rows are solved together (implicit cone)
template <class Row>
PRB_D bool quad_xfriction(const Row& r, const QuadMem& m, int c, unsigned qmask, float4& A, QuadFree& Fr) {
  const float4 H1 = *r.template q<3>(), H2 = *r.template q<7>(), G1 = *r.template q<11>(), G2 = *r.template q<15>();
  const int flags = __float_as_int(H1.x);
  const float4 J1 = r.template ld<0>(c), B1 = r.template ld<4>(c), J2 = r.template ld<8>(c), B2 = r.template ld<12>(c);
  const int tN = __float_as_int(H2.y);
  const float tot = *quad_lam0(m, tN);
  float ua = quad_sum(qmask, c < 3 ? dot4(J1, A, 0.f) : 0.f), ub = quad_sum(qmask, c < 3 ? dot4(J2, A, 0.f) : 0.f);
  const int sidx = (flags >> 3) & 3, fb = (flags >> 6) & 1;
  const float fs = (flags & 128) ? -1.0f : 1.0f;
  v3 t1 = V3(0, 0, 0), t2 = t1, rr = t1;
  if (flags & 4) { const float vs = f4comp(Fr.sl, sidx); ua = fmaf(H2.z, vs, ua); ub = fmaf(G2.z, vs, ub); }
  if (flags & 32) {
    const float4 g0 = m.ldq(tN + 8), g1 = m.ldq(tN + 9), g2 = m.ldq(tN + 10);
    const v3 n = V3(g0.x, g0.y, g0.z);
    t1 = V3(g1.x, g1.y, g1.z); t2 = cross(n, t1); rr = V3(g2.x, g2.y, g2.z);
    ua += Fr.jdot(fb, fs, rr, t1, false); ub += Fr.jdot(fb, fs, rr, t2, false);
  }
  const float lim = H2.x * tot;
  const float la = H1.w, lb = G1.w;
  const float sumA = la + (H1.y - ua * H1.z);
  const float sumB = lb + (G1.y - ub * G1.z);
  float na = sumA, nb = sumB;
  if (sumA < -lim || sumA > lim || sumB < -lim || sumB > lim) {
    const float ss = sumA * sumA + sumB * sumB;
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    const float ca_ = fabsf(lim * sumA) * inv, cb_ = ss > 0.f ? fabsf(lim * sumB) * inv : fabsf(lim);
    na = clampf(sumA, -ca_, ca_); nb = clampf(sumB, -cb_, cb_);
  }
  const float d1 = na - la, d2 = nb - lb;
  __syncwarp(qmask);
  if (d1 != 0.f || d2 != 0.f) {
    if (c == 3) { reinterpret_cast<float*>(r.template q<3>())[3] = na; reinterpret_cast<float*>(r.template q<11>())[3] = nb; }
    if (c < 3) { axpy4(A, B1, d1); axpy4(A, B2, d2); }
    if (flags & 4) f4add(Fr.sl, sidx, H2.w * d1 + G2.w * d2);
    if (flags & 32) Fr.apply(fb, fs, rr, t1 * d1 + t2 * d2, false);
  }
  __syncwarp(qmask);
  return d1 != 0.f || d2 != 0.f;
}
// a compact record (both sides free body / static) inside the arm island: the free-body solver's arithmetic
// on the replicated velocities.  PASS 0 normal, 1 spin (h: its pointer item), 2 friction
template <int PASS, class Row>
PRB_D bool quad_compact(const Row& r, float4 h, int c, unsigned qmask, QuadFree& Fr) {
  bool changed = false;
  const float4 q0 = *r.template q<0>(), q1 = *r.template q<1>(), q2 = *r.template q<2>();
  const int pk = __float_as_int(q0.x);
  const int iP = (pk >> 4) & 7, iS = (pk >> 7) & 7;
  const float sP = ((pk >> 10) & 1) ? -1.0f : 1.0f;
  const bool two = ((pk >> 2) & 3) == K_FREE;
  const v3 n = V3(q1.x, q1.y, q1.z), rP = V3(q2.x, q2.y, q2.z);
  v3 rS = V3(0, 0, 0);
  if (two) { const float4 g = *r.template q<7>(); rS = V3(g.x, g.y, g.z); }
  if (PASS == 0) {
    float u = Fr.jdot(iP, sP, rP, n, false);
    if (two) u += Fr.jdot(iS, -sP, rS, n, false);
    const float l0 = q1.w;
    const float nl = fmaxf(l0 + (q0.z - l0 * q0.y - u * q0.w), 0.f);
    const float dl = nl - l0;
    __syncwarp(qmask);
    changed = dl != 0.f;
    if (dl != 0.f) {
      if (c == 3) reinterpret_cast<float*>(r.template q<1>())[3] = nl;
      Fr.apply(iP, sP, rP, n * dl, false);
      if (two) Fr.apply(iS, -sP, rS, n * dl, false);
    }
  } else if (PASS == 1) {
    const float tot = q1.w;
    if (tot > 0.f) {                                   // Bullet skips the row while the normal impulse is 0
      float u = Fr.jdot(iP, sP, rP, n, true);
      if (two) u += Fr.jdot(iS, -sP, rS, n, true);
      const float lim = h.y * tot;
      float* pl = reinterpret_cast<float*>(r.template q<4>()) + 3;
      const float l0 = *pl;
      const float nl = clampf(l0 + (h.z - u * h.w), -lim, lim);
      const float dl = nl - l0;
      __syncwarp(qmask);
      changed = dl != 0.f;
      if (dl != 0.f) {
        if (c == 3) *pl = nl;
        Fr.apply(iP, sP, rP, n * dl, true);
        if (two) Fr.apply(iS, -sP, rS, n * dl, true);
      }
    }
  } else {
    const float4 q3 = *r.template q<4>(), q4 = *r.template q<5>(), q5 = *r.template q<6>();
    const v3 t1 = V3(q3.x, q3.y, q3.z), t2 = cross(n, t1);
    float ua = Fr.jdot(iP, sP, rP, t1, false), ub = Fr.jdot(iP, sP, rP, t2, false);
    if (two) { ua += Fr.jdot(iS, -sP, rS, t1, false); ub += Fr.jdot(iS, -sP, rS, t2, false); }
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
    __syncwarp(qmask);
    changed = d1 != 0.f || d2 != 0.f;
    if (d1 != 0.f || d2 != 0.f) {
      if (c == 3) *r.template q<6>() = make_float4(na, nb, 0.f, 0.f);
      const v3 Pv = t1 * d1 + t2 * d2;
      Fr.apply(iP, sP, rP, Pv, false);
      if (two) Fr.apply(iS, -sP, rS, Pv, false);
    }
  }
  __syncwarp(qmask);
  return changed;
}

template <int ND>
__global__ void __launch_bounds__(32) prb_pgs_arm_kernel(const DevModel* __restrict__ Mp, float* __restrict__ sbuf,
                                                                const int* __restrict__ heavy_list, int* __restrict__ heavy_cnt, int rows) {
  PRB_PGS_SMEM_DECL;
  // 32 threads = 8 envs per block.  (Smaller blocks were measured: 16 threads equal, 8 and 4 slower — the
  // kernel is bound by the instruction count per row visit, which a warp amortises over its converged quads.)
  const int lane = threadIdx.x, c = lane & 3, qb = lane & ~3;
  const unsigned qmask = 0xfu << qb;
  const int cnt = *heavy_cnt;
  const DevModel& M = *Mp;
  // persistent blocks, dynamic scheduling: every quad (the quads of a warp are independent) draws the
  // next env of the list from a device counter (heavy_cnt[2]), so long islands do not queue behind each other
  for (;;) {
  int i = 0;
  if (c == 0) i = atomicAdd(heavy_cnt + 2, 1);
  i = __shfl_sync(qmask, i, qb);
  if (i >= cnt) break;
  const int e = heavy_list[i];
  float4* G = stream_col(sbuf, e);
  const float4 hdr = G[Q_HDR * 32];
  const int h0 = __float_as_int(hdr.x), info = __float_as_int(G[(Q_HDR + 1) * 32].x);
  const int njr = h0 & 0xff, nc = (h0 >> 8) & 0xff, ns = (h0 >> 16) & 0xff;
  const int tS0 = __float_as_int(hdr.y), tT0 = __float_as_int(hdr.z), tEnd = __float_as_int(hdr.w);
  const int tN0 = (T_JROW + njr + ((njr + 3) >> 2) + 3) & ~3;
  QuadMem m;
  m.base = sm + qb; m.Gr = G + Q_ST * 32; m.cap = rows << 2;
  {
    const int tq = min(tEnd, m.cap);
#pragma unroll 4
    for (int q = c; q < tq; q += 4) *m.sp(q) = m.Gr[q * 32];
  }
  __syncwarp(qmask);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c == 0) for (int k = 0; k < ((njr + 3) >> 2); k++) *m.sp(T_JROW + njr + k) = z4;
  __syncwarp(qmask);
  float4 A = z4;
  QuadFree Fr;
  Fr.sl = z4;
#pragma unroll
  for (int b = 0; b < PRB_MAXFREE; b++) {
    Fr.v[b] = V3(0, 0, 0); Fr.w[b] = V3(0, 0, 0);
    const float4 i0 = *m.sp(T_BODY + 2 * b), i1 = *m.sp(T_BODY + 2 * b + 1);
    Fr.I[b][0] = i0.x; Fr.I[b][1] = i0.y; Fr.I[b][2] = i0.z; Fr.I[b][3] = i0.w; Fr.I[b][4] = i1.x; Fr.I[b][5] = i1.y;
    Fr.invm[b] = i1.z;
  }
  const float ratio = M.params[P_GEAR_RATIO];
  const int iters = M.solver_iters;
  const int t_jlam = T_JROW + njr;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    bool changed = false;                                  // fixed point reached: later sweeps are exact repeats
    // ---- non-contact rows, sweep direction alternating per iteration
#pragma unroll 1
    for (int ii = 0; ii < njr; ii++) {
      const int j = (it & 1) ? ii : njr - 1 - ii;
      const float4 r = *m.sp(T_JROW + j);
      const int pk = __float_as_int(r.x);
      const int pd = pk & 0xff, a = (pk >> 16) & 15, a2 = (pk >> 20) & 15;
      const int sidx = pd - DVW_SLIDE(0);                            // slide index when a == 15
      const float sg = ((pk >> 24) & 1) ? -1.0f : 1.0f;
      float4 m0 = z4, m1 = z4;                                       // this lane's slice of the M^-1 rows
      if (a != 15 && c < 3) m0 = *m.sp(T_MINV + 3 * a + c);
      if (a2 != 15 && c < 3) m1 = *m.sp(T_MINV + 3 * a2 + c);
      // u = J . dv: J = e_a (+ ratio e_a2): broadcast of the owning lane's word; slide rows are replicated
      float u = __shfl_sync(qmask, f4comp(A, a & 3), qb + ((a >> 2) & 3));
      if (a2 != 15) u = fmaf(ratio, __shfl_sync(qmask, f4comp(A, a2 & 3), qb + (a2 >> 2)), u);
      if (a == 15) u = f4comp(Fr.sl, sidx);
      u *= sg;
      float* pl = reinterpret_cast<float*>(m.sp(t_jlam + (j >> 2))) + (j & 3);
      const float l0 = *pl;
      const float hi = r.w, lo = ((pk >> 25) & 1) ? -hi : 0.f;
      const float nl = clampf(l0 + (r.y - u * r.z), lo, hi);
      const float dl = (nl - l0) * sg;
      __syncwarp(qmask);                                             // every lane of the quad has read l0
      if (dl != 0.f) {
        changed = true;
        if (c == 0) *pl = nl;
        if (a != 15) { axpy4(A, m0, dl); axpy4(A, m1, dl * ratio); }
        else f4add(Fr.sl, sidx, M.slide_minv[sidx] * dl);
      }
      __syncwarp(qmask);                                             // the new impulse is visible to the quad
    }
    // ---- contact normals
    {
      int t = tN0;
#pragma unroll 1
      for (int k = 0; k < nc; k++) {
        const bool st = m.staged(t);
        const int type = __float_as_int((st ? m.sp(t + 3) : m.Gr + (t + 3) * 32)->x) & 3;
        if (type == 2) changed |= st ? quad_compact<0>(RowS{m.sp(t)}, z4, c, qmask, Fr) : quad_compact<0>(RowG{m.Gr + t * 32}, z4, c, qmask, Fr);
        else changed |= st ? quad_xrow<0>(RowS{m.sp(t)}, m, c, qmask, A, Fr, 0.f) : quad_xrow<0>(RowG{m.Gr + t * 32}, m, c, qmask, A, Fr, 0.f);
        t += type == 1 ? 12 : 8;
      }
    }
    // ---- spinning friction (Bullet skips the row while the normal impulse is 0)
    {
      int t = tS0;
#pragma unroll 1
      for (int k = 0; k < ns; k++) {
        const bool st = m.staged(t);
        const int type = __float_as_int((st ? m.sp(t + 3) : m.Gr + (t + 3) * 32)->x) & 3;
        if (type == 3) {
          const float4 h = *(st ? m.sp(t) : m.Gr + t * 32);
          const int tr = __float_as_int(h.x);
          changed |= m.staged(tr) ? quad_compact<1>(RowS{m.sp(tr)}, h, c, qmask, Fr) : quad_compact<1>(RowG{m.Gr + tr * 32}, h, c, qmask, Fr);
          t += 4;
        } else {
          const float4 H2 = *(st ? m.sp(t + 7) : m.Gr + (t + 7) * 32);
          const float tot = *quad_lam0(m, __float_as_int(H2.y));     // normal impulse of the contact
          if (tot > 0.f) {
            changed |= st ? quad_xrow<1>(RowS{m.sp(t)}, m, c, qmask, A, Fr, tot) : quad_xrow<1>(RowG{m.Gr + t * 32}, m, c, qmask, A, Fr, tot);
          }
          t += 8;
        }
      }
    }
    // ---- lateral friction
    {
      int t = tT0;
#pragma unroll 1
      for (int k = 0; k < nc; k++) {
        const bool st = m.staged(t);
        const int type = __float_as_int((st ? m.sp(t + 3) : m.Gr + (t + 3) * 32)->x) & 3;
        if (type == 3) {
          const int tr = __float_as_int((st ? m.sp(t) : m.Gr + t * 32)->x);
          changed |= m.staged(tr) ? quad_compact<2>(RowS{m.sp(tr)}, z4, c, qmask, Fr) : quad_compact<2>(RowG{m.Gr + tr * 32}, z4, c, qmask, Fr);
          t += 4;
        } else {
          changed |= st ? quad_xfriction(RowS{m.sp(t)}, m, c, qmask, A, Fr) : quad_xfriction(RowG{m.Gr + t * 32}, m, c, qmask, A, Fr);
          t += 16;
        }
      }
    }
    if (!changed) break;
  }
  // ---- velocity change -> stream (linear DoF order): arm, slides, and the free bodies this island owns
  float* gd = reinterpret_cast<float*>(G + Q_DV * 32);
  const int nd = M.nd, nf = M.n_free;
  const float a4[4] = {A.x, A.y, A.z, A.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int wa = 4 * c + k;                    // arm DoF of this lane's word k
    if (c < 3 && wa < nd) gd[(wa >> 2) * 128 + (wa & 3)] = a4[k];
  }
  if (c == 3) {
    for (int s_ = 0; s_ < M.n_slide; s_++) { const int d = nd + 6 * nf + s_; gd[(d >> 2) * 128 + (d & 3)] = f4comp(Fr.sl, s_); }
#pragma unroll
    for (int b = 0; b < PRB_MAXFREE; b++)
      if (b < nf && ((info >> (16 + 2 * b)) & 3) == 0) {
        const float v6[6] = {Fr.v[b].x, Fr.v[b].y, Fr.v[b].z, Fr.w[b].x, Fr.w[b].y, Fr.w[b].z};
#pragma unroll
        for (int k = 0; k < 6; k++) { const int d = nd + 6 * b + k; gd[(d >> 2) * 128 + (d & 3)] = v6[k]; }
      }
  }
  __syncwarp(qmask);                                     // the quad's columns are re-staged by the next env
  }
}
