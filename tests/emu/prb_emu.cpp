// prb_emu.cpp — runs the product's kernel source under the CPU SIMT emulator (tests only).
#include "cuda_emu.h"
#include "../../roboticsplayroompybullet_b200/csrc/prb_reset.cuh"
#include "../../roboticsplayroompybullet_b200/csrc/prb_convert.h"

static DevModel g_M;
static std::string g_err;

static std::vector<float> g_sbuf;
static std::vector<float4> g_hbuf;
static long long g_class_count[ARM_NCLASS] = {0};     // envs solved per arm-island class since the last query
static int g_fused = 0;     // 1: run the fused warp-per-env kernel instead of the split pipeline
template <int ND>
static void run_step_nd(float* state, DevOut O, int N, int nsub, int observe, const unsigned char* active = nullptr,
                        const int* elist = nullptr, int n_list = 0) {
  if (g_fused && active == nullptr && elist == nullptr) {
    emu_dim3 g, b; b.x = 32 * CfgL::WPB; g.x = (N + CfgL::WPB - 1) / CfgL::WPB;
    emu::launch(g, b, [&]() { prb_step_kernel<ND, CfgL>(&g_M, state, O, nullptr, nullptr, nullptr, nullptr, N, nsub, observe); });
    return;
  }
  g_sbuf.assign(sbuf_bytes(N) / sizeof(float), 0.f);
  g_hbuf.assign(hbuf_bytes(N) / sizeof(float4), make_float4(0.f, 0.f, 0.f, 0.f));
  int heavy_cnt[4 * ARM_NCLASS] = {0};
  emu_dim3 gs, bs, gp, bp;
  const int n = elist != nullptr ? n_list : N;         // work items (prb_capi.cu launch_step)
  bs.x = 32 * SetupCfg::WPB; gs.x = (n + SetupCfg::WPB - 1) / SetupCfg::WPB;
  bp.x = PGS_BLOCK; gp.x = (n + PGS_BLOCK - 1) / PGS_BLOCK;
  for (int i = 0; i <= nsub; i++) {
    int flags = (i > 0 ? SETUP_INTEGRATE : 0) | (i < nsub ? SETUP_BUILD : 0) | ((i == nsub && observe) ? SETUP_OBSERVE : 0);
    if (flags == 0) break;
    memset(heavy_cnt, 0, sizeof(heavy_cnt));
    emu::launch(gs, bs, [&]() { prb_setup_kernel<ND>(&g_M, state, g_sbuf.data(), O, N, flags, g_hbuf.data(), heavy_cnt, active, elist, n_list); });
    if (i < nsub) {
      emu::launch(gp, bp, [&]() { prb_pgs_joint_kernel<ND>(&g_M, g_sbuf.data(), N, active, elist, n_list); });
      for (int y = 0; y < g_M.n_free; y++) {
        emu_dim3 gy = gp;
        emu::launch_y(gy, bp, y, [&]() { prb_pgs_free_kernel(&g_M, g_sbuf.data(), N, active, elist, n_list); });
      }
      for (int k = 0; k < ARM_NCLASS; k++) {
        g_class_count[k] += heavy_cnt[4 * k];
        emu_dim3 gh, bq; gh.x = 2; bq.x = 32;             // two persistent blocks share the class's work counter
        float4* hc = g_hbuf.data() + arm_class_base(k, N);
        int* wc = heavy_cnt + 4 * k;
        const int cq = arm_capq(k), bq_ = arm_bufq(k);
        switch (k) {
          case 0: emu::launch(gh, bq, [&]() { prb_pgs_arm_kernel<ND, false, arm_lanes(0)>(&g_M, g_sbuf.data(), hc, wc, cq, bq_); }); break;
          case 1: emu::launch(gh, bq, [&]() { prb_pgs_arm_kernel<ND, false, arm_lanes(1)>(&g_M, g_sbuf.data(), hc, wc, cq, bq_); }); break;
          case 2: emu::launch(gh, bq, [&]() { prb_pgs_arm_kernel<ND, false, arm_lanes(2)>(&g_M, g_sbuf.data(), hc, wc, cq, bq_); }); break;
          case 3: emu::launch(gh, bq, [&]() { prb_pgs_arm_kernel<ND, false, arm_lanes(3)>(&g_M, g_sbuf.data(), hc, wc, cq, bq_); }); break;
          default: emu::launch(gh, bq, [&]() { prb_pgs_arm_kernel<ND, true, arm_lanes(4)>(&g_M, g_sbuf.data(), hc, wc, cq, bq_); }); break;
        }
      }
    }
  }
}
static void run_step(float* state, DevOut O, int N, int nsub, int observe) {
  if (g_M.nd == 12) run_step_nd<12>(state, O, N, nsub, observe); else run_step_nd<9>(state, O, N, nsub, observe);
}

static unsigned long long g_overflow = 0;
static int g_reset_rounds = 0;
// mirrors reset_rounds() of prb_capi.cu
template <int ND>
static void emu_reset_nd(float* state, DevOut o, const unsigned char* mask, int N, unsigned long long seed, unsigned env_offset) {
  std::vector<unsigned char> pending(N, 0), ovf(N, 0);
  std::vector<int> ctl(2 * N, 0), elist[2];
  elist[0].assign(N, 0); elist[1].assign(N, 0);
  int n_prev = 0;
  o.ovf_env = ovf.data();
  emu_dim3 gp, bp, gf, bf;
  bp.x = 128; gp.x = (N + 127) / 128;
  bf.x = 32 * SetupCfg::WPB; gf.x = (N + SetupCfg::WPB - 1) / SetupCfg::WPB;
  g_reset_rounds = 0;
  for (int round = 0; round < RESET_MAX_ATTEMPTS * RESET_MAX_TRIES; round++) {
    const bool list0 = round == 0 && mask != nullptr;
    int n0 = 0;
    emu::launch(gp, bp, [&]() { prb_reset_place_kernel(&g_M, state, ctl.data(), mask, pending.data(), N, seed, env_offset, round == 0,
                                                       list0 ? elist[1].data() : nullptr, list0 ? &n0 : nullptr); });
    if (list0) { n_prev = n0; if (n0 == 0) return; }
    if (round == 0 && !list0) run_step_nd<ND>(state, o, N, g_M.settle_steps, 0, pending.data());
    else run_step_nd<ND>(state, o, N, g_M.settle_steps, 0, nullptr, elist[(round - 1) & 1].data(), n_prev);
    int n_pending = 0;
    int* lout = elist[round & 1].data();
    emu::launch(gf, bf, [&]() { prb_reset_finish_kernel<ND>(&g_M, state, o, ctl.data(), pending.data(), &n_pending, N, seed, env_offset, lout); });
    g_reset_rounds = round + 1;
    n_prev = n_pending;
    if (n_pending == 0) break;
  }
}
extern "C" {
int emu_reset_rounds() { return g_reset_rounds; }
int emu_set_model(const prb_model* m) { g_err = prb_convert_model(m, &g_M); return g_err.empty() ? 0 : -1; }
const char* emu_error() { return g_err.c_str(); }
int emu_state_stride() { return g_M.state_stride; }
int emu_state_dim() { return g_M.state_dim; }
int emu_warpmem_bytes() { return (int)sizeof(SetupMemT<SetupCfg>); }
int emu_warpmem_large_bytes() { return (int)sizeof(WarpMemT<CfgL>); }
int emu_devmodel_bytes() { return (int)sizeof(DevModel); }

void emu_init(float* state, int N) {
  emu_dim3 g, b; b.x = 128; g.x = (N + 127) / 128;
  emu::launch(g, b, [&]() { prb_init_kernel(&g_M, state, N); });
}
void emu_ik(float* state, const float* action, float* target, int N) {
  emu_dim3 g, b; b.x = 128; g.x = (N + 127) / 128;
  emu::launch(g, b, [&]() { prb_ik_kernel(&g_M, state, action, target, N); });
}
void emu_set_fused(int f) { g_fused = f; }
float* emu_sbuf() { return g_sbuf.data(); }
int emu_sbuf_q() { return SB_Q; }
void emu_class_counts(long long* out, int clear) { for (int k = 0; k < ARM_NCLASS; k++) { out[k] = g_class_count[k]; if (clear) g_class_count[k] = 0; } }
int emu_arm_nclass() { return ARM_NCLASS; }
// narrow-phase candidates (before manifold reduction) left in the scratch of warp `w` of the last launched block
int emu_last_candidates(int w, int n, float* out) {
  const SetupMemT<SetupCfg>& W = ((const SetupMemT<SetupCfg>*)g_emu_smem2)[w];
  for (int i = 0; i < n; i++) {
    const Contact& c = W.cand[i];
    float* o = out + 9 * i;
    o[0] = c.pbx; o[1] = c.pby; o[2] = c.pbz; o[3] = c.nx; o[4] = c.ny; o[5] = c.nz; o[6] = c.dist; o[7] = (float)(c.cols & 0xff); o[8] = (float)((c.cols >> 8) & 0xff);
  }
  return n;
}
// contacts the setup kernel generated for the env of warp `w` of the last launched block: {pb, n, dist, ca, cb} each (tests)
int emu_last_contacts(int w, float* out) {
  const SetupMemT<SetupCfg>& W = ((const SetupMemT<SetupCfg>*)g_emu_smem2)[w];
  for (int i = 0; i < W.n_contact; i++) {
    const Contact& c = W.ct[i];
    float* o = out + 9 * i;
    o[0] = c.pbx; o[1] = c.pby; o[2] = c.pbz; o[3] = c.nx; o[4] = c.ny; o[5] = c.nz; o[6] = c.dist; o[7] = (float)(c.cols & 0xff); o[8] = (float)((c.cols >> 8) & 0xff);
  }
  return W.n_contact;
}
int emu_arm_capq(int k) { return arm_capq(k); }
unsigned long long emu_overflow() { return g_overflow; }
void emu_substeps(float* state, int N, int nsub) { DevOut O; memset(&O, 0, sizeof(O)); O.overflow = &g_overflow; O.dbg = nullptr; O.ovf_env = nullptr; run_step(state, O, N, nsub, 0); }
void emu_step(float* state, const float* action, DevOut* O, int N) {
  O->overflow = &g_overflow; O->dbg = nullptr; O->ovf_env = nullptr;
  emu_ik(state, action, O->target_poses, N);
  run_step(state, *O, N, g_M.n_substeps, 1);
}
void emu_observe(float* state, DevOut* O, int N) { O->overflow = &g_overflow; O->dbg = nullptr; O->ovf_env = nullptr; run_step(state, *O, N, 0, 1); }
void emu_reset(float* state, DevOut* O, const unsigned char* mask, int N, unsigned long long seed, unsigned env_offset) {
  DevOut o = *O;
  o.overflow = &g_overflow; o.dbg = nullptr;
  if (g_M.nd == 12) emu_reset_nd<12>(state, o, mask, N, seed, env_offset);
  else emu_reset_nd<9>(state, o, mask, N, seed, env_offset);
}
void emu_reset_to(float* state, DevOut* O, const float* obs, const unsigned char* mask, int N, unsigned long long seed, unsigned env_offset, int restore_env) {
  DevOut o = *O;
  o.overflow = &g_overflow; o.dbg = nullptr; o.ovf_env = nullptr;
  emu_dim3 g, b; b.x = 32 * SetupCfg::WPB; g.x = (N + SetupCfg::WPB - 1) / SetupCfg::WPB;
  if (g_M.nd == 12) emu::launch(g, b, [&]() { prb_reset_to_kernel<12>(&g_M, state, o, obs, mask, N, seed, env_offset, restore_env); });
  else emu::launch(g, b, [&]() { prb_reset_to_kernel<9>(&g_M, state, o, obs, mask, N, seed, env_offset, restore_env); });
}
void emu_reward(const float* ag, const float* dg, long long B, float* out) {
  emu_dim3 g, b; b.x = 128; g.x = (unsigned)((B + 127) / 128);
  emu::launch(g, b, [&]() { prb_reward_kernel(&g_M, ag, dg, B, out); });
}
int emu_box_box(const float* p1, const float* R1, const float* h1, const float* p2, const float* R2, const float* h2, float* out) {
  CPoint c[4];
  int n = box_box(ld3(p1), ldm(R1), ld3(h1), ld3(p2), ldm(R2), ld3(h2), c);
  for (int i = 0; i < n; i++) { st3(out + 7 * i, c[i].pos); st3(out + 7 * i + 3, c[i].n); out[7 * i + 6] = c[i].depth; }
  return n;
}
}
