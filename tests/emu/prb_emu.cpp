// prb_emu.cpp — runs the product's kernel source under the CPU SIMT emulator (tests only).
#include "cuda_emu.h"
#include "../../roboticsplayroompybullet_b200/csrc/prb_kernels.cuh"
#include "../../roboticsplayroompybullet_b200/csrc/prb_convert.h"

static DevModel g_M;
static std::string g_err;

extern "C" {
int emu_set_model(const prb_model* m) { g_err = prb_convert_model(m, &g_M); return g_err.empty() ? 0 : -1; }
const char* emu_error() { return g_err.c_str(); }
int emu_state_stride() { return g_M.state_stride; }
int emu_state_dim() { return g_M.state_dim; }
int emu_warpmem_bytes() { return (int)sizeof(WarpMem); }
int emu_devmodel_bytes() { return (int)sizeof(DevModel); }

void emu_init(float* state, int N) {
  emu_dim3 g, b; b.x = 128; g.x = (N + 127) / 128;
  emu::launch(g, b, [&]() { prb_init_kernel(&g_M, state, N); });
}
void emu_ik(float* state, const float* action, float* target, int N) {
  emu_dim3 g, b; b.x = 128; g.x = (N + 127) / 128;
  emu::launch(g, b, [&]() { prb_ik_kernel(&g_M, state, action, target, N); });
}
static void run_step(float* state, DevOut O, int N, int nsub, int observe) {
  emu_dim3 g, b; b.x = 32 * PRB_WPB; g.x = (N + PRB_WPB - 1) / PRB_WPB;
  if (g_M.nd == 12) emu::launch(g, b, [&]() { prb_step_kernel<12>(&g_M, state, O, N, nsub, observe); });
  else emu::launch(g, b, [&]() { prb_step_kernel<9>(&g_M, state, O, N, nsub, observe); });
}
static unsigned long long g_overflow = 0;
unsigned long long emu_overflow() { return g_overflow; }
void emu_substeps(float* state, int N, int nsub) { DevOut O; memset(&O, 0, sizeof(O)); O.overflow = &g_overflow; run_step(state, O, N, nsub, 0); }
void emu_step(float* state, const float* action, DevOut* O, int N) {
  O->overflow = &g_overflow;
  emu_ik(state, action, O->target_poses, N);
  run_step(state, *O, N, g_M.n_substeps, 1);
}
void emu_observe(float* state, DevOut* O, int N) { O->overflow = &g_overflow; run_step(state, *O, N, 0, 1); }
void emu_reset(float* state, DevOut* O, const unsigned char* mask, int N, unsigned long long seed, unsigned env_offset) {
  emu_dim3 g, b; b.x = 32 * PRB_WPB; g.x = (N + PRB_WPB - 1) / PRB_WPB;
  DevOut o = *O;
  o.overflow = &g_overflow;
  if (g_M.nd == 12) emu::launch(g, b, [&]() { prb_reset_kernel<12>(&g_M, state, o, mask, N, seed, env_offset); });
  else emu::launch(g, b, [&]() { prb_reset_kernel<9>(&g_M, state, o, mask, N, seed, env_offset); });
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
