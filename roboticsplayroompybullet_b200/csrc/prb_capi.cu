// prb_capi.cu — the C-ABI shared library (include/prb.h) around the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>

#include <string>

#include "../../include/prb.h"
#include "prb_convert.h"
#include "prb_kernels.cuh"

struct prb_handle {
  DevModel hm;
  DevModel* dm = nullptr;
  int N = 0, device = 0;
  unsigned env_offset = 0;
  unsigned long long seed = 0;
  float* state = nullptr;
  float* out = nullptr;
  float* action_stage = nullptr;   // device staging for prb_step_host
  int64_t out_floats = 0;
  DevOut O;
  int64_t launches = 0;
  int smem = 0, regs = 0;          // small tier (reported)
  int smem_large = 0, regs_large = 0, smem_reset = 0;
  int* redo_list = nullptr;        // [2][N] envs handed from the small to the medium / medium to the large tier
  int* redo_count = nullptr;       // [2]
  int smem_medium = 0;
  int timing = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
};

static std::string g_err;  // errors before a handle exists

#define CK(h, call)                                                              \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return PRB_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

// One env step (or n raw substeps) = small tier over every env, then the large tier over the envs
// the small tier marked.  Two launches, no host synchronisation in between.
template <int ND>
static int launch_step(prb_handle* h, int nsub, int observe, cudaStream_t s) {
  CK(h, cudaMemsetAsync(h->redo_count, 0, 2 * sizeof(int), s));
  int* l1 = h->redo_list; int* l2 = h->redo_list + h->N;
  int* c1 = h->redo_count; int* c2 = h->redo_count + 1;
  dim3 gs((h->N + CfgS::WPB - 1) / CfgS::WPB), bs(32 * CfgS::WPB);
  prb_step_kernel<ND, CfgS><<<gs, bs, h->smem, s>>>(h->dm, h->state, h->O, nullptr, nullptr, l1, c1, h->N, nsub, observe);
  h->launches++;
  CK(h, cudaGetLastError());
  if (h->timing) CK(h, cudaEventRecord(h->ev[3], s));
  if (nsub > 0) {
    dim3 gm((h->N + CfgM::WPB - 1) / CfgM::WPB), bm(32 * CfgM::WPB);
    prb_step_kernel<ND, CfgM><<<gm, bm, h->smem_medium, s>>>(h->dm, h->state, h->O, l1, c1, l2, c2, h->N, nsub, observe);
    dim3 gl((h->N + CfgL::WPB - 1) / CfgL::WPB), bl(32 * CfgL::WPB);
    prb_step_kernel<ND, CfgL><<<gl, bl, h->smem_large, s>>>(h->dm, h->state, h->O, l2, c2, nullptr, nullptr, h->N, nsub, observe);
    h->launches += 2;
    CK(h, cudaGetLastError());
  }
  return PRB_OK;
}
static int run_step(prb_handle* h, int nsub, int observe, cudaStream_t s) {
  return h->hm.nd == 12 ? launch_step<12>(h, nsub, observe, s) : launch_step<9>(h, nsub, observe, s);
}

template <int ND>
static int setup_kernels(prb_handle* h) {
  h->smem = CfgS::WPB * (int)sizeof(WarpMemT<CfgS>);
  h->smem_large = CfgL::WPB * (int)sizeof(WarpMemT<CfgL>);
  h->smem_reset = (int)sizeof(WarpMemT<CfgL>);
  h->smem_medium = CfgM::WPB * (int)sizeof(WarpMemT<CfgM>);
  CK(h, cudaFuncSetAttribute(prb_step_kernel<ND, CfgM>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_medium));
  CK(h, cudaFuncSetAttribute(prb_step_kernel<ND, CfgS>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
  CK(h, cudaFuncSetAttribute(prb_step_kernel<ND, CfgL>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_large));
  CK(h, cudaFuncSetAttribute(prb_reset_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_reset));
  cudaFuncAttributes fa;
  CK(h, cudaFuncGetAttributes(&fa, prb_step_kernel<ND, CfgS>));
  h->regs = fa.numRegs;
  CK(h, cudaFuncGetAttributes(&fa, prb_step_kernel<ND, CfgL>));
  h->regs_large = fa.numRegs;
  return PRB_OK;
}

extern "C" {

const char* prb_version(void) { return "prb_b200 0.1 (sm_100a)"; }

const char* prb_last_error(prb_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int prb_create(const prb_model* model, const prb_config* cfg, prb_handle** out) {
  if (!model || !cfg || !out || cfg->num_envs <= 0) { g_err = "prb_create: bad arguments"; return PRB_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_err = "prb_create: no CUDA device (this library has no CPU fallback)";
    return PRB_ERR_NO_DEVICE;
  }
  prb_handle* h = new prb_handle();
  std::string why = prb_convert_model(model, &h->hm);
  if (!why.empty()) { g_err = "prb_create: " + why; delete h; return PRB_ERR_INVALID; }
  if (h->hm.nd != 12 && h->hm.nd != 9) { g_err = "prb_create: arm must have 12 (UR5+Robotiq) or 9 (Panda) DoF"; delete h; return PRB_ERR_INVALID; }
  h->N = cfg->num_envs; h->device = cfg->device; h->env_offset = (unsigned)cfg->env_offset; h->seed = cfg->seed;
  *out = h;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMalloc(&h->dm, sizeof(DevModel)));
  CK(h, cudaMemcpy(h->dm, &h->hm, sizeof(DevModel), cudaMemcpyHostToDevice));
  const DevModel& M = h->hm;
  const int64_t N = h->N;
  CK(h, cudaMalloc(&h->state, sizeof(float) * N * M.state_stride));
  const int dims[12] = {M.obs_dim, M.goal_dim, M.goal_dim, 4, M.fps_dim, 8, 6, M.observation_dim, 1, 1, 1, M.n_ik};
  int64_t tot = 0;
  for (int i = 0; i < 12; i++) tot += N * dims[i];
  h->out_floats = tot;
  CK(h, cudaMalloc(&h->out, sizeof(float) * tot));
  CK(h, cudaMemset(h->out, 0, sizeof(float) * tot));
  CK(h, cudaMalloc(&h->action_stage, sizeof(float) * N * 7));
  CK(h, cudaMalloc(&h->O.overflow, sizeof(unsigned long long)));
  CK(h, cudaMemset(h->O.overflow, 0, sizeof(unsigned long long)));
  CK(h, cudaMalloc(&h->O.dbg, sizeof(int) * 4 * N));
  CK(h, cudaMemset(h->O.dbg, 0, sizeof(int) * 4 * N));
  float** slots[12] = {&h->O.obs_quat, &h->O.achieved_goal, &h->O.desired_goal, &h->O.cag, &h->O.fps, &h->O.joints,
                       &h->O.velocity, &h->O.observation, &h->O.proprio, &h->O.reward, &h->O.success, &h->O.target_poses};
  int64_t off = 0;
  for (int i = 0; i < 12; i++) { *slots[i] = h->out + off; off += N * dims[i]; }
  CK(h, cudaMalloc(&h->redo_list, sizeof(int) * 2 * N));
  CK(h, cudaMalloc(&h->redo_count, 2 * sizeof(int)));
  CK(h, cudaMemset(h->redo_count, 0, 2 * sizeof(int)));
  {
    int rc = M.nd == 12 ? setup_kernels<12>(h) : setup_kernels<9>(h);
    if (rc != PRB_OK) return rc;
  }
  prb_init_kernel<<<(h->N + 127) / 128, 128>>>(h->dm, h->state, h->N);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  return PRB_OK;
}

int prb_destroy(prb_handle* h) {
  if (!h) return PRB_ERR_INVALID;
  cudaSetDevice(h->device);
  cudaFree(h->dm); cudaFree(h->state); cudaFree(h->out); cudaFree(h->action_stage); cudaFree(h->O.overflow); cudaFree(h->O.dbg); cudaFree(h->redo_list); cudaFree(h->redo_count);
  delete h;
  return PRB_OK;
}

int prb_get_buffers(prb_handle* h, prb_buffers* b) {
  if (!h || !b) return PRB_ERR_INVALID;
  const DevModel& M = h->hm;
  b->state = h->state; b->out_base = h->out;
  b->obs_quat = h->O.obs_quat; b->achieved_goal = h->O.achieved_goal; b->desired_goal = h->O.desired_goal;
  b->controllable_achieved_goal = h->O.cag; b->full_positional_state = h->O.fps; b->joints = h->O.joints;
  b->velocity = h->O.velocity; b->observation = h->O.observation; b->gripper_proprioception = h->O.proprio;
  b->reward = h->O.reward; b->is_success = h->O.success; b->target_poses = h->O.target_poses;
  b->out_floats = h->out_floats; b->num_envs = h->N; b->state_dim = M.state_dim; b->state_stride = M.state_stride;
  b->obs_dim = M.obs_dim; b->goal_dim = M.goal_dim; b->fps_dim = M.fps_dim; b->observation_dim = M.observation_dim; b->n_ik = M.n_ik;
  return PRB_OK;
}

int prb_step(prb_handle* h, const float* action_dev, void* stream) {
  if (!h || !action_dev) return PRB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->timing) CK(h, cudaEventRecord(h->ev[0], s));
  prb_ik_kernel<<<(h->N + 127) / 128, 128, 0, s>>>(h->dm, h->state, action_dev, h->O.target_poses, h->N);
  h->launches++;
  CK(h, cudaGetLastError());
  if (h->timing) CK(h, cudaEventRecord(h->ev[1], s));
  int rc = run_step(h, h->hm.n_substeps, 1, s);
  if (h->timing) CK(h, cudaEventRecord(h->ev[2], s));
  return rc;
}

int prb_enable_kernel_timing(prb_handle* h, int32_t enable) {
  if (!h) return PRB_ERR_INVALID;
  if (enable && !h->ev[0]) for (int i = 0; i < 4; i++) CK(h, cudaEventCreate(&h->ev[i]));
  h->timing = enable ? 1 : 0;
  return PRB_OK;
}

int prb_last_kernel_ms(prb_handle* h, float* ik_ms, float* step_ms) {
  if (!h || !h->ev[0]) return PRB_ERR_INVALID;
  CK(h, cudaEventSynchronize(h->ev[2]));
  if (ik_ms) CK(h, cudaEventElapsedTime(ik_ms, h->ev[0], h->ev[1]));
  if (step_ms) CK(h, cudaEventElapsedTime(step_ms, h->ev[1], h->ev[2]));
  return PRB_OK;
}

int prb_last_tier_ms(prb_handle* h, float* small_ms, float* large_ms) {
  if (!h || !h->ev[0]) return PRB_ERR_INVALID;
  CK(h, cudaEventSynchronize(h->ev[2]));
  if (small_ms) CK(h, cudaEventElapsedTime(small_ms, h->ev[1], h->ev[3]));
  if (large_ms) CK(h, cudaEventElapsedTime(large_ms, h->ev[3], h->ev[2]));
  return PRB_OK;
}

int prb_observe(prb_handle* h, void* stream) {
  if (!h) return PRB_ERR_INVALID;
  return run_step(h, 0, 1, (cudaStream_t)stream);
}

int prb_substeps(prb_handle* h, int32_t n, void* stream) {
  if (!h || n < 0) return PRB_ERR_INVALID;
  return run_step(h, n, 0, (cudaStream_t)stream);
}

int prb_reset(prb_handle* h, const uint8_t* mask_dev, void* stream) {
  if (!h) return PRB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(h->N), block(32);
  if (h->hm.nd == 12) prb_reset_kernel<12><<<grid, block, h->smem_reset, s>>>(h->dm, h->state, h->O, mask_dev, h->N, h->seed, h->env_offset);
  else prb_reset_kernel<9><<<grid, block, h->smem_reset, s>>>(h->dm, h->state, h->O, mask_dev, h->N, h->seed, h->env_offset);
  h->launches++;
  CK(h, cudaGetLastError());
  return PRB_OK;
}

__global__ void prb_set_goal_kernel(const DevModel* __restrict__ Mp, float* __restrict__ state, const float* __restrict__ goal,
                                    const unsigned char* __restrict__ mask, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const DevModel& M = *Mp;
  if (i >= N * M.goal_dim) return;
  int e = i / M.goal_dim, k = i % M.goal_dim;
  if (mask && !mask[e]) return;
  state[(size_t)e * M.state_stride + 5 * M.nd + 13 * M.n_free + 2 * M.n_slide + k] = goal[i];
}

int prb_set_goal(prb_handle* h, const float* goal_dev, const uint8_t* mask_dev, void* stream) {
  if (!h || !goal_dev) return PRB_ERR_INVALID;
  int n = h->N * h->hm.goal_dim;
  prb_set_goal_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->dm, h->state, goal_dev, mask_dev, h->N);
  h->launches++;
  CK(h, cudaGetLastError());
  return PRB_OK;
}

int prb_compute_reward(prb_handle* h, const float* ag_dev, const float* dg_dev, int64_t B, float* out_dev, void* stream) {
  if (!h || !ag_dev || !dg_dev || !out_dev || B < 0) return PRB_ERR_INVALID;
  if (B == 0) return PRB_OK;
  prb_reward_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->dm, ag_dev, dg_dev, (long long)B, out_dev);
  h->launches++;
  CK(h, cudaGetLastError());
  return PRB_OK;
}

int prb_get_state(prb_handle* h, float* host_out) {
  if (!h || !host_out) return PRB_ERR_INVALID;
  const DevModel& M = h->hm;
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy2D(host_out, sizeof(float) * M.state_dim, h->state, sizeof(float) * M.state_stride,
                     sizeof(float) * M.state_dim, h->N, cudaMemcpyDeviceToHost));
  return PRB_OK;
}

int prb_set_state(prb_handle* h, const float* host_in) {
  if (!h || !host_in) return PRB_ERR_INVALID;
  const DevModel& M = h->hm;
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy2D(h->state, sizeof(float) * M.state_stride, host_in, sizeof(float) * M.state_dim,
                     sizeof(float) * M.state_dim, h->N, cudaMemcpyHostToDevice));
  return PRB_OK;
}

int prb_step_host(prb_handle* h, const float* action_host, float* out_host, void* stream) {
  if (!h || !action_host || !out_host) return PRB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  CK(h, cudaMemcpyAsync(h->action_stage, action_host, sizeof(float) * h->N * 7, cudaMemcpyHostToDevice, s));
  int rc = prb_step(h, h->action_stage, stream);
  if (rc != PRB_OK) return rc;
  CK(h, cudaMemcpyAsync(out_host, h->out, sizeof(float) * h->out_floats, cudaMemcpyDeviceToHost, s));
  CK(h, cudaStreamSynchronize(s));
  return PRB_OK;
}

int64_t prb_launch_count(prb_handle* h) { return h ? h->launches : 0; }

int prb_debug_usage(prb_handle* h, int32_t* host_out /* [N,4] */) {
  if (!h || !host_out) return PRB_ERR_INVALID;
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(host_out, h->O.dbg, sizeof(int) * 4 * h->N, cudaMemcpyDeviceToHost));
  return PRB_OK;
}

int64_t prb_overflow_count(prb_handle* h) {
  if (!h) return -1;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, h->O.overflow, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

int prb_kernel_info(prb_handle* h, int32_t* smem_bytes_per_block, int32_t* envs_per_block, int32_t* regs_per_thread) {
  if (!h) return PRB_ERR_INVALID;
  if (smem_bytes_per_block) *smem_bytes_per_block = h->smem;
  if (envs_per_block) *envs_per_block = CfgS::WPB;
  if (regs_per_thread) *regs_per_thread = h->regs;
  return PRB_OK;
}

}  // extern "C"
