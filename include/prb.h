/* prb.h — C ABI of the B200-native batched playroom simulator (libprb_b200.so).
 *
 * The reference has no FFI/plugin API of its own: its boundary is the Python gym.GoalEnv
 * surface of `playEnv` (roboticsPlayroomPybullet/envs/environments.py:58-314), which talks to
 * PyBullet through ~60 C-API calls per step.  Every entry point below replaces one piece of
 * that surface for a whole batch of environments; the reference lines it stands in for are
 * cited per function.  The Python mirror (roboticsplayroompybullet_b200/envs.py) binds these
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: return 0 on success, a negative prb_status on failure (never throws across the
 * ABI); prb_last_error() gives the message.  All `*_dev` pointers are device pointers on the
 * handle's GPU; calls are asynchronous on the given CUDA stream (a cudaStream_t passed as
 * void*; NULL = legacy default stream) and serialised per handle; one handle per GPU/process.
 * No host synchronisation happens inside prb_step / prb_observe / prb_substeps / prb_set_goal; prb_reset
 * synchronises the stream (once per reset round, see below), like the reference's reset() it replaces.
 */
#ifndef PRB_H
#define PRB_H
#include <stdint.h>
#include "prb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct prb_handle prb_handle;

typedef enum prb_status {
  PRB_OK = 0,
  PRB_ERR_INVALID = -1,   /* bad argument / model not representable */
  PRB_ERR_CUDA = -2,      /* CUDA runtime failure (message has the cudaError string) */
  PRB_ERR_NO_DEVICE = -3  /* no CUDA device: there is no CPU fallback */
} prb_status;

typedef struct prb_config {
  int32_t num_envs;     /* environments owned by this handle (this rank's shard) */
  int32_t env_offset;   /* global index of local env 0: RNG streams are keyed by global env id */
  int32_t device;       /* CUDA device ordinal */
  int32_t reserved;
  uint64_t seed;        /* base seed of the counter-based per-env RNG (reset sampling) */
} prb_config;

/* Device buffers owned by the handle.  Each observation key of the reference's calc_state dict
 * (environments.py:849-861) is its own dense [num_envs, dim] fp32 array inside one contiguous
 * allocation `out_base` of `out_floats` floats (so one copy moves a whole step's results).
 * `state` is [num_envs, state_stride] fp32, see DESIGN.md "state layout". */
typedef struct prb_buffers {
  float* state;
  float* out_base;
  float* obs_quat;
  float* achieved_goal;
  float* desired_goal;
  float* controllable_achieved_goal;
  float* full_positional_state;
  float* joints;                   /* [N,8]  environments.py:758 */
  float* velocity;                 /* [N,6]  environments.py:857 */
  float* observation;              /* [N,6|12|18] environments.py:859 (dimension quirk kept) */
  float* gripper_proprioception;   /* [N]    environments.py:720-743 */
  float* reward;                   /* [N]    environments.py:211 */
  float* is_success;               /* [N]    environments.py:213 */
  float* target_poses;             /* [N,n_ik] environments.py:214,1034 */
  int64_t out_floats;
  int32_t num_envs, state_dim, state_stride;
  int32_t obs_dim, goal_dim, fps_dim, observation_dim, n_ik;
} prb_buffers;

/* playEnv.__init__ + activate_physics_client + instance.__init__ (environments.py:64-170,
 * 218-249, 321-454): builds the world for `num_envs` independent environments from a compiled
 * model.  All environments start in the load state (arm at q=0, default joint motors). */
int prb_create(const prb_model* model, const prb_config* cfg, prb_handle** out);
int prb_destroy(prb_handle* h);

/* playEnv.reset() (environments.py:173-187 -> instance.reset :599-603): for every env whose
 * mask byte is non-zero (NULL = all) re-seat objects, settle 100 substeps, reset the arm through
 * one IK call, sample a goal, and repeat while the sampled state already satisfies the goal.
 * Refreshes all output buffers of the reset envs.  Runs as rounds of {seat objects, settle on the masked step
 * pipeline, finish} over the envs still pending and reads one device counter per round (synchronous on `stream`);
 * prb_reset_rounds returns the number of rounds the most recent call needed. */
int prb_reset(prb_handle* h, const uint8_t* mask_dev, void* stream);
int prb_reset_rounds(prb_handle* h);

/* playEnv.reset(o) (environments.py:173-187, 541-556, 582-596): re-seat the masked envs FROM AN OBSERVATION — trajectory
 * replay.  obs_dev is [N, obs_dim] in the obs_quat layout: the object pose is taken from it (no settle steps), the arm
 * goes from its rest pose through one IK call to the observed end-effector pose, a new goal is drawn (again while already
 * satisfied).  Two documented differences from the reference: the object is read at its real offset in the layout (the
 * reference's hard-coded 11 / 10 indexing is wrong for the 19-D play layout), and with restore_env != 0 the drawer, door,
 * button and dial are restored from the observation too (the reference leaves them at their defaults).  Asynchronous. */
int prb_reset_to(prb_handle* h, const float* obs_dev, const uint8_t* mask_dev, int32_t restore_env, void* stream);

/* playEnv.reset_goal_pos(goal) (environments.py:190-191, 492-501): goal_dev is [N, goal_dim]. */
int prb_set_goal(prb_handle* h, const float* goal_dev, const uint8_t* mask_dev, void* stream);

/* playEnv.step(action) (environments.py:206-214): action_dev is [N,A]; A = 7 (8 for the quaternion decoders), the
 * decoder is the model's action_type parameter (environments.py:915-981).  Default [N,7] absolute xyz + rpy +
 * gripper; clip -> IK -> motor targets -> 12 substeps -> calc_state -> reward / is_success. */
int prb_step(prb_handle* h, const float* action_dev, void* stream);

/* instance.calc_state() (environments.py:799-864) without stepping: refresh the output buffers
 * from the current state (used after prb_set_state / prb_set_goal). */
int prb_observe(prb_handle* h, void* stream);

/* Run `n` raw stepSimulation() substeps without touching motors or outputs
 * (environments.py:490,535); used by the parity tests. */
int prb_substeps(prb_handle* h, int32_t n, void* stream);

int prb_get_buffers(prb_handle* h, prb_buffers* out);

/* playEnv.compute_reward / compute_reward_sparse (environments.py:278-304,
 * playRewardFunc.py:66-77), stateless and batched for goal relabelling: ag, dg are [B, goal_dim]. */
int prb_compute_reward(prb_handle* h, const float* ag_dev, const float* dg_dev, int64_t B, float* out_dev, void* stream);

/* Raw simulation state (host copies, synchronous): [N, state_dim] fp32.  The reference has no
 * save/restore; these exist so tests can start oracle and device from identical
 * states, and double as checkpoint/restore. */
int prb_get_state(prb_handle* h, float* host_out);
int prb_set_state(prb_handle* h, const float* host_in);

/* Same as prb_step, but through HOST buffers: copies action_host [N,A] to the device, steps,
 * copies the whole output block (out_floats floats) back into out_host and synchronises the
 * stream.  This is the call the Python gym mirror makes for numpy in / numpy out stepping.
 * out_host may be pinned (one asynchronous copy) or ordinary pageable memory, e.g. the fresh array a
 * gym step returns: then the block comes back through the handle's pinned staging in chunks that
 * PRB_HOST_THREADS worker threads (default: 8 on hosts with >= 16 cores) copy out while the next chunk
 * is still in flight. */
int prb_step_host(prb_handle* h, const float* action_host, float* out_host, void* stream);

/* Per-kernel device timing for reports: when enabled, prb_step brackets its two kernels with
 * CUDA events on the launching stream; prb_last_kernel_ms synchronises on them and returns the
 * durations of the most recent step's IK kernel and fused substep kernel. */
int prb_enable_kernel_timing(prb_handle* h, int32_t enable);
int prb_last_kernel_ms(prb_handle* h, float* ik_ms, float* step_ms);
/* split of step_ms into the summed durations of the warp-per-env setup launches and the
 * thread-per-env solver launches of the most recent step */
int prb_last_tier_ms(prb_handle* h, float* setup_ms, float* pgs_ms);

/* Number of kernels launched by this handle since creation (bench.py reports it). */
int64_t prb_launch_count(prb_handle* h);
/* Env steps (since creation; a reset counts as one) in which a contact had to be dropped because the per-env
 * capacity (overlapping pairs, contacts, joint rows) was exceeded: an env is counted once per step however many of
 * its 12 substeps dropped one.  Synchronises the device. */
int64_t prb_overflow_count(prb_handle* h);
/* Per-env facts of the last substep built, for sizing reports and for tests that must sample every solver path:
 * [N,4] int32 = {arm-island solver class (0: joint-row kernel, 1..4: size class of the arm-island kernel) | q count of
 * the arm island's region << 8, contacts, q count of the env's whole record stream, joint rows | capacity flags << 8
 * (1: more than 32 overlapping collider pairs, 2: more than 32 contacts, 4: contact between two slide bodies)}; synchronous. */
int prb_debug_usage(prb_handle* h, int32_t* host_out);
/* Static facts of the step kernel for reports: dynamic shared memory per block, warps per block. */
int prb_kernel_info(prb_handle* h, int32_t* smem_bytes_per_block, int32_t* envs_per_block, int32_t* regs_per_thread);

const char* prb_last_error(prb_handle* h);
const char* prb_version(void);

#ifdef __cplusplus
}
#endif
#endif
