"""Vectorised mirror of the reference's gym surface (roboticsPlayroomPybullet/envs/
environments.py:58-314 `playEnv`, envList.py:18-22, 89-99) on top of the C-ABI.

`VecPlayEnv` keeps the reference's method names, argument meaning, dict keys and per-key
layouts, with a leading [num_envs] dimension:

    reset(mask=None) -> obs dict          (playEnv.reset, :173-187)
    step(action[N,7]) -> (obs, r[N], done[N], info)   (playEnv.step, :206-214)
    reset_goal_pos(goal[N,G])             (:190-191)
    compute_reward(ag, dg[, info])        (:274-304; sparse variant bound when sparse=True, :169-170)
    _max_episode_steps, action_space / observation_space bounds as arrays

Two calling styles: numpy in / numpy out (`step`), which goes through host buffers exactly like a
reference user would (H2D of actions, D2H of the whole observation block every step), and
`step_device` for learners that keep everything on the GPU (torch tensors, zero-copy views of the
library's buffers).
"""
import ctypes

import numpy as np

from . import lib as _lib
from .model import ACTION_VARIANTS, ENV_KINDS, action_dim, load_model

OUT_KEYS = ['obs_quat', 'achieved_goal', 'desired_goal', 'controllable_achieved_goal', 'full_positional_state',
            'joints', 'velocity', 'observation', 'gripper_proprioception', 'reward', 'is_success', 'target_poses']
OBS_KEYS = OUT_KEYS[:9]


class _DevArray:
    """Exposes a library-owned device buffer through __cuda_array_interface__ (zero-copy)."""

    def __init__(self, ptr, shape, owner):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}
        self._owner = owner


class VecPlayEnv:
    env_id = None

    def __init__(self, env_id=None, num_envs=1, device=0, seed=1234, env_offset=0, sparse=True, model=None):
        import torch  # device memory / streams only
        self.torch = torch
        env_id = env_id or self.env_id
        if not sparse:
            # The reference binds compute_reward once and uses it in step() AND in the reset retry loop
            # (environments.py:169-170, 185, 210); with the dense reward that loop only ends when the sampled goal
            # is >= 1 m away.  No registered id uses it; the in-kernel reward is the sparse one.
            raise NotImplementedError('sparse=False (dense reward inside step/reset) is not supported; '
                                      'use dense_reward(ag, dg) for the dense metric')
        if env_id not in ENV_KINDS and env_id not in ACTION_VARIANTS:
            raise NotImplementedError(env_id)           # environments.py:376,416,933
        self.env_id = env_id
        self.model = model if model is not None else load_model(env_id)
        m = self.model
        self.num_envs = int(num_envs)
        self.device = torch.device('cuda', device)
        self.L = _lib.load()
        self._ms = m.as_struct()
        cfg = _lib.PrbConfig(self.num_envs, int(env_offset), int(device), 0, int(seed))
        self._h = ctypes.c_void_p()
        rc = self.L.prb_create(ctypes.byref(self._ms), ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise _lib.PrbError('prb_create failed (%d): %s' % (rc, self.L.prb_last_error(self._h).decode()))
        b = _lib.PrbBuffers()
        _lib.check(self.L, self._h, self.L.prb_get_buffers(self._h, ctypes.byref(b)))
        self.buffers = b
        N = self.num_envs
        self.dims = {'obs_quat': b.obs_dim, 'achieved_goal': b.goal_dim, 'desired_goal': b.goal_dim,
                     'controllable_achieved_goal': 4, 'full_positional_state': b.fps_dim, 'joints': 8,
                     'velocity': 6, 'observation': b.observation_dim, 'gripper_proprioception': 1, 'reward': 1,
                     'is_success': 1, 'target_poses': b.n_ik}
        self.out_floats = int(b.out_floats)
        with torch.cuda.device(self.device):
            self.dev = {k: torch.as_tensor(_DevArray(getattr(b, k), (N, self.dims[k]), self), device=self.device)
                        for k in OUT_KEYS}
            self.state_dev = torch.as_tensor(_DevArray(b.state, (N, b.state_stride), self), device=self.device)
        # pinned host staging for the numpy path
        self.action_dim = action_dim(self.model)
        self._h_action = torch.empty((N, self.action_dim), dtype=torch.float32).pin_memory()
        self._h_out = torch.empty((self.out_floats,), dtype=torch.float32).pin_memory()
        self._h_views = {}
        self._h_slices = {}
        off = 0
        for k in OUT_KEYS:
            n = N * self.dims[k]
            self._h_views[k] = self._h_out[off:off + n].view(N, self.dims[k]).numpy()
            self._h_slices[k] = (off, off + n)
            off += n
        self._h_out_np = self._h_out.numpy()
        # gym-like metadata (environments.py:84,108-117)
        self._max_episode_steps = None if m['play'] else 250
        high = np.array([m.param('action_high_xyz')] * (self.action_dim - 1) + [m.param('action_high_grip')], np.float32)   # :88-112
        self.action_low, self.action_high = -high, high
        self.play = bool(m['play'])
        self.sparse = sparse
        self.h2d_bytes_per_step = N * self.action_dim * 4
        self.d2h_bytes_per_step = self.out_floats * 4

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_dev(self):
        return {k: self.dev[k] for k in OBS_KEYS}

    def _host_copy(self):
        """Fresh arrays for the caller (the reference returns copies, environments.py:850-855): ONE copy of the pinned result
        block, the per-key arrays are views of it."""
        buf = self._h_out_np.copy()
        return {k: buf[a:b].reshape(self.num_envs, self.dims[k]) for k, (a, b) in self._h_slices.items()}

    def _obs_host(self, c=None):
        c = c if c is not None else self._host_copy()
        d = {k: c[k] for k in OBS_KEYS}
        d['gripper_proprioception'] = d['gripper_proprioception'][:, 0].astype(np.int64)
        d['img'] = None                                  # environments.py:844-845
        return d

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self.L.prb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device-resident API
    def reset_device(self, mask=None, o=None, restore_env=True):
        """reset() of the masked envs (all when mask is None).  With `o` ([N, obs_dim] float32 CUDA tensor in the obs_quat
        layout) the envs are re-seated from that observation instead of sampled: playEnv.reset(o), trajectory replay."""
        mp = None
        if mask is not None:
            self._mask = mask.to(self.device, self.torch.uint8).contiguous()
            mp = ctypes.c_void_p(self._mask.data_ptr())
        if o is not None:
            ob = o.to(self.device, self.torch.float32).contiguous()
            assert ob.shape == (self.num_envs, self.dims['obs_quat']), 'o must be [num_envs, %d] (obs_quat layout)' % self.dims['obs_quat']
            self._o_keep = ob
            _lib.check(self.L, self._h, self.L.prb_reset_to(self._h, ctypes.c_void_p(ob.data_ptr()), mp, 1 if restore_env else 0, self._stream()))
        else:
            _lib.check(self.L, self._h, self.L.prb_reset(self._h, mp, self._stream()))
        return self._obs_dev()

    def step_device(self, action):
        """action: float32 CUDA tensor [N,7].  Returns views of the library's output buffers."""
        a = action.to(self.device, self.torch.float32).contiguous()
        assert a.shape == (self.num_envs, self.action_dim), 'action must be [num_envs, %d]' % self.action_dim   # environments.py:937,956
        self._a_keep = a
        _lib.check(self.L, self._h, self.L.prb_step(self._h, ctypes.c_void_p(a.data_ptr()), self._stream()))
        info = {'is_success': self.dev['is_success'][:, 0], 'target_poses': self.dev['target_poses']}
        return self._obs_dev(), self.dev['reward'][:, 0], None, info

    def observe_device(self):
        _lib.check(self.L, self._h, self.L.prb_observe(self._h, self._stream()))
        return self._obs_dev()

    # ------------------------------------------------------------------ reference-style (numpy) API
    def reset(self, o=None, mask=None, restore_env=True):
        """playEnv.reset(o=None) (environments.py:173-187) for every env, or for the envs of `mask`.  `o`: [N, obs_dim]
        obs_quat rows to re-seat the envs from (trajectory replay) instead of sampling."""
        mt = None
        if mask is not None:
            mt = self.torch.as_tensor(np.asarray(mask, np.uint8))
        ot = None
        if o is not None:
            ot = self.torch.as_tensor(np.ascontiguousarray(np.asarray(o, np.float32)).reshape(self.num_envs, -1))
        self.reset_device(mt, ot, restore_env)
        self.torch.cuda.current_stream(self.device).synchronize()
        self._pull()
        return self._obs_host()

    def _pull(self):
        t = self.torch
        with t.cuda.device(self.device):
            src = t.as_tensor(_DevArray(self.buffers.out_base, (self.out_floats,), self), device=self.device)
            self._h_out.copy_(src, non_blocking=True)
            t.cuda.current_stream(self.device).synchronize()

    def step(self, action):
        a = np.asarray(action, np.float32)
        assert a.shape == (self.num_envs, self.action_dim), 'action must be [num_envs, %d]' % self.action_dim
        self._h_action.numpy()[...] = a
        # the result block lands in a fresh array (the reference returns new arrays every step, environments.py:850-855):
        # prb_step_host reads it back through pinned staging in chunks and copies them out with worker threads
        buf = np.empty(self.out_floats, np.float32)
        _lib.check(self.L, self._h, self.L.prb_step_host(self._h, ctypes.c_void_p(self._h_action.data_ptr()),
                                                        ctypes.c_void_p(buf.ctypes.data), self._stream()))
        c = {k: buf[i:j].reshape(self.num_envs, self.dims[k]) for k, (i, j) in self._h_slices.items()}
        r = c['reward'][:, 0]
        info = {'is_success': c['is_success'][:, 0].astype(np.int64), 'target_poses': c['target_poses']}
        done = np.zeros(self.num_envs, dtype=bool)       # environments.py:212: always False
        return self._obs_host(c), r, done, info

    def reset_goal_pos(self, goal):
        t = self.torch
        g = t.as_tensor(np.asarray(goal, np.float32)).reshape(self.num_envs, -1).to(self.device).contiguous()
        assert g.shape == (self.num_envs, self.dims['desired_goal']), 'goal must be [num_envs, %d]' % self.dims['desired_goal']
        self._g_keep = g
        _lib.check(self.L, self._h, self.L.prb_set_goal(self._h, ctypes.c_void_p(g.data_ptr()), None, self._stream()))

    def reset_goal_pos_device(self, goal, mask=None):
        """Device-resident reset_goal_pos (environments.py:190-191) for goal relabelling inside a rollout loop:
        goal float32 CUDA tensor [N, goal_dim], optional uint8 mask [N]."""
        t = self.torch
        g = goal.to(self.device, t.float32).reshape(self.num_envs, -1).contiguous()
        assert g.shape == (self.num_envs, self.dims['desired_goal']), 'goal must be [num_envs, %d]' % self.dims['desired_goal']
        self._g_keep = g
        mp = None
        if mask is not None:
            self._gmask = mask.to(self.device, t.uint8).contiguous()
            mp = ctypes.c_void_p(self._gmask.data_ptr())
        _lib.check(self.L, self._h, self.L.prb_set_goal(self._h, ctypes.c_void_p(g.data_ptr()), mp, self._stream()))

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        """Batched compute_reward_sparse (sparse=True) or dense -||ag-dg|| (environments.py:274-304)."""
        t = self.torch
        ag = np.asarray(achieved_goal, np.float32)
        dg = np.asarray(desired_goal, np.float32)
        single = ag.ndim == 1
        G = self.dims['achieved_goal']
        assert ag.shape[-1] == G and dg.shape == ag.shape, 'goals must be [..., %d]' % G
        agd = t.as_tensor(ag.reshape(-1, G)).to(self.device).contiguous()
        dgd = t.as_tensor(dg.reshape(-1, G)).to(self.device).contiguous()
        out = t.empty(agd.shape[0], dtype=t.float32, device=self.device)
        _lib.check(self.L, self._h, self.L.prb_compute_reward(self._h, ctypes.c_void_p(agd.data_ptr()),
                                                             ctypes.c_void_p(dgd.data_ptr()), agd.shape[0],
                                                             ctypes.c_void_p(out.data_ptr()), self._stream()))
        r = out.cpu().numpy()
        return r[0] if single else r

    compute_reward_sparse = compute_reward

    def dense_reward(self, achieved_goal, desired_goal):
        """playEnv.compute_reward before the sparse rebinding (environments.py:269-275): -||ag - dg|| (host numpy)."""
        G = self.dims['achieved_goal']
        ag = np.asarray(achieved_goal, np.float32)
        d = -np.linalg.norm(ag.reshape(-1, G) - np.asarray(desired_goal, np.float32).reshape(-1, G), axis=1)
        return d[0] if ag.ndim == 1 else d

    def compute_reward_device(self, ag, dg):
        t = self.torch
        G = self.dims['achieved_goal']
        agd = ag.reshape(-1, G).to(self.device, t.float32).contiguous()
        dgd = dg.reshape(-1, G).to(self.device, t.float32).contiguous()
        out = t.empty(agd.shape[0], dtype=t.float32, device=self.device)
        _lib.check(self.L, self._h, self.L.prb_compute_reward(self._h, ctypes.c_void_p(agd.data_ptr()),
                                                             ctypes.c_void_p(dgd.data_ptr()), agd.shape[0],
                                                             ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ raw state (tests, checkpoints)
    def get_state(self):
        out = np.zeros((self.num_envs, self.buffers.state_dim), np.float32)
        _lib.check(self.L, self._h, self.L.prb_get_state(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def set_state(self, state):
        s = np.ascontiguousarray(state, np.float32)
        assert s.shape == (self.num_envs, self.buffers.state_dim)
        _lib.check(self.L, self._h, self.L.prb_set_state(self._h, s.ctypes.data_as(ctypes.c_void_p)))

    def substeps(self, n):
        _lib.check(self.L, self._h, self.L.prb_substeps(self._h, int(n), self._stream()))

    def observe(self):
        self.observe_device()
        self._pull()
        return self._obs_host()

    def enable_kernel_timing(self, on=True):
        _lib.check(self.L, self._h, self.L.prb_enable_kernel_timing(self._h, 1 if on else 0))

    def last_kernel_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        _lib.check(self.L, self._h, self.L.prb_last_kernel_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def last_tier_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        _lib.check(self.L, self._h, self.L.prb_last_tier_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def reset_rounds(self):
        """Rounds (seat objects, settle, finish) the most recent reset needed."""
        return int(self.L.prb_reset_rounds(self._h))

    def launch_count(self):
        return int(self.L.prb_launch_count(self._h))

    def debug_usage(self):
        out = np.zeros((self.num_envs, 4), np.int32)
        _lib.check(self.L, self._h, self.L.prb_debug_usage(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def overflow_count(self):
        return int(self.L.prb_overflow_count(self._h))

    def kernel_info(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self.L.prb_kernel_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {'smem_bytes_per_block': a.value, 'envs_per_block': b.value, 'regs_per_thread': c.value}


# the three registered ids in scope (roboticsPlayroomPybullet/__init__.py:23-26,65-68,91-94)
class UR5Reach(VecPlayEnv):
    env_id = 'UR5Reach-v0'


class pandaPick(VecPlayEnv):
    env_id = 'pandaPick-v0'


class UR5PlayAbsRPY1Obj(VecPlayEnv):
    env_id = 'UR5PlayAbsRPY1Obj-v0'


# the other UR5 playroom ids: same world, other action decoder (envList.py:101-140, environments.py:915-981)
class UR5Play1Obj(VecPlayEnv):
    env_id = 'UR5Play1Obj-v0'               # absolute_quat, 8-D action


class UR5PlayRel1Obj(VecPlayEnv):
    env_id = 'UR5PlayRel1Obj-v0'            # relative_quat, 8-D action


class UR5PlayRelRPY1Obj(VecPlayEnv):
    env_id = 'UR5PlayRelRPY1Obj-v0'         # relative_rpy


class UR5PlayAbsJoints1Obj(VecPlayEnv):
    env_id = 'UR5PlayAbsJoints1Obj-v0'      # absolute_joints


class UR5PlayRelJoints1Obj(VecPlayEnv):
    env_id = 'UR5PlayRelJoints1Obj-v0'      # relative_joints


# the Panda ids (roboticsPlayroomPybullet/__init__.py:3-63, envList.py:8-86); pandaPlay-v0 / pandaPlayJoints-v0 have TWO
# objects (a second free body and 26-D observations) and are not compiled yet
class pandaReach(VecPlayEnv):
    env_id = 'pandaReach-v0'


class pandaReach2D(VecPlayEnv):
    env_id = 'pandaReach2D-v0'


class pandaPush(VecPlayEnv):
    env_id = 'pandaPush-v0'


class pandaPlayAbsRPY1Obj(VecPlayEnv):
    env_id = 'pandaPlayAbsRPY1Obj-v0'


class pandaPlay1Obj(VecPlayEnv):
    env_id = 'pandaPlay1Obj-v0'             # absolute_quat


class pandaPlayRel1Obj(VecPlayEnv):
    env_id = 'pandaPlayRel1Obj-v0'          # relative_quat


class pandaPlayRelRPY1Obj(VecPlayEnv):
    env_id = 'pandaPlayRelRPY1Obj-v0'


class pandaPlayAbsJoints1Obj(VecPlayEnv):
    env_id = 'pandaPlayAbsJoints1Obj-v0'    # 8-D: 7 joints + gripper


class pandaPlayRelJoints1Obj(VecPlayEnv):
    env_id = 'pandaPlayRelJoints1Obj-v0'


_REGISTRY = {c.env_id: c for c in (UR5Reach, pandaPick, UR5PlayAbsRPY1Obj, UR5Play1Obj, UR5PlayRel1Obj, UR5PlayRelRPY1Obj,
                                   UR5PlayAbsJoints1Obj, UR5PlayRelJoints1Obj, pandaReach, pandaReach2D, pandaPush,
                                   pandaPlayAbsRPY1Obj, pandaPlay1Obj, pandaPlayRel1Obj, pandaPlayRelRPY1Obj,
                                   pandaPlayAbsJoints1Obj, pandaPlayRelJoints1Obj)}


def make(env_id, num_envs=1, **kw):
    """gym.make analogue for the registered ids."""
    if env_id not in _REGISTRY:
        raise NotImplementedError('unknown env id %r (in scope: %s)' % (env_id, sorted(_REGISTRY)))
    return _REGISTRY[env_id](num_envs=num_envs, **kw)
