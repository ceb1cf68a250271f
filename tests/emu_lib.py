"""ctypes front-end of tests/emu/libprb_emu.so: the product's kernel source compiled with g++
against the SIMT emulator (tests/emu/cuda_emu.h).  CPU tests only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu')
_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'roboticsplayroompybullet_b200', 'csrc')
_LIB = None

OUT_KEYS = ['obs_quat', 'achieved_goal', 'desired_goal', 'controllable_achieved_goal', 'full_positional_state',
            'joints', 'velocity', 'observation', 'gripper_proprioception', 'reward', 'is_success', 'target_poses']


class DevOut(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in OUT_KEYS] + [('overflow', ctypes.c_void_p), ('ovf_env', ctypes.c_void_p), ('dbg', ctypes.c_void_p)]


def out_dims(m):
    return {'obs_quat': m['obs_dim'], 'achieved_goal': m['goal_dim'], 'desired_goal': m['goal_dim'],
            'controllable_achieved_goal': 4, 'full_positional_state': m['fps_dim'], 'joints': 8, 'velocity': 6,
            'observation': m['observation_dim'], 'gripper_proprioception': 1, 'reward': 1, 'is_success': 1,
            'target_poses': m['n_ik']}


def lib_variant(tag, defines):
    """A second build of the emulated kernels with other compile-time constants (e.g. tiny solver stages)."""
    so = os.path.join(_HERE, 'libprb_emu_%s.so' % tag)
    deps = [os.path.join(_HERE, f) for f in ('prb_emu.cpp', 'cuda_emu.h')] + \
           [os.path.join(_SRC, f) for f in ('prb_kernels.cuh', 'prb_stream.cuh', 'prb_reset.cuh', 'prb_device.h', 'prb_convert.h')]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-U_FORTIFY_SOURCE', '-Wno-unknown-pragmas'] +
                              ['-D%s' % d for d in defines] + ['-o', so, os.path.join(_HERE, 'prb_emu.cpp')])
    L = ctypes.CDLL(so)
    L.emu_error.restype = ctypes.c_char_p
    return L


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libprb_emu.so')
        deps = [os.path.join(_HERE, f) for f in ('prb_emu.cpp', 'cuda_emu.h')] + \
               [os.path.join(_SRC, f) for f in ('prb_kernels.cuh', 'prb_stream.cuh', 'prb_reset.cuh', 'prb_device.h', 'prb_convert.h')]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
            subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-U_FORTIFY_SOURCE',
                                   '-Wno-unknown-pragmas', '-o', so, os.path.join(_HERE, 'prb_emu.cpp')])
        _LIB = ctypes.CDLL(so)
        _LIB.emu_error.restype = ctypes.c_char_p
    return _LIB


class EmuSim:
    def __init__(self, model, N, seed=1234, env_offset=0, library=None):
        self.m = model
        self.L = library if library is not None else lib()
        self.ms = model.as_struct()
        if self.L.emu_set_model(ctypes.byref(self.ms)) != 0:
            raise RuntimeError(self.L.emu_error().decode())
        self.N = N
        self.stride = self.L.emu_state_stride()
        self.state_dim = self.L.emu_state_dim()
        self.state = np.zeros((N, self.stride), np.float32)
        self.seed, self.env_offset = seed, env_offset
        self.out = {k: np.zeros((N, d), np.float32) for k, d in out_dims(model).items()}
        self.O = DevOut(*([self.out[k].ctypes.data for k in OUT_KEYS] + [None, None]))
        self.L.emu_init(self.state.ctypes.data_as(ctypes.c_void_p), N)

    def _sel(self):
        self.L.emu_set_model(ctypes.byref(self.ms))

    def substeps(self, n):
        self._sel()
        self.L.emu_substeps(self.state.ctypes.data_as(ctypes.c_void_p), self.N, n)

    def step(self, action):
        self._sel()
        from roboticsplayroompybullet_b200.model import action_dim
        a = np.ascontiguousarray(action, np.float32).reshape(self.N, action_dim(self.m))
        self.L.emu_step(self.state.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                        ctypes.byref(self.O), self.N)
        return {k: v.copy() for k, v in self.out.items()}

    def observe(self):
        self._sel()
        self.L.emu_observe(self.state.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.O), self.N)
        return {k: v.copy() for k, v in self.out.items()}

    def reset(self, mask=None):
        self._sel()
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mp = mask.ctypes.data_as(ctypes.c_void_p)
        self.L.emu_reset(self.state.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.O), mp, self.N,
                         ctypes.c_uint64(self.seed), ctypes.c_uint32(self.env_offset))
        return {k: v.copy() for k, v in self.out.items()}

    def reset_to(self, obs, mask=None, restore_env=True):
        self._sel()
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mp = mask.ctypes.data_as(ctypes.c_void_p)
        ob = np.ascontiguousarray(obs, np.float32).reshape(self.N, -1)
        self.L.emu_reset_to(self.state.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.O), ob.ctypes.data_as(ctypes.c_void_p), mp,
                            self.N, ctypes.c_uint64(self.seed), ctypes.c_uint32(self.env_offset), 1 if restore_env else 0)
        return {k: v.copy() for k, v in self.out.items()}

    def ik(self, action):
        self._sel()
        from roboticsplayroompybullet_b200.model import action_dim
        a = np.ascontiguousarray(action, np.float32).reshape(self.N, action_dim(self.m))
        t = np.zeros((self.N, self.m['n_ik']), np.float32)
        self.L.emu_ik(self.state.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                      t.ctypes.data_as(ctypes.c_void_p), self.N)
        return t


def box_box(p1, R1, h1, p2, R2, h2):
    f = lambda x: np.ascontiguousarray(x, np.float32)
    out = np.zeros(28, np.float32)
    a = [f(x) for x in (p1, R1, h1, p2, R2, h2)]
    n = lib().emu_box_box(*[x.ctypes.data_as(ctypes.c_void_p) for x in a], out.ctypes.data_as(ctypes.c_void_p))
    return out[:7 * n].reshape(n, 7)
