"""ctypes front-end of the CPU oracle (oracle/prb_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB32 = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libprb_oracle.so')
        src = os.path.join(_HERE, 'prb_oracle.c')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_state_dim.restype = ctypes.c_int
        _LIB.orc_out_dim.restype = ctypes.c_int
        _LIB.orc_box_box.restype = ctypes.c_int
        _LIB.orc_last_contacts.restype = ctypes.c_int
        _LIB.orc_last_rows.restype = ctypes.c_int
    return _LIB


def lib32():
    """The same restatement built with ORC_REAL=float: conditioning probe for the parity tests only."""
    global _LIB32
    if _LIB32 is None:
        so = os.path.join(_HERE, 'libprb_oracle_f32.so')
        src = os.path.join(_HERE, 'prb_oracle.c')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build()
        _LIB32 = ctypes.CDLL(so)
        _LIB32.orc_state_dim.restype = ctypes.c_int
        _LIB32.orc_out_dim.restype = ctypes.c_int
    return _LIB32


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _d(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class Oracle:
    """One environment stepped by the fp64 restatement."""

    def __init__(self, model, seed=1234, env_id=0, f32=False):
        self.model = model
        self.ms = model.as_struct()
        self.mp = ctypes.byref(self.ms)
        self.L = lib32() if f32 else lib()
        self.dtype = np.float32 if f32 else np.float64
        self.state_dim = self.L.orc_state_dim(self.mp)
        self.out_dim = self.L.orc_out_dim(self.mp)
        self.state = np.zeros(self.state_dim, self.dtype)
        self.seed = seed
        self.env_id = env_id
        self.L.orc_init_state(self.mp, _p(self.state))

    def split_out(self, o):
        m = self.model
        dims = [('obs_quat', m['obs_dim']), ('achieved_goal', m['goal_dim']), ('desired_goal', m['goal_dim']),
                ('controllable_achieved_goal', 4), ('full_positional_state', m['fps_dim']), ('joints', 8),
                ('velocity', 6), ('observation', m['observation_dim']), ('gripper_proprioception', 1),
                ('reward', 1), ('is_success', 1), ('target_poses', m['n_ik'])]
        d, k = {}, 0
        for n, s in dims:
            d[n] = o[k:k + s].copy()
            k += s
        return d

    def reset(self):
        out = np.zeros(self.out_dim, self.dtype)
        self.L.orc_reset(self.mp, _p(self.state), ctypes.c_uint64(self.seed), ctypes.c_uint32(self.env_id), _p(out))
        return self.split_out(out)

    def reset_to(self, obs, restore_env=True):
        out = np.zeros(self.out_dim, self.dtype)
        o = np.ascontiguousarray(np.asarray(obs, dtype=self.dtype))
        self.L.orc_reset_to(self.mp, _p(self.state), _p(o), ctypes.c_int(1 if restore_env else 0), ctypes.c_uint64(self.seed),
                            ctypes.c_uint32(self.env_id), _p(out))
        return self.split_out(out)

    def step(self, action):
        out = np.zeros(self.out_dim, self.dtype)
        a = np.ascontiguousarray(np.asarray(action, dtype=self.dtype))
        self.L.orc_step(self.mp, _p(self.state), _p(a), _p(out))
        return self.split_out(out)

    def calc_state(self):
        out = np.zeros(self.out_dim, self.dtype)
        self.L.orc_calc_state(self.mp, _p(self.state), _p(out))
        return self.split_out(out)

    def substeps(self, n):
        self.L.orc_substeps(self.mp, _p(self.state), ctypes.c_int(n))

    def last_rows(self):
        """Rows of the most recent substep: J [R, nv], B = M^-1 J^T [R, nv], scalars [R, 8] =
        {rhs, cfm, invD, lo, hi, lambda, mu, normal_row}, number of contacts."""
        self.L.orc_nv.restype = ctypes.c_int
        nv, R = self.L.orc_nv(self.mp), self.L.orc_last_rows()
        J, B, sc = np.zeros((R, nv), self.dtype), np.zeros((R, nv), self.dtype), np.zeros((R, 8), self.dtype)
        for i in range(R):
            self.L.orc_last_row(self.mp, ctypes.c_int(i), _p(J[i]), _p(B[i]), _p(sc[i]))
        return J, B, sc, self.L.orc_last_contacts()

    def ik(self, q, pos, quat, iters):
        q = _d(q); pos = _d(pos); quat = _d(quat)
        out = np.zeros(self.model['nd'])
        self.L.orc_ik(self.mp, _p(q), _p(pos), _p(quat), ctypes.c_int(iters), _p(out))
        return out

    def calc_angles(self, q, pos, quat):
        q = _d(q); pos = _d(pos); quat = _d(quat)
        out = np.zeros(self.model['nd'])
        self.L.orc_calc_angles(self.mp, _p(q), _p(pos), _p(quat), _p(out))
        return out

    def fk_sites(self, q):
        q = _d(q)
        out = np.zeros(28)
        self.L.orc_fk_sites(self.mp, _p(q), _p(out))
        return out.reshape(4, 7)

    def minv(self, q):
        q = _d(q)
        nd = self.model['nd']
        out = np.zeros((nd, nd))
        self.L.orc_arm_minv_matrix(self.mp, _p(q), _p(out))
        return out

    def qdd(self, q, qd):
        q = _d(q); qd = _d(qd)
        out = np.zeros(self.model['nd'])
        self.L.orc_arm_qdd(self.mp, _p(q), _p(qd), _p(out))
        return out

    def compute_reward(self, ag, dg):
        ag = _d(ag).reshape(-1, self.model['goal_dim']); dg = _d(dg).reshape(-1, self.model['goal_dim'])
        out = np.zeros(len(ag))
        self.L.orc_compute_reward(self.mp, _p(ag), _p(dg), ctypes.c_int64(len(ag)), _p(out))
        return out


def box_box(p1, R1, h1, p2, R2, h2):
    out = np.zeros(56)
    n = lib().orc_box_box(_p(_d(p1)), _p(_d(R1)), _p(_d(h1)), _p(_d(p2)), _p(_d(R2)), _p(_d(h2)), _p(out))
    return out[:7 * n].reshape(n, 7)


def quat_from_euler(rpy):
    q = np.zeros(4)
    lib().orc_quat_from_euler(_p(_d(rpy)), _p(q))
    return q


def euler_from_quat(q):
    e = np.zeros(3)
    lib().orc_euler_from_quat(_p(_d(q)), _p(e))
    return e


def rng4(seed, env, attempt, block):
    u = np.zeros(4)
    lib().orc_rng4(ctypes.c_uint64(seed), ctypes.c_uint32(env), ctypes.c_uint32(attempt), ctypes.c_uint32(block), _p(u))
    return u
