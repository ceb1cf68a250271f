"""Compiled simulation model: the struct-of-arrays description handed to the C-ABI.

The reference builds its world imperatively (loadURDF, createMultiBody, ...:
environments.py:321-454, scenes.py:8-426).  Here the same world is compiled ONCE
on the host into flat int32/float64 arrays (`CompiledModel`), stored as .npz under
`assets/`, and passed by pointer to `prb_create` (include/prb.h), which converts
it to the fp32 device layout.  The schema below is the single source of truth:
`include/prb_model.h` is generated from it (tools/gen_header.py) and a CPU test
checks they stay in sync.
"""
import ctypes
import os

import numpy as np

# (name, kind)   kind: 'i' int32 scalar | 'I' int32 array | 'R' float64 array
SCHEMA = [
    # ---- scalars
    ('env_kind', 'i'),          # 0 reach (no object), 1 push / pick (one object on the tray), 2 playroom
    ('arm_kind', 'i'),          # 0 UR5, 1 Panda
    ('nd', 'i'),                # arm DoF (movable links after folding fixed joints)
    ('n_ik', 'i'),              # leading arm DoF driven by IK / position control (6 or 7)
    ('n_free', 'i'),            # free 6-DoF bodies (block, drawer)
    ('n_slide', 'i'),           # single-joint fixed-base bodies (door, button, dial)
    ('n_col', 'i'),             # box colliders
    ('n_pair', 'i'),            # candidate collider pairs
    ('n_grip', 'i'),            # gripper motor entries
    ('ik_calls', 'i'),          # chained calculateInverseKinematics calls per step (4 UR5, 1 Panda)
    ('ik_iters', 'i'),          # max DLS iterations per call (20 UR5, 200 Panda)
    ('ik_reset_iters', 'i'),    # iterations of the single reset-time IK call (20)
    ('n_substeps', 'i'),        # physics substeps per env step (12)
    ('solver_iters', 'i'),      # PGS iterations (50)
    ('settle_steps', 'i'),      # physics substeps after re-seating objects in reset (100)
    ('obs_dim', 'i'),           # len(obs_quat)
    ('goal_dim', 'i'),          # len(achieved_goal)
    ('fps_dim', 'i'),           # len(full_positional_state)
    ('observation_dim', 'i'),   # len(observation)  (6/12/18, reference quirk kept)
    ('use_orientation', 'i'),
    ('return_velocity', 'i'),
    ('play', 'i'),
    ('grip_obs_dof', 'i'),      # arm DoF whose position is the gripper observation
    ('gear_a', 'i'),            # gear constraint DoFs (Panda fingers) or -1
    ('gear_b', 'i'),
    # ---- int arrays
    ('arm_parent', 'I'), ('arm_jtype', 'I'), ('arm_urdf_index', 'I'),
    ('joints_obs_dof', 'I'),    # [8] DoF index per PyBullet joint 0..7 (-1 = fixed joint -> 0.0)
    ('site_link', 'I'),         # [4] movable link of sites: EE, wrist(EE-1), pad A (18), pad B (20)
    ('col_body', 'I'),          # -1 static, 0 arm, 1..n_free free bodies, then slide bodies
    ('col_link', 'I'),          # arm movable link (or -1: arm base, static)
    ('col_urdf_link', 'I'),     # PyBullet link index (ray-test classification)
    ('col_obj', 'I'),           # collision-object id: contacts are reduced to <= 4 per object pair
    ('pair_a', 'I'), ('pair_b', 'I'),
    ('slide_jtype', 'I'),
    ('grip_dof', 'I'), ('grip_mimic', 'I'),
    # ---- real arrays
    ('arm_jpos', 'R'), ('arm_jrot', 'R'), ('arm_axis', 'R'), ('arm_com', 'R'), ('arm_mass', 'R'),
    ('arm_inertia', 'R'), ('arm_lo', 'R'), ('arm_hi', 'R'), ('arm_jdamp', 'R'), ('arm_rest', 'R'),
    ('arm_base_pos', 'R'), ('arm_base_rot', 'R'),
    ('site_pos', 'R'), ('site_rot', 'R'),
    ('col_pos', 'R'), ('col_rot', 'R'), ('col_half', 'R'), ('col_friction', 'R'), ('col_spin', 'R'),
    ('col_stiffness', 'R'), ('col_damping', 'R'),
    ('free_mass', 'R'), ('free_inertia', 'R'), ('free_lin_damp', 'R'), ('free_ang_damp', 'R'),
    ('free_pos0', 'R'), ('free_quat0', 'R'),
    ('slide_pos', 'R'), ('slide_rot', 'R'), ('slide_axis', 'R'), ('slide_mass', 'R'),
    ('slide_inertia', 'R'), ('slide_ang_damp', 'R'), ('slide_motor', 'R'),
    ('ctrl_ll', 'R'), ('ctrl_ul', 'R'), ('ctrl_inc', 'R'),
    ('grip_scale', 'R'), ('grip_offset', 'R'), ('grip_force', 'R'),
    ('goal_lo', 'R'), ('goal_hi', 'R'), ('obj_lo', 'R'), ('obj_hi', 'R'), ('env_hi', 'R'),
    ('default_orn', 'R'),       # [4] default end-effector orientation quaternion (reset IK target)
    ('params', 'R'),            # see PARAM_NAMES
]

PARAM_NAMES = [
    'dt', 'gravity_z', 'erp_joint', 'erp_contact', 'linear_slop', 'ik_damping', 'ik_threshold',
    'arm_force', 'sparse_thresh', 'reset_z_offset', 'default_motor_impulse', 'motor_kp', 'motor_kd',
    'limit_max_impulse', 'gear_ratio', 'gear_erp', 'gear_max_impulse', 'max_coord_vel',
    'action_high_xyz', 'action_high_grip', 'obj_reset_dz', 'arm_lin_damp', 'arm_ang_damp',
    'contact_breaking', 'action_type',
]
N_PARAMS = len(PARAM_NAMES)

# compiled worlds (one .npz each); the other registered ids are action-decoder variants of these (ACTION_VARIANTS)
ENV_KINDS = {'UR5Reach-v0': 0, 'pandaPick-v0': 1, 'UR5PlayAbsRPY1Obj-v0': 2,
             'pandaReach-v0': 0, 'pandaReach2D-v0': 0, 'pandaPush-v0': 1, 'pandaPlayAbsRPY1Obj-v0': 2}


class PrbModelStruct(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32 if k == 'i' else
                 (ctypes.POINTER(ctypes.c_int32) if k == 'I' else ctypes.POINTER(ctypes.c_double)))
                for n, k in SCHEMA]


class CompiledModel:
    def __init__(self, d):
        self.d = {}
        for n, k in SCHEMA:
            v = d[n]
            if k == 'i':
                self.d[n] = int(v)
            elif k == 'I':
                self.d[n] = np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(-1))
            else:
                self.d[n] = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
        self.meta = {k: d[k] for k in d if k not in self.d}

    def __getitem__(self, k):
        return self.d[k]

    def param(self, name):
        return float(self.d['params'][PARAM_NAMES.index(name)])

    def save(self, path):
        np.savez(path, **{k: np.asarray(v) for k, v in self.d.items()})
        import json
        meta = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in self.meta.items()
                if isinstance(v, (list, tuple, str, int, float))}
        json.dump(meta, open(os.path.splitext(path)[0] + '.meta.json', 'w'), indent=1)

    @staticmethod
    def load(path):
        z = np.load(path)
        d = {k: z[k] for k in z.files}
        mp = os.path.splitext(path)[0] + '.meta.json'
        if os.path.exists(mp):
            import json
            d.update(json.load(open(mp)))
        return CompiledModel(d)

    def as_struct(self):
        """ctypes view; keeps references to the arrays alive on the returned object."""
        s = PrbModelStruct()
        keep = []
        for n, k in SCHEMA:
            v = self.d[n]
            if k == 'i':
                setattr(s, n, v)
            elif k == 'I':
                a = v if v.size else np.zeros(1, np.int32)
                keep.append(a)
                setattr(s, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
            else:
                a = v if v.size else np.zeros(1, np.float64)
                keep.append(a)
                setattr(s, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        s._keep = keep
        return s


def asset_path(env_id):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets', env_id + '.npz')


# action decoders of the reference (environments.py:915-981); value of the `action_type` parameter
ACTION_TYPES = {'absolute_rpy': 0, 'relative_rpy': 1, 'absolute_quat': 2, 'relative_quat': 3,
                'absolute_joints': 4, 'relative_joints': 5}
# env ids that differ from a compiled asset only by the action decoder (envList.py:101-140): base asset,
# action type, bound of the non-gripper action entries (environments.py:88-112)
ACTION_VARIANTS = {
    'UR5Play1Obj-v0': ('UR5PlayAbsRPY1Obj-v0', 'absolute_quat', 1.0),
    'UR5PlayRel1Obj-v0': ('UR5PlayAbsRPY1Obj-v0', 'relative_quat', 1.0),
    'UR5PlayRelRPY1Obj-v0': ('UR5PlayAbsRPY1Obj-v0', 'relative_rpy', 1.0),
    'UR5PlayAbsJoints1Obj-v0': ('UR5PlayAbsRPY1Obj-v0', 'absolute_joints', 6.0),
    'UR5PlayRelJoints1Obj-v0': ('UR5PlayAbsRPY1Obj-v0', 'relative_joints', 1.0),
    # the Panda in the playroom, one object (envList.py:45-86)
    'pandaPlay1Obj-v0': ('pandaPlayAbsRPY1Obj-v0', 'absolute_quat', 1.0),
    'pandaPlayRel1Obj-v0': ('pandaPlayAbsRPY1Obj-v0', 'relative_quat', 1.0),
    'pandaPlayRelRPY1Obj-v0': ('pandaPlayAbsRPY1Obj-v0', 'relative_rpy', 1.0),
    'pandaPlayAbsJoints1Obj-v0': ('pandaPlayAbsRPY1Obj-v0', 'absolute_joints', 6.0),
    'pandaPlayRelJoints1Obj-v0': ('pandaPlayAbsRPY1Obj-v0', 'relative_joints', 1.0),
}


def action_dim(model):
    """Length of one action (environments.py:88-112): 8 for the quaternion decoders (xyz + quat + gripper), one entry per IK
    joint + gripper for the joint decoders (7 UR5, 8 Panda), else 7 (xyz + rpy + gripper)."""
    t = int(round(model.param('action_type')))
    return 8 if t in (2, 3) else (int(model['n_ik']) + 1 if t >= 4 else 7)


def load_model(env_id):
    if env_id in ACTION_VARIANTS:
        base, atype, high = ACTION_VARIANTS[env_id]
        m = load_model(base)
        prm = m.d['params'].copy()
        prm[PARAM_NAMES.index('action_type')] = float(ACTION_TYPES[atype])
        prm[PARAM_NAMES.index('action_high_xyz')] = high
        m.d['params'] = prm
        return m
    p = asset_path(env_id)
    if not os.path.exists(p):
        raise FileNotFoundError(
            'compiled model %s missing; run tools/compile_models.py with the reference assets' % p)
    return CompiledModel.load(p)


def c_header():
    """Text of include/prb_model.h generated from SCHEMA."""
    out = ['/* GENERATED by tools/gen_header.py from roboticsplayroompybullet_b200/model.py: SCHEMA.',
           ' * Compiled world description consumed by prb_create (include/prb.h).  It replaces the',
           ' * imperative world construction of the reference (environments.py:321-454,',
           ' * scenes.py:8-426, ur5e2.urdf, panda.urdf).  All reals are float64 on the host side;',
           ' * the library converts to the fp32 device layout. */',
           '#ifndef PRB_MODEL_H', '#define PRB_MODEL_H', '#include <stdint.h>', '',
           '#define PRB_N_PARAMS %d' % N_PARAMS, 'enum prb_param {']
    for i, n in enumerate(PARAM_NAMES):
        out.append('  PRB_P_%s = %d,' % (n.upper(), i))
    out += ['};', '', 'typedef struct prb_model {']
    for n, k in SCHEMA:
        t = {'i': 'int32_t ', 'I': 'const int32_t* ', 'R': 'const double* '}[k]
        out.append('  %s%s;' % (t, n))
    out += ['} prb_model;', '', '#endif', '']
    return '\n'.join(out)
