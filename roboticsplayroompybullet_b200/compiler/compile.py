"""World compiler: reference assets + programmatic scenes -> CompiledModel.

Restates, as data, the world construction of the reference:
  * arm loading and constants       environments.py:356-422
  * default_scene / push_scene      scenes.py:8-43
  * complex_scene (playroom)        scenes.py:46-85 with add_static :88-114,
    add_door :117-182, add_button :184-259, add_drawer :262-340, add_dial :345-426
  * env variants                    envList.py:18-22 (pandaPick), :89-91 (UR5Reach),
                                    :93-99 (UR5PlayAbsRPY1Obj)
`tray/traybox.urdf` (scenes.py:24) ships with pybullet_data, not with the
reference; its five collision boxes are restated from the published file.
"""
import os

import numpy as np

from ..model import CompiledModel, PARAM_NAMES, N_PARAMS
from .geom import Xf, rpy_to_mat, rpy_to_quat, box_inertia
from .urdf import parse_urdf, reduce_arm, URDF_MARGIN
from . import mesh_boxes

STATIC = -1


class World:
    def __init__(self):
        self.cols = []
        self.free = []
        self.slide = []
        self.next_obj = 0

    def new_obj(self):
        self.next_obj += 1
        return self.next_obj - 1

    def add_box(self, body, half, pos, R=None, link=-1, urdf_link=-1, friction=0.5, spin=0.0,
                stiffness=-1.0, damping=-1.0, obj=None):
        # obj: Bullet keeps one <=4-point manifold per pair of collision objects (per child-shape
        # pair for compounds).  Boxes cut out of ONE concave trimesh share an object id; every
        # other collider is its own object.
        if obj is None:
            obj = self.next_obj
            self.next_obj += 1
        self.cols.append(dict(body=body, link=link, urdf_link=urdf_link, obj=obj,
                              R=np.eye(3) if R is None else np.array(R, float),
                              p=np.array(pos, float), half=np.array(half, float),
                              friction=friction, spin=spin, stiffness=stiffness, damping=damping))


def _free_inertia(mass, boxes):
    """Box inertia of the collision AABB (body frame) -- what Bullet computes for
    createMultiBody bodies (inertia always recomputed from the shape)."""
    pts = []
    for c, h in boxes:
        pts += [c - h, c + h]
    pts = np.array(pts)
    return box_inertia(mass, pts.max(0) - pts.min(0))


def build_arm(world, ref_envs, arm_kind):
    if arm_kind == 0:
        links = parse_urdf(os.path.join(ref_envs, 'ur_e_description', 'ur5e2.urdf'))
        base = Xf(rpy_to_mat([0, 0, np.pi / 2]), [0.5, -0.1, 0.0])        # environments.py:367,373
        rest = [-1.50189075, -1.6291067, -1.87020409, -1.21324173, 1.57003561, 0.06970189]  # :371
        ee_urdf, n_ik = 7, 6                                            # :368,372
        site_urdf = [7, 6, 18, 20]                                      # :722-725
    else:
        links = parse_urdf(os.path.join(ref_envs, 'franka_panda', 'panda.urdf'))
        base = Xf(np.eye(3), [-0.5, 0.0, -0.05])                        # environments.py:359,363
        rest = [-0.6, 0.437, 0.217, -2.09, 1.1, 1.4, 1.3, 0.0, 0.0]     # :361 (10th entry never used)
        ee_urdf, n_ik = 11, 7                                           # :360,362
        site_urdf = [11, 10, 9, 10]
    arm, cols, sites = reduce_arm(links)
    nd = arm['nd']
    for c in cols:
        if c['link'] < 0:
            # arm base link: immovable (useFixedBase) but still part of the arm body, so the
            # same-body filter keeps it from colliding with the arm's own links
            world.add_box(0, c['half'], c['p'], c['R'], link=-1, urdf_link=c['urdf_link'],
                          friction=c['friction'])
        else:
            world.add_box(0, c['half'], c['p'], c['R'], link=c['link'], urdf_link=c['urdf_link'],
                          friction=c['friction'], spin=c['spin'], stiffness=c['stiffness'],
                          damping=c['damping'])
    arm_rest = np.zeros(nd)
    arm_rest[:len(rest)] = rest[:nd]
    site_link, site_pos, site_rot = [], [], []
    for u in site_urdf:
        m, X = sites[u]
        site_link.append(m)
        site_pos.append(X.p)
        site_rot.append(X.R)
    urdf_to_dof = {int(u): k for k, u in enumerate(arm['urdf_index'])}
    joints_obs = [urdf_to_dof.get(j, -1) for j in range(8)]            # environments.py:758
    d = dict(nd=nd, n_ik=n_ik, arm_parent=arm['parent'], arm_jtype=arm['jtype'],
             arm_urdf_index=arm['urdf_index'], arm_jpos=arm['jpos'], arm_jrot=arm['jrot'],
             arm_axis=arm['axis'], arm_com=arm['com'], arm_mass=arm['mass'],
             arm_inertia=arm['inertia'], arm_lo=arm['lo'], arm_hi=arm['hi'], arm_jdamp=arm['jdamp'],
             arm_rest=arm_rest, arm_base_pos=base.p, arm_base_rot=base.R,
             site_link=site_link, site_pos=np.array(site_pos), site_rot=np.array(site_rot),
             joints_obs_dof=joints_obs)
    if arm_kind == 0:
        # close_gripper, environments.py:1048-1073:  a -= 0.2; target = scale * a
        g = [(18, 0.055, 100.0, -1), (20, 0.0, 1000.0, 18), (12, 0.5, 100.0, -1), (15, 0.5, 100.0, -1),
             (10, 0.8, 100.0, -1), (13, 0.8, 100.0, -1)]
        d['grip_dof'] = [urdf_to_dof[u] for u, _, _, _ in g]
        d['grip_scale'] = [s for _, s, _, _ in g]
        d['grip_offset'] = [-0.2 * s for _, s, _, _ in g]
        d['grip_force'] = [f for _, _, f, _ in g]
        d['grip_mimic'] = [urdf_to_dof[m] if m >= 0 else -1 for _, _, _, m in g]
        d['grip_obs_dof'] = urdf_to_dof[18]                               # :756
        d['gear_a'], d['gear_b'] = -1, -1
        d['ctrl_ll'] = [-2 * np.pi] * 6                                   # :1019
        d['ctrl_ul'] = [-0.7, 2 * np.pi, -0.5, 2 * np.pi, 2 * np.pi, 2 * np.pi]   # :1020
        d['ctrl_inc'] = [0.1, 0.1, 0.2, 0.2, 0.2, 0.2]                    # :1021
    else:
        # environments.py:1042-1047: target = 0.04 - a/25 on fingers 9, 10
        d['grip_dof'] = [urdf_to_dof[9], urdf_to_dof[10]]
        d['grip_scale'] = [-1.0 / 25, -1.0 / 25]
        d['grip_offset'] = [0.04, 0.04]
        d['grip_force'] = [100.0, 100.0]
        d['grip_mimic'] = [-1, -1]
        d['grip_obs_dof'] = urdf_to_dof[9]                                # :754
        d['gear_a'], d['gear_b'] = urdf_to_dof[9], urdf_to_dof[10]        # :400-405
        d['ctrl_ll'] = [-0.6, -2.2, -3.0, -3.04878596, -np.pi, -np.pi, -np.pi]   # :1015
        d['ctrl_ul'] = [3, 1.8, 0.5, -0.5002492, 3., 3.45266257, 2.40072908]     # :1016
        d['ctrl_inc'] = [0.1, 0.1, 0.2, 0.2, 0.2, 0.2, 0.2]               # :1017
    d['n_grip'] = len(d['grip_dof'])
    d['names'] = arm['names']
    # PyBullet joint index -> joint name for EVERY joint (fixed ones included), for the indexing fixture
    d['urdf_joint_names'] = [L.jname for L in sorted(links, key=lambda L: L.index) if L.index >= 0]
    return d


def default_scene(world, z):
    world.add_box(STATIC, [2, 2, 0.0001], [0, 0, z])                      # scenes.py:12-19 / :49-55


def tray_box(world):
    # pybullet_data/tray/traybox.urdf at [0,0,-0.1] (scenes.py:23-25): floor + four slanted walls
    T = Xf(np.eye(3), [0, 0, -0.1])
    parts = [([.3, .3, .01], [0, 0, 0.005], [0, 0, 0]),
             ([.01, .3, .075], [0.25, 0, 0.059], [0, 0.575469961, 0]),
             ([.01, .3, .075], [-0.25, 0, 0.059], [0, -0.575469961, 0]),
             ([.3, .01, .075], [0, -0.25, 0.059], [0.575469961, 0, 0]),
             ([.3, .01, .075], [0, 0.25, 0.059], [-0.575469961, 0, 0])]
    for half, xyz, rpy in parts:
        X = T * Xf(rpy_to_mat(rpy), xyz)
        world.add_box(STATIC, half, X.p, X.R)


def add_free_box(world, half, mass, friction, pos0, quat0=(0, 0, 0, 1)):
    body = 1 + len(world.free)
    world.add_box(body, half, [0, 0, 0], friction=friction)
    world.free.append(dict(mass=mass, inertia=box_inertia(mass, 2 * np.array(half, float)),
                           lin_damp=0.04, ang_damp=0.04, pos0=np.array(pos0, float),
                           quat0=np.array(quat0, float)))
    return body


def complex_scene(world, ref_envs):
    default_scene(world, -0.27)                                           # scenes.py:49-55
    n_free_total = 2
    slide_body0 = 1 + n_free_total
    # --- block (scenes.py:58-82): half (0.05,0.025,0.025), mass 0.3, lateralFriction 1.5
    add_free_box(world, [0.05, 0.025, 0.025], 0.3, 1.5, [-0.6, -0.06, -0.006])
    # --- drawer (scenes.py:294-333): blockers (static) + free concave body
    world.add_box(STATIC, [0.1, 0.28, 0.005], [-0.13, 0.25, -0.13])
    world.add_box(STATIC, [0.1, 0.05, 0.015], [0, 0.25, -0.06])
    world.add_box(STATIC, [0.03, 0.01, 0.045], [-0.25, -0.02, -0.08])
    world.add_box(STATIC, [0.03, 0.01, 0.045], [-0.0, -0.02, -0.08])
    boxes, _ = mesh_boxes.decompose(os.path.join(ref_envs, 'env_meshes', 'drawer2.obj'), 1.25)
    body = 1 + len(world.free)
    drawer_obj = world.new_obj()
    for c, h in boxes:
        world.add_box(body, h, c, obj=drawer_obj)
    world.free.append(dict(mass=0.1, inertia=_free_inertia(0.1, boxes), lin_damp=0.04, ang_damp=0.04,
                           pos0=np.array([-0.10, -0.00, -0.04]),
                           quat0=rpy_to_quat([np.pi / 2, 0, 0])))
    # --- door (scenes.py:117-182): static 0.1 cube base + prismatic link, concave mesh x0.0015
    world.add_box(STATIC, [0.1, 0.1, 0.1], [0, 0.4, -0.2])
    Rl = rpy_to_mat([0, np.pi / 2, 0])
    boxes, _ = mesh_boxes.decompose(os.path.join(ref_envs, 'env_meshes', 'door.obj'), 0.0015)
    sb = slide_body0 + len(world.slide)
    door_obj = world.new_obj()
    for c, h in boxes:
        world.add_box(sb, h, c, obj=door_obj)
    world.slide.append(dict(jtype=1, pos=np.array([0, 0.4, -0.2]) + np.array([0, 0, 0.27]), R=Rl,
                            axis=[0, 0, 1], mass=0.1, inertia=0.0, ang_damp=0.04,
                            motor=[0.0, 0.0, 1.0, -1.0]))     # default velocity motor (target,kp,kd,maxImp<0 => default)
    # --- button (scenes.py:184-238): prismatic z, box (0.02,0.02,0.005), motor to 0.03 with force 1
    sb = slide_body0 + len(world.slide)
    world.add_box(sb, [0.02, 0.02, 0.005], [0, 0, 0])
    world.slide.append(dict(jtype=1, pos=np.array([0, 0, -0.7]) + np.array([-0.25, 0.45, 0.70]),
                            R=np.eye(3), axis=[0, 0, 1], mass=0.1, inertia=0.0, ang_damp=0.04,
                            motor=[0.03, 0.1, 1.0, 1.0 / 300.0]))
    world.add_box(STATIC, [0.02, 0.02, 0.005], [0, 0, -0.7])              # button base box
    r = 0.03 * (np.pi / 6) ** (1 / 3)                                     # toggle sphere -> equal-volume box
    world.add_box(STATIC, [r, r, r], [-0.25, 0.45, 0.24])                 # scenes.py:250-257
    # --- dial (scenes.py:345-401): revolute z (link frame Rx(pi/2)), box half (0.03,0.01125,0.03)
    sb = slide_body0 + len(world.slide)
    half = np.array([0.0075 * 4, 0.0075 * 1.5, 0.0075 * 4])
    world.add_box(sb, half, [0, 0, 0])
    Izz = box_inertia(0.1, 2 * half)[2]
    world.slide.append(dict(jtype=0, pos=np.array([0.2, 0.0, -0.07]) + np.array([0.0, -0.055, 0.0]),
                            R=rpy_to_mat([np.pi / 2, 0, 0]), axis=[0, 0, 1], mass=0.1, inertia=Izz,
                            ang_damp=0.04, motor=[0.0, 0.0, 1.0, -1.0]))
    world.add_box(STATIC, [0.07, 0.07, 0.01], [0.2, 0.1, -0.03])          # toggle grill scenes.py:417-423
    # --- add_static (scenes.py:88-114)
    world.add_box(STATIC, [0.35, 0.28, 0.005], [0, 0.25, -0.03])
    world.add_box(STATIC, [0.35, 0.01, 0.235], [0., 0.52, -0.00])
    world.add_box(STATIC, [0.37, 0.065, 0.005], [0., 0.45, 0.24])
    world.add_box(STATIC, [0.03, 0.065, 0.235], [-0.34, 0.45, -0.00])
    world.add_box(STATIC, [0.03, 0.065, 0.235], [0.34, 0.45, -0.00])


def make_pairs(world):
    pa, pb = [], []
    n = len(world.cols)
    for i in range(n):
        for j in range(i + 1, n):
            a, b = world.cols[i], world.cols[j]
            if a['body'] == b['body']:
                continue                                   # no self collision (environments.py:327)
            def fixed(c):
                return c['body'] == STATIC or (c['body'] == 0 and c['link'] < 0)
            if fixed(a) and fixed(b):
                continue
            # first collider of the pair is always the movable one of lowest body index
            if fixed(a):
                i2, j2 = j, i
            else:
                i2, j2 = i, j
            pa.append(i2)
            pb.append(j2)
    # pairs of the same object pair must be consecutive (manifold reduction works on runs)
    order = sorted(range(len(pa)), key=lambda k: (world.cols[pa[k]]['obj'], world.cols[pb[k]]['obj'], pa[k], pb[k]))
    return [pa[k] for k in order], [pb[k] for k in order]


def compile_env(env_id, ref_envs):
    world = World()
    p = dict(dt=1.0 / 300, gravity_z=-9.8, erp_joint=0.2, erp_contact=0.08, linear_slop=1e-5,
             ik_damping=0.5, ik_threshold=1e-4, arm_force=240.0, sparse_thresh=0.05,
             reset_z_offset=0.0, default_motor_impulse=1.0, motor_kp=0.1, motor_kd=1.0,
             limit_max_impulse=100.0, gear_ratio=-1.0, gear_erp=0.1, gear_max_impulse=50.0 / 300,
             max_coord_vel=100.0, action_high_xyz=6.0, action_high_grip=1.0, obj_reset_dz=0.03,
             arm_lin_damp=0.0, arm_ang_damp=0.0, contact_breaking=0.02, action_type=0.0)
    # playEnv.__init__ defaults (environments.py:64-67)
    DEF = dict(env_lo=[-0.18, -0.18, -0.05], env_hi=[0.18, 0.18, 0.15], goal_lo=[-0.18, -0.18, -0.05], goal_hi=[0.18, 0.18, 0.05],
               obj_lo=[-0.18, -0.18, -0.05], obj_hi=[-0.18, -0.18, -0.05])
    PLAY = dict(env_hi=[1.0, 1.0, 1.0], goal_lo=[-0.18, 0, 0.05], goal_hi=[0.18, 0.3, 0.1], obj_lo=[-0.18, 0, 0.05], obj_hi=[0.18, 0.3, 0.1])
    TABLE = {
        # env id: arm (0 UR5 + Robotiq, 1 Panda), scene, ranges                                           envList.py
        'UR5Reach-v0': (0, 'default', {}),                                                               # :89-91
        'pandaReach-v0': (1, 'default', {}),                                                             # :8-10
        'pandaReach2D-v0': (1, 'default', dict(env_hi=[0.18, 0.18, 0.0], goal_lo=[-0.18, -0.18, -0.06], goal_hi=[0.18, 0.18, -0.05])),   # :24-26
        'pandaPush-v0': (1, 'push', dict(env_hi=[0.18, 0.18, -0.04], goal_lo=[-0.1, -0.1, -0.06], goal_hi=[0.1, 0.1, -0.05],
                                         obj_lo=[-0.1, -0.1, -0.06], obj_hi=[0.1, 0.1, -0.05])),          # :12-16
        'pandaPick-v0': (1, 'push', dict(env_hi=[0.18, 0.18, 0.2], goal_lo=[-0.18, -0.18, 0.0], goal_hi=[0.18, 0.18, 0.1],
                                         obj_lo=[-0.18, -0.18, 0.0], obj_hi=[0.18, 0.18, 0.1])),          # :18-22
        'UR5PlayAbsRPY1Obj-v0': (0, 'play', PLAY),                                                       # :93-99
        'pandaPlayAbsRPY1Obj-v0': (1, 'play', PLAY),                                                     # :73-79
    }
    if env_id not in TABLE:
        raise NotImplementedError(env_id)
    arm_kind, scene, rng = TABLE[env_id]
    R = dict(DEF)
    R.update(rng)
    d = build_arm(world, ref_envs, arm_kind)
    ik = dict(ik_calls=4, ik_iters=20) if arm_kind == 0 else dict(ik_calls=1, ik_iters=200)   # inverseKinematics.py:44-50 / environments.py:995-997
    if scene == 'default':                                  # scenes.py:8-19: ground only, no object
        default_scene(world, -0.07)
        d.update(env_kind=0, play=0, use_orientation=0, return_velocity=1, obs_dim=7, goal_dim=3, fps_dim=4, observation_dim=6)
    elif scene == 'push':                                   # scenes.py:21-43: ground, tray, one cube
        default_scene(world, -0.07)
        tray_box(world)
        add_free_box(world, [0.025] * 3, 0.1, 0.5, [0, -0.06, -0.06])     # scenes.py:33-38
        d.update(env_kind=1, play=0, use_orientation=0, return_velocity=1, obs_dim=13, goal_dim=3, fps_dim=7, observation_dim=12)
    else:                                                   # scenes.py:46-85: the playroom, one block
        complex_scene(world, ref_envs)
        d.update(env_kind=2, play=1, use_orientation=1, return_velocity=0, obs_dim=19, goal_dim=11, fps_dim=19, observation_dim=18)
    d.update(goal_lo=R['goal_lo'], goal_hi=R['goal_hi'], obj_lo=R['obj_lo'], obj_hi=R['obj_hi'], env_hi=R['env_hi'], **ik)
    if arm_kind == 0:
        p['reset_z_offset'] = 0.2                           # environments.py:580-581 (UR5 only)
    d['arm_kind'] = arm_kind
    d['ik_reset_iters'] = 20
    d['n_substeps'] = 12                                    # environments.py:489
    d['solver_iters'] = 50
    d['settle_steps'] = 100                                 # environments.py:534
    d['default_orn'] = rpy_to_quat([0, 0, 0])               # environments.py:357-358,365-366
    cols = world.cols
    d['n_col'] = len(cols)
    d['col_body'] = [c['body'] for c in cols]
    d['col_link'] = [c['link'] for c in cols]
    d['col_urdf_link'] = [c['urdf_link'] for c in cols]
    d['col_obj'] = [c['obj'] for c in cols]
    d['col_pos'] = np.array([c['p'] for c in cols])
    d['col_rot'] = np.array([c['R'] for c in cols])
    d['col_half'] = np.array([c['half'] for c in cols])
    d['col_friction'] = [c['friction'] for c in cols]
    d['col_spin'] = [c['spin'] for c in cols]
    d['col_stiffness'] = [c['stiffness'] for c in cols]
    d['col_damping'] = [c['damping'] for c in cols]
    d['n_free'] = len(world.free)
    d['free_mass'] = [f['mass'] for f in world.free]
    d['free_inertia'] = np.array([f['inertia'] for f in world.free]).reshape(-1)
    d['free_lin_damp'] = [f['lin_damp'] for f in world.free]
    d['free_ang_damp'] = [f['ang_damp'] for f in world.free]
    d['free_pos0'] = np.array([f['pos0'] for f in world.free]).reshape(-1)
    d['free_quat0'] = np.array([f['quat0'] for f in world.free]).reshape(-1)
    d['n_slide'] = len(world.slide)
    d['slide_jtype'] = [s['jtype'] for s in world.slide]
    d['slide_pos'] = np.array([s['pos'] for s in world.slide]).reshape(-1)
    d['slide_rot'] = np.array([s['R'] for s in world.slide]).reshape(-1)
    d['slide_axis'] = np.array([s['axis'] for s in world.slide], float).reshape(-1)
    d['slide_mass'] = [s['mass'] for s in world.slide]
    d['slide_inertia'] = [s['inertia'] for s in world.slide]
    d['slide_ang_damp'] = [s['ang_damp'] for s in world.slide]
    d['slide_motor'] = np.array([s['motor'] for s in world.slide], float).reshape(-1)
    pa, pb = make_pairs(world)
    d['n_pair'] = len(pa)
    d['pair_a'], d['pair_b'] = pa, pb
    d['params'] = [p[n] for n in PARAM_NAMES]
    assert len(d['params']) == N_PARAMS
    return CompiledModel(d)
