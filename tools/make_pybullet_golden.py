#!/usr/bin/env python
"""Dump golden steps from the REAL reference (PyBullet) -- to be run on a machine that has `pybullet`, `pybullet_data`
and `gym<0.22`; neither this container nor the GPU box does (profiles/r2_pybullet_probe.log), which is why the oracle's
parity is "unpinned" today.  The output pins it:

    python tools/make_pybullet_golden.py /path/to/RoboticsPlayroomPybullet tests/golden/pybullet_steps.npz

For each of the three north-star env ids the script resets the reference env (environments.py:173-187), then repeatedly
  1. records the observation dict `o0` (environments.py:849-861),
  2. lets the scene come to rest with the hold-still action of the current pose (so that velocities, which an observation
     does not carry, are ~0), records `o1`,
  3. applies one seeded action `a` (same distribution as tests/helpers.py random_actions) and records (o2, r, info).
It also records the raw simulator state before the action (every joint position / velocity of the arm body and of the
scene bodies, base poses and twists: `raw_*`), which is what a one-step parity check needs: an observation carries neither
the six Robotiq joints nor velocities.

tests/test_cpu_oracle.py::test_pybullet_fixture (skipped while the file is absent) checks what needs no state mapping:
reward / is_success recomputed from (achieved_goal, desired_goal), the quaternion -> Euler convention between `obs_quat`
and `observation`, the dial read-out.  Mapping `raw_*` (PyBullet joint indices) onto the oracle's state vector and
comparing o2 key by key at 1e-4 m / 1e-3 rad is the step that turns "parity unpinned" into "pinned"; it is not written
because it cannot be exercised here.  THIS SCRIPT HAS NOT BEEN RUN (no PyBullet in reach).
"""
import sys

import numpy as np

KEYS = ['obs_quat', 'achieved_goal', 'desired_goal', 'controllable_achieved_goal', 'full_positional_state', 'joints',
        'velocity', 'observation', 'gripper_proprioception']
ENVS = ['UR5Reach', 'pandaPick', 'UR5PlayAbsRPY1Obj']
N_CASES = 24
HOLD_STEPS = 20


def random_action(rng, play):
    lo, hi = ([-0.30, -0.05, 0.0], [0.30, 0.50, 0.35]) if play else ([-0.18, -0.18, -0.05], [0.18, 0.18, 0.2])
    return np.concatenate([rng.uniform(lo, hi), rng.uniform(-0.5, 0.5, 3), rng.uniform(-1, 1, 1)]).astype(np.float32)


def hold_action(obs, env):
    """The action that asks for the pose the arm already has (absolute xyz + rpy, gripper as it is)."""
    import pybullet as p
    q = obs['obs_quat']
    if env.use_orientation:
        xyz, rpy, grip = q[0:3], p.getEulerFromQuaternion(q[3:7]), q[7]
        rpy = np.asarray(rpy) - np.asarray(env.instance.default_arm_orn_RPY)
    else:
        xyz, rpy, grip = q[0:3], np.zeros(3), q[6]
    return np.concatenate([xyz, rpy, [np.clip(grip, -1, 1)]]).astype(np.float32)


def raw_state(env):
    """Joint states of every body with joints, base pose / twist of every body (PyBullet body ids in creation order)."""
    p = env.p
    out = {}
    for b in range(p.getNumBodies()):
        pos, orn = p.getBasePositionAndOrientation(b)
        lin, ang = p.getBaseVelocity(b)
        out['base%d' % b] = np.asarray(list(pos) + list(orn) + list(lin) + list(ang), np.float64)
        nj = p.getNumJoints(b)
        if nj:
            st = p.getJointStates(b, list(range(nj)))
            out['joints%d' % b] = np.asarray([[s[0], s[1]] for s in st], np.float64)
    return out


def main():
    ref_root, out = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ref_root)
    from roboticsPlayroomPybullet.envs import envList
    rec = {}
    for name in ENVS:
        env = getattr(envList, name)()
        rng = np.random.default_rng(1234)
        play = 'Play' in name
        rows = {k: [] for k in ['o1_' + x for x in KEYS] + ['o2_' + x for x in KEYS] + ['action', 'reward', 'is_success', 'target_poses']}
        obs = env.reset()
        for case in range(N_CASES):
            if case % 8 == 0:
                obs = env.reset()
            for _ in range(HOLD_STEPS):
                obs, _, _, _ = env.step(hold_action(obs, env))
            a = random_action(rng, play)
            raw = raw_state(env)
            o2, r, _, info = env.step(a)
            for k, v in raw.items():
                rows.setdefault('raw_' + k, []).append(v)
            for k in KEYS:
                rows['o1_' + k].append(np.atleast_1d(np.asarray(obs[k], np.float64)))
                rows['o2_' + k].append(np.atleast_1d(np.asarray(o2[k], np.float64)))
            rows['action'].append(a)
            rows['reward'].append(float(r))
            rows['is_success'].append(int(info['is_success']))
            rows['target_poses'].append(np.asarray(info['target_poses'], np.float64))
            obs = o2
        for k, v in rows.items():
            rec[name + '/' + k] = np.asarray(v)
    import pybullet
    rec['pybullet_api_version'] = np.asarray(pybullet.getAPIVersion())
    np.savez_compressed(out, **rec)
    print('wrote', out, 'cases per env', N_CASES)


if __name__ == '__main__':
    main()
