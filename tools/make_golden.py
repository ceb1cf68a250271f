"""Generate the committed fixtures under tests/golden/.

1. notebook_ur5.json -- the ONLY numeric outputs the reference records anywhere
   (envs/ur_e_description/testing_bullet_ik.ipynb): joint count (cell 1), the 22-row
   index -> joint-name table (cell 2) and one FK read-out at `default_joints` (cells 12, 18).
   They were produced by PyBullet itself, so they pin the URDF indexing / frame conventions of the
   model compiler and the oracle's forward kinematics.  (The recorded z is stale by exactly the
   shoulder-height change 0.163 -> 0.083 between the URDF the notebook used and the shipped
   ur5e2.urdf: SURVEY.md section 4; x, y and the orientation are valid.)
2. oracle_regression.npz -- seeded runs of the CPU oracle (oracle/prb_oracle.c).  These are NOT
   reference outputs (PyBullet is not installable here: parity unpinned); they freeze the oracle so
   an accidental change to the checker is caught.

Run where /root/reference exists:  python tools/make_golden.py
"""
import ast
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
REF = '/root/reference/roboticsPlayroomPybullet/envs'


def notebook_fixture():
    nb = json.load(open(os.path.join(REF, 'ur_e_description', 'testing_bullet_ik.ipynb')))
    cells = nb['cells']

    def out(i):
        return ''.join(''.join(o.get('text', o.get('data', {}).get('text/plain', ''))) for o in cells[i]['outputs'])
    n_joints = int(out(1).strip())
    names = {}
    for line in out(2).strip().splitlines():
        i, n = line.split(' ', 1)
        names[int(i)] = ast.literal_eval(n).decode()
    src12 = ''.join(cells[12]['source'])
    dj = [float(x) for x in re.findall(r'-?\d+\.\d+', src12)][:6]
    o18 = out(18).strip().splitlines()
    euler = list(ast.literal_eval(o18[0]))
    pos = list(ast.literal_eval(o18[1]))
    return {'source': 'envs/ur_e_description/testing_bullet_ik.ipynb cells 1, 2, 12, 18 (PyBullet outputs)',
            'n_joints': n_joints, 'joint_names': [names[i] for i in range(n_joints)],
            'default_joints': dj, 'ee_index': 6, 'ee_euler': euler, 'ee_pos_recorded': pos,
            'ee_pos_stale_dz': 0.163 - 0.083}


def oracle_fixture():
    from roboticsplayroompybullet_b200.model import load_model
    from oracle.oracle import Oracle
    out = {}
    for env_id in ['UR5Reach-v0', 'pandaPick-v0', 'UR5PlayAbsRPY1Obj-v0']:
        m = load_model(env_id)
        o = Oracle(m, seed=77, env_id=3)
        d = o.reset()
        rng = np.random.default_rng(5)
        acts = np.concatenate([rng.uniform(-0.15, 0.25, (6, 3)), rng.uniform(-0.3, 0.3, (6, 3)), rng.uniform(-1, 1, (6, 1))], 1)
        tag = env_id.replace('-', '_')
        out[tag + '__reset_state'] = o.state.copy()
        out[tag + '__reset_obs_quat'] = d['obs_quat']
        out[tag + '__reset_goal'] = d['desired_goal']
        out[tag + '__actions'] = acts
        for a in acts:
            d = o.step(a)
        out[tag + '__final_state'] = o.state.copy()
        out[tag + '__final_obs_quat'] = d['obs_quat']
        out[tag + '__final_target_poses'] = d['target_poses']
        out[tag + '__final_reward'] = d['reward']
    return out


if __name__ == '__main__':
    gd = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gd, exist_ok=True)
    json.dump(notebook_fixture(), open(os.path.join(gd, 'notebook_ur5.json'), 'w'), indent=1)
    np.savez(os.path.join(gd, 'oracle_regression.npz'), **oracle_fixture())
    print('wrote', os.listdir(gd))
