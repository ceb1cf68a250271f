"""Diagnostic: one-step parity vs the oracle in the scripted steady state, split pipeline vs the fused
A/B path (PRB_PIPELINE=fused) from the SAME states."""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from helpers import compare_step, oracle_step_from, POS_TOL
from roboticsplayroompybullet_b200.envs import make
from roboticsplayroompybullet_b200.model import load_model
from oracle.oracle import Oracle
env_id = 'UR5PlayAbsRPY1Obj-v0'
n = 64
env = make(env_id, num_envs=n, seed=13)
obs = env.reset()
acts = bench.synth_actions(np.random.default_rng(2), n, 64, env_id, block_xyz=obs['achieved_goal'][:, :3], ee_xyz=obs['obs_quat'][:, :3])
for s in range(60): env.step(acts[s])
os.environ['PRB_PIPELINE'] = 'fused'
envf = make(env_id, num_envs=n, seed=13)
m = load_model(env_id)
KEYS = ['obs_quat', 'achieved_goal', 'controllable_achieved_goal', 'full_positional_state', 'joints']
for s in range(60, 63):
    st = env.get_state()
    envf.set_state(st)
    obs, r, _, info = env.step(acts[s])
    obf, rf, _, inff = envf.step(acts[s])
    us = env.debug_usage()
    outs = [oracle_step_from(m, st[i], acts[s][i], Oracle) for i in range(n)]
    for i in range(n):
        e = max(float(np.abs(np.asarray(obs[k][i], np.float64) - outs[i][0][k]).max()) for k in KEYS)
        ef = max(float(np.abs(np.asarray(obf[k][i], np.float64) - outs[i][0][k]).max()) for k in KEYS)
        if e > POS_TOL or ef > POS_TOL:
            print('step %d env %2d  split err %.2e  fused err %.2e  contacts %d stream q %d' % (s, i, e, ef, us[i, 1], us[i, 2]))
