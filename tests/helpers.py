"""Shared helpers for the parity tests: seeded synthetic actions and comparison
of a batched result against per-env oracle runs started from identical states."""
import numpy as np

OBS_KEYS = ['obs_quat', 'achieved_goal', 'desired_goal', 'controllable_achieved_goal', 'full_positional_state',
            'joints', 'velocity', 'observation']

# tolerances (BASELINE.json north_star): poses 1e-4 m / 1e-3 rad after one env step from an
# identical state; velocities are compared relative to their magnitude.
POS_TOL = 1e-4
VEL_TOL = 5e-3


def random_actions(rng, n, env_id):
    if env_id == 'UR5PlayAbsRPY1Obj-v0':
        lo, hi = [-0.30, -0.05, 0.0], [0.30, 0.50, 0.35]
    else:
        lo, hi = [-0.18, -0.18, -0.05], [0.18, 0.18, 0.2]
    a = np.concatenate([rng.uniform(lo, hi, (n, 3)), rng.uniform(-0.5, 0.5, (n, 3)), rng.uniform(-1, 1, (n, 1))], 1)
    tail = rng.random(n) < 0.05                      # 5% from the full clip box: exercises the +-inc clamp
    a[tail, :6] = rng.uniform(-6, 6, (int(tail.sum()), 6))
    return a.astype(np.float32)


def oracle_step_from(model, state_row, action, Oracle):
    o = Oracle(model)
    o.state[:] = state_row.astype(np.float64)
    d = o.step(action.astype(np.float64))
    return d, o.state.copy()


def compare_step(obs, r, info, oracle_outs, max_outlier_frac=0.05):
    """Returns (n_bad_envs, worst) over position-like keys."""
    n = len(oracle_outs)
    bad, worst = 0, 0.0
    for i in range(n):
        d = oracle_outs[i]
        env_bad = False
        for k in ['obs_quat', 'achieved_goal', 'controllable_achieved_goal', 'full_positional_state', 'joints']:
            a = np.asarray(obs[k][i], np.float64)
            b = np.asarray(d[k], np.float64)
            if k == 'obs_quat' and len(b) in (7, 13):      # velocity entries of the non-play layouts
                mask = np.ones(len(b), bool)
                mask[3:6] = False
                if len(b) == 13:
                    mask[10:13] = False
                a, b = a[mask], b[mask]
            e = float(np.abs(a - b).max())
            worst = max(worst, e)
            if e > POS_TOL:
                env_bad = True
        bad += env_bad
    return bad, worst
