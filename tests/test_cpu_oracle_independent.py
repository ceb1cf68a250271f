"""Oracle-free checks of the oracle (and of the emulated kernels where cheap).

PyBullet is absent from this image AND from the GPU box (profiles/r2_pybullet_probe.log), so the fp64 restatement in
oracle/ cannot be pinned to the reference's own arithmetic.  These tests harden it from the other side: every check
below compares it with something that shares NO code with it:

  * the constraint solve against an exact active-set solution of the same bounded QP (scipy BVLS on an
    eigen-factorisation of the Delassus matrix) — the projected Gauss-Seidel row updates, bounds, CFM terms and the
    M^-1 J^T products are all exercised, only the friction cone is outside a QP and is checked physically instead;
  * Coulomb friction against closed-form sliding / sticking of a box on the table;
  * free flight against the closed-form trajectory of semi-implicit Euler, and momentum / energy drift of a spinning body;
  * box-box penetration depth, normal and contact points against a direction-sampling estimate of the minimum
    translation distance built on support functions (no separating-axis code).
"""
import copy

import numpy as np
import pytest

from roboticsplayroompybullet_b200.model import CompiledModel, load_model
from oracle.oracle import Oracle, box_box as orc_box_box


def _variant(env_id, **over):
    m = load_model(env_id)
    d = dict(m.d)
    d.update(m.meta)
    for k, v in over.items():
        d[k] = v
    return CompiledModel(d)


def _frictionless(env_id, iters):
    m = load_model(env_id)
    return _variant(env_id, col_friction=np.zeros_like(m['col_friction']), col_spin=np.zeros_like(m['col_spin']), solver_iters=iters)


def _kkt_residual(H, b, lo, hi, lam):
    g = H @ lam - b
    return np.abs(np.where(lam <= lo + 1e-13, np.minimum(g, 0), np.where(lam >= hi - 1e-13, np.maximum(g, 0), g))).max()


def _bounded_qp(J, B, sc):
    """min 1/2 l^T (A + C) l - b^T l, lo <= l <= hi  with A = J M^-1 J^T (PSD, singular when contact points are
    redundant), solved WITHOUT Gauss-Seidel: L-BFGS-B finds the active set, then the free rows are solved exactly by
    least squares and the KKT conditions of the result are verified.  Rows pinned by lo = hi (the friction rows of the
    frictionless variant) are taken out."""
    from scipy.optimize import minimize
    rhs, cfm, invD, lo, hi = sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3].copy(), sc[:, 4].copy()
    free = (invD > 0) & (hi > lo)
    lam = np.where(free, 0.0, lo)
    idx = np.nonzero(free)[0]
    A = J @ B.T
    Af = A[np.ix_(idx, idx)]
    C = np.diag(cfm[idx] / invD[idx])
    b = rhs[idx] / invD[idx] - A[np.ix_(idx, np.nonzero(~free)[0])] @ lam[~free]
    H = 0.5 * ((Af + C) + (Af + C).T)
    lo_f, hi_f = lo[idx], np.where(hi[idx] > 1e9, np.inf, hi[idx])
    res = minimize(lambda x: 0.5 * x @ H @ x - b @ x, np.clip(np.zeros(len(idx)), lo_f, hi_f), jac=lambda x: H @ x - b,
                   method='L-BFGS-B', bounds=list(zip(lo_f, hi_f)), options={'maxcor': 60, 'ftol': 1e-30, 'gtol': 1e-13, 'maxiter': 200000, 'maxfun': 400000})
    x = res.x
    for _ in range(20):                                     # polish: exact solve on the free set, re-clip, repeat
        at_lo, at_hi = x <= lo_f + 1e-12, x >= hi_f - 1e-12
        fr = ~(at_lo | at_hi)
        if fr.any():
            xa = np.where(at_lo, lo_f, np.where(at_hi, hi_f, 0.0))
            sol = np.linalg.lstsq(H[np.ix_(fr, fr)], b[fr] - H[np.ix_(fr, ~fr)] @ xa[~fr], rcond=1e-14)[0]
            xn = xa.copy()
            xn[fr] = sol
            xn = np.clip(xn, lo_f, hi_f)
            if _kkt_residual(H, b, lo_f, hi_f, xn) <= _kkt_residual(H, b, lo_f, hi_f, x):
                x = xn
    assert _kkt_residual(H, b, lo_f, hi_f, x) < 1e-8, _kkt_residual(H, b, lo_f, hi_f, x)
    lam[idx] = x
    obj = lambda l: 0.5 * l[idx] @ H @ l[idx] - b @ l[idx]
    kkt = lambda l: _kkt_residual(H, b, lo_f, hi_f, l[idx])
    return lam, A, obj, kkt


@pytest.mark.parametrize('env_id', ['pandaPick-v0', 'UR5PlayAbsRPY1Obj-v0'])
def test_pgs_converges_to_the_exact_bounded_qp(env_id):
    """Frictionless variant of the model (mu = 0 turns the friction rows into lambda = 0): motors / limits / gear with box
    bounds plus contact normals with lambda >= 0 and soft-contact CFM are exactly a bounded QP.  The oracle's PGS run to
    convergence must reach the QP optimum found by an active-set method, in velocity space (the impulses of redundant
    contact points are not unique, their velocity change is)."""
    m = _frictionless(env_id, 50)
    o = Oracle(m, seed=5)
    o.reset()
    rng = np.random.default_rng(0)
    nd = m['nd']
    # a state with work for every row kind: arm commanded away from where it is (motors saturate), block slightly
    # pressed into its support (normals with position correction), velocities everywhere
    o.state[nd:2 * nd] = rng.uniform(-0.5, 0.5, nd)
    o.step(np.array([0.05, 0.1, 0.1, 0.2, -0.1, 0.3, 1.0]))
    o.state[5 * nd + 2] -= 0.002
    o.state[5 * nd + 7:5 * nd + 13] = rng.uniform(-0.2, 0.2, 6)
    s0 = o.state.copy()
    # a finger at its joint limit pushed on by its motor makes Gauss-Seidel creep (two rows on one DoF): run it long
    oc = Oracle(_frictionless(env_id, 400000))
    oc.state[:] = s0
    oc.substeps(1)
    J, B, sc, nc = oc.last_rows()
    assert nc >= 4 and len(J) >= nd + 4
    lam_qp, A, obj, kkt = _bounded_qp(J, B, sc)
    lam_pgs = sc[:, 5]
    # (1) the converged Gauss-Seidel impulses satisfy the KKT conditions of the QP, evaluated here from the exported rows
    assert kkt(lam_pgs) < 1e-9, kkt(lam_pgs)
    # (2) and reach the optimum the quasi-Newton / active-set solver found (A is singular for redundant contact points:
    #     the impulses are not unique, the objective and the velocity change are)
    assert abs(obj(lam_pgs) - obj(lam_qp)) < 1e-9 * max(1.0, abs(obj(lam_qp))), (obj(lam_pgs), obj(lam_qp))
    dv_qp, dv_pgs = B.T @ lam_qp, B.T @ lam_pgs
    scale = max(1e-9, np.abs(dv_qp).max())
    assert np.abs(dv_qp - dv_pgs).max() < 1e-8 * max(1.0, scale), np.abs(dv_qp - dv_pgs).max()
    # the bounded rows that saturate are the same
    act_qp = (lam_qp >= sc[:, 4] - 1e-9) | (lam_qp <= sc[:, 3] + 1e-9)
    act_pgs = (lam_pgs >= sc[:, 4] - 1e-9) | (lam_pgs <= sc[:, 3] + 1e-9)
    motors = np.arange(len(J)) < len(J) - 4 * nc
    assert (act_qp[motors] == act_pgs[motors]).mean() > 0.9
    # the 50 iterations Bullet runs are a truncation of that iteration (solverResidualThreshold = 0, environments.py:326):
    # close to, not at, the optimum
    o.substeps(1)
    _, B50, sc50, _ = o.last_rows()
    dv50 = B50.T @ sc50[:, 5]
    assert 1e-9 < np.abs(dv50 - dv_qp).max() < 0.2 * max(1.0, scale)


def _block_on_table(m, vx=0.0, vy=0.0):
    o = Oracle(m, seed=2)
    o.reset()
    nd = m['nd']
    f0 = 5 * nd
    # park the block on the table top, at rest, far from the arm and the furniture
    o.state[f0:f0 + 3] = [0.0, 0.18, o.state[f0 + 2]]
    o.state[f0 + 3:f0 + 7] = [0, 0, 0, 1]
    o.state[f0 + 7:f0 + 13] = 0
    o.substeps(60)                                   # settle
    o.state[f0 + 7] = vx
    o.state[f0 + 8] = vy
    return o, f0


def test_coulomb_friction_sliding_and_sticking():
    """A box sliding on the table decelerates at mu g (mu = product of the two friction coefficients, clamped as in
    Bullet) — also when it slides along a direction that is not a friction-row axis (implicit cone; load transfer between
    the corner contacts may then turn the box, so only the magnitude is checked) — and a
    box pushed below the static limit does not move."""
    m = load_model('UR5PlayAbsRPY1Obj-v0')
    dt, g = m.param('dt'), -m.param('gravity_z')
    for vx, vy in [(0.4, 0.0), (0.0, -0.4), (0.3, 0.3)]:
        o, f0 = _block_on_table(m, vx, vy)
        v0 = np.array([vx, vy])
        o.substeps(1)
        J, B, sc, nc = o.last_rows()
        mu = sc[len(J) - 2 * nc, 6]                          # first friction row of the step
        assert mu > 0
        n = 10
        o.substeps(n - 1)
        v = o.state[f0 + 7:f0 + 9]
        lin_damp = 0.04                                      # Bullet's default linear damping also acts (environments.py:421-422 zeroes only the arm's)
        dec = (np.linalg.norm(v0) - np.linalg.norm(v)) / (n * dt)
        # 50 truncated sweeps share the load between four redundant points: within 10 % of mu g
        assert abs(dec - mu * g) < 0.10 * mu * g + lin_damp * np.linalg.norm(v0) * 1.5, (dec, mu * g)
        u0, u1 = v0 / np.linalg.norm(v0), v / np.linalg.norm(v)
        if vx == 0.0 or vy == 0.0:
            assert abs(u0[0] * u1[1] - u0[1] * u1[0]) < 2e-2  # along a symmetry axis of the box the motion keeps its direction
        assert abs(o.state[f0 + 9]) < 1e-3                   # stays on the table
    o, f0 = _block_on_table(m, 0.0, 0.0)
    z0 = o.state[f0 + 2]
    o.state[f0 + 7] = 0.5 * mu * g * dt                      # a nudge the static friction absorbs within one substep
    o.substeps(3)
    assert np.abs(o.state[f0 + 7:f0 + 9]).max() < 1e-6 and abs(o.state[f0 + 2] - z0) < 1e-5


def test_resting_contact_carries_the_weight():
    """Block at rest on the table: the normal impulses of its (redundant) contact points sum to m g dt."""
    m = load_model('UR5PlayAbsRPY1Obj-v0')
    o, f0 = _block_on_table(m)
    o.substeps(1)
    J, B, sc, nc = o.last_rows()
    n0 = len(J) - 4 * nc + 0
    nrm = [i for i in range(len(J)) if sc[i, 4] > 1e9 and J[i, m['nd'] + 2] != 0 and np.abs(J[i, m['nd'] + 6:m['nd'] + 12]).max() == 0]
    w = m['free_mass'][0] * -m.param('gravity_z') * m.param('dt')
    assert len(nrm) >= 3
    assert abs(sc[nrm, 5].sum() - w) < 0.02 * w, (sc[nrm, 5].sum(), w)


@pytest.mark.parametrize('impl', ['oracle', 'emu'])
def test_free_flight_closed_form_and_drift(impl):
    """No contacts, damping switched off in a model variant: semi-implicit Euler gives v_n = v_0 + n g dt and
    z_n = z_0 + dt sum v_k exactly; a spinning asymmetric body keeps its angular momentum (world frame) and kinetic energy
    to the accuracy of the explicit gyroscopic term over 0.3 s."""
    m0 = load_model('UR5PlayAbsRPY1Obj-v0')
    m = _variant('UR5PlayAbsRPY1Obj-v0', free_lin_damp=np.zeros_like(m0['free_lin_damp']), free_ang_damp=np.zeros_like(m0['free_ang_damp']))
    nd, dt, g = m['nd'], m.param('dt'), m.param('gravity_z')
    o = Oracle(m, seed=1)
    o.reset()
    f0 = 5 * nd
    st = o.state.copy()
    st[f0:f0 + 3] = [0.0, 0.2, 1.0]
    st[f0 + 3:f0 + 7] = [0.1, -0.2, 0.3, 0.9273618495495703]
    st[f0 + 7:f0 + 13] = [0.3, -0.1, 0.5, 2.0, -1.0, 1.5]
    I = np.array(m['free_inertia'][:3], np.float64)

    def Lw(s):
        x, y, z, w = s[f0 + 3:f0 + 7]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        wb = R.T @ s[f0 + 10:f0 + 13]
        return R @ (I * wb), 0.5 * float(wb @ (I * wb))
    n = 90
    if impl == 'oracle':
        o.state[:] = st
        traj = []
        for _ in range(n):
            o.substeps(1)
            traj.append(o.state.copy())
        tol_v = 1e-12
    else:
        from emu_lib import EmuSim
        sim = EmuSim(m, 1)
        sim.state[0, :len(st)] = st.astype(np.float32)
        traj = []
        for _ in range(n):
            sim.substeps(1)
            traj.append(sim.state[0, :len(st)].astype(np.float64))
        tol_v = 2e-5
    traj = np.array(traj)
    k = np.arange(1, n + 1)
    assert np.abs(traj[:, f0 + 9] - (0.5 + k * g * dt)).max() < tol_v * 10
    assert np.abs(traj[:, f0 + 7] - 0.3).max() < tol_v and np.abs(traj[:, f0 + 8] + 0.1).max() < tol_v
    z_ref = 1.0 + dt * np.cumsum(0.5 + k * g * dt)
    assert np.abs(traj[:, f0 + 2] - z_ref).max() < tol_v * 10
    L0, E0 = Lw(st)
    L1, E1 = Lw(traj[-1])
    assert np.linalg.norm(L1 - L0) < 2e-2 * np.linalg.norm(L0), (L0, L1)
    assert abs(E1 - E0) < 3e-2 * E0, (E0, E1)
    q = traj[:, f0 + 3:f0 + 7]
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-5


def _support(R, h, d):
    return np.abs(d @ R) @ h                      # half-extent of the box along direction(s) d


def test_box_box_against_direction_sampling():
    """Penetration depth = min over unit directions of the overlap of the two boxes' projections (support functions);
    200k sampled directions + the 15 candidate axes refined locally bound it from above to ~1e-3.  Bullet's detector (restated
    in the oracle and ported to the kernels) must report that depth (its 1.05 fudge may prefer a face axis that is up to
    5 % deeper), a unit normal pointing from box 2 to box 1 along such a direction, and contact points that lie on box 2
    and inside box 1 up to the depth."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(11)
    dirs = rng.standard_normal((200000, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    n_checked = 0
    for trial in range(120):
        R1 = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        R2 = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        h1, h2 = rng.uniform(0.02, 0.1, 3), rng.uniform(0.02, 0.1, 3)
        p2 = rng.uniform(-0.12, 0.12, 3)
        ov = _support(R1, h1, dirs) + _support(R2, h2, dirs) - np.abs(dirs @ p2)
        i = int(np.argmin(ov))
        est = float(ov[i])
        # local refinement around the best sampled direction
        d = dirs[i]
        for scale in (3e-2, 1e-2, 3e-3, 1e-3):
            cand = d + scale * rng.standard_normal((4000, 3))
            cand /= np.linalg.norm(cand, axis=1, keepdims=True)
            o2 = _support(R1, h1, cand) + _support(R2, h2, cand) - np.abs(cand @ p2)
            j = int(np.argmin(o2))
            if o2[j] < est:
                est, d = float(o2[j]), cand[j]
        c = orc_box_box([0, 0, 0], R1.reshape(-1), h1, p2, R2.reshape(-1), h2)
        if est < -2e-3:
            assert len(c) == 0                      # a separating direction exists
            continue
        if est < 2e-3:
            continue                                # touching: either answer is within the sampling error
        n_checked += 1
        assert len(c) >= 1
        depth = c[:, 6].max()
        assert est * (1 - 1e-2) - 1e-4 <= depth <= 1.05 * est + 1e-3, (depth, est)
        nrm = c[0, 3:6]
        assert abs(np.linalg.norm(nrm) - 1) < 1e-9
        assert nrm @ (-p2) > -1e-9                  # from box 2 towards box 1
        # the reported normal is itself a near-minimal overlap direction
        ov_n = _support(R1, h1, nrm[None]) + _support(R2, h2, nrm[None]) - abs(nrm @ p2)
        assert ov_n[0] <= 1.05 * est + 1e-3
        for pt in c[:, :3]:
            l2 = R2.T @ (pt - p2)
            assert (np.abs(l2) <= h2 + 1e-6).all()              # on / in box 2
            l1 = R1.T @ pt
            assert (np.abs(l1) <= h1 + depth + 1e-6).all()      # within the depth of box 1
    assert n_checked > 25
