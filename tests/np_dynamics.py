"""Independent numpy fp64 rigid-body dynamics (composite-rigid-body + recursive Newton-Euler,
world frame, COM-referenced) used by the CPU tests to cross-check the oracle's
articulated-body algorithm.  Same formulation the CUDA kernels use."""
import numpy as np


def _rot(axis, q):
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K


def fk(m, q):
    nd = m['nd']
    R = [None] * nd; P = [None] * nd; A = [None] * nd; C = [None] * nd
    bR = m['arm_base_rot'].reshape(3, 3); bp = m['arm_base_pos']
    jrot = m['arm_jrot'].reshape(nd, 3, 3); jpos = m['arm_jpos'].reshape(nd, 3)
    axis = m['arm_axis'].reshape(nd, 3); com = m['arm_com'].reshape(nd, 3)
    for i in range(nd):
        p = m['arm_parent'][i]
        pR, pp = (bR, bp) if p < 0 else (R[p], P[p])
        if m['arm_jtype'][i] == 0:
            R[i] = pR @ jrot[i] @ _rot(axis[i], q[i]); P[i] = pR @ jpos[i] + pp
        else:
            R[i] = pR @ jrot[i]; P[i] = pR @ jpos[i] + pp + R[i] @ axis[i] * q[i]
        A[i] = R[i] @ axis[i]; C[i] = P[i] + R[i] @ com[i]
    return R, P, A, C


def mass_matrix(m, q):
    nd = m['nd']
    R, P, A, C = fk(m, q)
    mass = m['arm_mass']; I = m['arm_inertia'].reshape(nd, 3, 3)
    cm = [mass[i] for i in range(nd)]; cc = [C[i].copy() for i in range(nd)]
    cI = [R[i] @ I[i] @ R[i].T for i in range(nd)]
    par = m['arm_parent']
    for i in range(nd - 1, -1, -1):
        p = par[i]
        if p >= 0:
            mt = cm[p] + cm[i]; c = (cm[p] * cc[p] + cm[i] * cc[i]) / mt
            def pa(mm, d): return mm * (d @ d * np.eye(3) - np.outer(d, d))
            cI[p] = cI[p] + pa(cm[p], cc[p] - c) + cI[i] + pa(cm[i], cc[i] - c)
            cm[p] = mt; cc[p] = c
    M = np.zeros((nd, nd))
    for i in range(nd):
        if m['arm_jtype'][i] == 0:
            w = A[i]; vc = np.cross(A[i], cc[i] - P[i])
        else:
            w = np.zeros(3); vc = A[i]
        lin = cm[i] * vc; ang = cI[i] @ w
        j = i
        while j >= 0:
            if m['arm_jtype'][j] == 0:
                M[i, j] = M[j, i] = A[j] @ (ang + np.cross(cc[i] - P[j], lin))
            else:
                M[i, j] = M[j, i] = A[j] @ lin
            j = par[j]
    return M


def bias(m, q, qd, g=-9.8):
    """tau such that M qdd + bias = 0 for the free arm (gravity + Coriolis/centrifugal)."""
    nd = m['nd']
    R, P, A, C = fk(m, q)
    mass = m['arm_mass']; I = m['arm_inertia'].reshape(nd, 3, 3); par = m['arm_parent']
    w = [None] * nd; al = [None] * nd; ap = [None] * nd; vp = [None] * nd  # ang vel, ang acc, acc of link origin, vel of origin
    f = [None] * nd; n = [None] * nd
    for i in range(nd):
        p = par[i]
        if p < 0:
            wp = np.zeros(3); alp = np.zeros(3); app = np.array([0, 0, -g]); Pp = m['arm_base_pos']
        else:
            wp, alp, app, Pp = w[p], al[p], ap[p], P[p]
        r = P[i] - Pp
        if m['arm_jtype'][i] == 0:
            # link origin is fixed in the parent
            a_org = app + np.cross(alp, r) + np.cross(wp, np.cross(wp, r))
            w[i] = wp + A[i] * qd[i]
            al[i] = alp + np.cross(wp, A[i] * qd[i])
            ap[i] = a_org
        else:
            # origin slides along the axis: r = r0 + a q
            w[i] = wp; al[i] = alp
            ap[i] = app + np.cross(alp, r) + np.cross(wp, np.cross(wp, r)) + 2 * np.cross(wp, A[i] * qd[i])
        rc = C[i] - P[i]
        ac = ap[i] + np.cross(al[i], rc) + np.cross(w[i], np.cross(w[i], rc))
        Iw = R[i] @ I[i] @ R[i].T
        f[i] = mass[i] * ac
        n[i] = Iw @ al[i] + np.cross(w[i], Iw @ w[i])   # about COM
    tau = np.zeros(nd)
    F = [f[i].copy() for i in range(nd)]
    N = [n[i] + np.cross(C[i] - P[i], f[i]) for i in range(nd)]   # moment about link origin
    for i in range(nd - 1, -1, -1):
        tau[i] = A[i] @ N[i] if m['arm_jtype'][i] == 0 else A[i] @ F[i]
        p = par[i]
        if p >= 0:
            F[p] += F[i]; N[p] += N[i] + np.cross(P[i] - P[p], F[i])
    return tau
