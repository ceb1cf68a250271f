"""URDF -> reduced articulated-arm description (host side, one-time, fp64).

Replaces what the reference obtains from `loadURDF` (environments.py:397,409;
inverseKinematics.py:18) plus the importer behaviour it relies on:

* links are indexed by pre-order DFS with children in file order (matches the
  22-row table recorded in testing_bullet_ik.ipynb cell 2);
* no URDF_USE_INERTIA_FROM_FILE flag (environments.py:327) => link inertia is
  recomputed from the collision geometry as a box of the collision AABB in the
  inertial frame; links without <inertial> get mass 1 and an identity inertial
  frame; links without collision get the inertia of a 1 mm-margin point box;
* fixed joints carry no DoF; they are folded into the nearest movable ancestor
  (mass, centre of mass, inertia tensor, colliders and "sites"), which leaves the
  rigid-body dynamics unchanged;
* every collision geometry is reduced to an oriented box: <box> exactly,
  <cylinder> by its bounding box, meshes by the AABB of their vertices in the
  collision frame (Bullet uses the vertices' convex hull; see DESIGN.md
  "deviations").
"""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from .geom import Xf, rpy_to_mat, box_inertia, parallel_axis

URDF_MARGIN = 0.001  # Bullet's gUrdfDefaultCollisionMargin


def _floats(s, n=3, default=0.0):
    if s is None:
        return np.full(n, default)
    return np.array([float(v) for v in s.split()], dtype=np.float64)


def _origin(el):
    if el is None:
        return Xf()
    o = el.find('origin')
    if o is None:
        return Xf()
    return Xf(rpy_to_mat(_floats(o.get('rpy'))), _floats(o.get('xyz')))


def mesh_vertices(fn):
    ext = fn.lower().rsplit('.', 1)[-1]
    if ext == 'obj':
        return np.array([[float(x) for x in l.split()[1:4]] for l in open(fn) if l.startswith('v ')],
                        dtype=np.float64)
    if ext == 'stl':
        d = open(fn, 'rb').read()
        if d[:5] == b'solid' and b'facet' in d[:400]:
            V = [[float(x) for x in l.split()[1:4]] for l in d.decode(errors='ignore').splitlines()
                 if l.strip().startswith('vertex')]
            return np.array(V, dtype=np.float64)
        n = struct.unpack('<I', d[80:84])[0]
        a = np.frombuffer(d[84:84 + 50 * n], dtype=np.uint8).reshape(n, 50)
        f = a[:, :48].copy().view('<f4').reshape(n, 12)
        return f[:, 3:].reshape(-1, 3).astype(np.float64)
    raise ValueError('unsupported mesh format: ' + fn)


class UrdfLink:
    pass


def parse_urdf(path):
    """Returns links in PyBullet order: index -1 = base, 0.. = DFS pre-order."""
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(path)
    links = {l.get('name'): l for l in root.findall('link')}
    joints = root.findall('joint')
    children = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.find('parent').get('link'), []).append(j)
        child_names.add(j.find('child').get('link'))
    roots = [n for n in links if n not in child_names]
    assert len(roots) == 1, roots
    out = []

    def geom_boxes(link_el):
        boxes = []  # (Xf collision frame in link frame, centre in coll frame, half extents)
        for c in link_el.findall('collision'):
            X = _origin(c)
            g = c.find('geometry')[0]
            if g.tag == 'box':
                boxes.append((X, np.zeros(3), _floats(g.get('size')) / 2))
            elif g.tag == 'cylinder':
                r, L = float(g.get('radius')), float(g.get('length'))
                boxes.append((X, np.zeros(3), np.array([r, r, L / 2])))
            elif g.tag == 'sphere':
                r = float(g.get('radius'))
                boxes.append((X, np.zeros(3), np.array([r, r, r])))
            elif g.tag == 'mesh':
                fn = g.get('filename').replace('package://', '')
                V = mesh_vertices(os.path.join(base_dir, fn)) * _floats(g.get('scale'), 3, 1.0)
                lo, hi = V.min(0), V.max(0)
                boxes.append((X, (lo + hi) / 2, (hi - lo) / 2))
            else:
                raise ValueError(g.tag)
        return boxes

    def make_link(name, joint_el, parent_idx):
        el = links[name]
        L = UrdfLink()
        L.name = name
        L.parent = parent_idx
        L.boxes = geom_boxes(el)
        inert = el.find('inertial')
        if inert is not None:
            L.mass = float(inert.find('mass').get('value'))
            L.inertial = _origin(inert)
        else:
            L.mass = 1.0  # Bullet's UrdfParser default for a link with no <inertial>
            L.inertial = Xf()
        # inertia recomputed from the collision AABB expressed in the inertial frame
        if L.mass > 0:
            pts = []
            for X, c, h in L.boxes:
                Xi = L.inertial.inv() * X
                for sx in (-1, 1):
                    for sy in (-1, 1):
                        for sz in (-1, 1):
                            pts.append(Xi.apply(c + h * np.array([sx, sy, sz])))
            if pts:
                pts = np.array(pts)
                dims = (pts.max(0) - pts.min(0)) + 2 * 3 * URDF_MARGIN
            else:
                dims = np.full(3, 2 * URDF_MARGIN)
            L.inertia_diag = box_inertia(L.mass, dims)
        else:
            L.inertia_diag = np.zeros(3)
        ct = el.find('contact')
        L.lateral_friction = 0.5
        L.spinning_friction = 0.0
        L.stiffness = -1.0
        L.damping = -1.0
        if ct is not None:
            for e in ct:
                v = e.get('value')
                if e.tag == 'lateral_friction':
                    L.lateral_friction = float(v)
                elif e.tag == 'spinning_friction':
                    L.spinning_friction = float(v)
                elif e.tag == 'stiffness':
                    L.stiffness = float(v)
                elif e.tag == 'damping':
                    L.damping = float(v)
        if joint_el is None:
            L.jtype = 'base'
            L.X = Xf()
            L.axis = np.array([0, 0, 1.0])
            L.lo, L.hi = 0.0, -1.0
            L.jdamping = 0.0
            L.jname = ''
        else:
            L.jname = joint_el.get('name')
            L.jtype = joint_el.get('type')
            L.X = _origin(joint_el)
            ax = joint_el.find('axis')
            L.axis = _floats(ax.get('xyz')) if ax is not None else np.array([1.0, 0, 0])
            n = np.linalg.norm(L.axis)
            if n > 0:
                L.axis = L.axis / n
            lim = joint_el.find('limit')
            if lim is not None and L.jtype in ('revolute', 'prismatic'):
                L.lo = float(lim.get('lower', 0))
                L.hi = float(lim.get('upper', 0))
            else:
                L.lo, L.hi = 0.0, -1.0
            dyn = joint_el.find('dynamics')
            L.jdamping = float(dyn.get('damping', 0)) if dyn is not None else 0.0
        return L

    def visit(name, joint_el, parent_idx):
        L = make_link(name, joint_el, parent_idx)
        my = len(out) - 1  # base gets -1
        out.append(L)
        L.index = my
        for j in children.get(name, []):
            visit(j.find('child').get('link'), j, my)

    visit(roots[0], None, -2)
    return out  # out[0] is the base (index -1), out[k+1] has PyBullet link index k


def reduce_arm(urdf_links):
    """Fold fixed joints.  Returns dict of arrays over the movable links (DoF order =
    ascending PyBullet link index, which is Bullet's own DoF order) plus colliders
    and per-URDF-link sites."""
    by_index = {L.index: L for L in urdf_links}
    movable = [L for L in urdf_links if L.jtype in ('revolute', 'prismatic', 'continuous')]
    mov_of = {}     # urdf index -> (movable dof idx or -1 for base, Xf link frame in movable frame)
    dof_of = {L.index: k for k, L in enumerate(movable)}

    def resolve(idx):
        if idx in mov_of:
            return mov_of[idx]
        L = by_index[idx]
        if idx == -1:
            r = (-1, Xf())
        elif idx in dof_of:
            r = (dof_of[idx], Xf())
        else:
            pm, pX = resolve(L.parent)
            r = (pm, pX * L.X)
        mov_of[idx] = r
        return r

    nd = len(movable)
    acc = [dict(m=0.0, mc=np.zeros(3), parts=[]) for _ in range(nd)]
    colliders = []
    sites = {}
    for L in urdf_links:
        m_idx, X = resolve(L.index)
        # site = the link's inertial ("COM") frame, which is what getLinkState()[0:2] reports
        sites[L.index] = (m_idx, X * L.inertial)
        for Xc, c, h in L.boxes:
            Xb = X * Xc * Xf(np.eye(3), c)
            colliders.append(dict(link=m_idx, urdf_link=L.index, R=Xb.R, p=Xb.p, half=h,
                                  friction=L.lateral_friction, spin=L.spinning_friction,
                                  stiffness=L.stiffness, damping=L.damping))
        if m_idx >= 0 and L.mass > 0:
            Xi = X * L.inertial
            I = Xi.R @ np.diag(L.inertia_diag) @ Xi.R.T
            acc[m_idx]['parts'].append((L.mass, Xi.p, I))
            acc[m_idx]['m'] += L.mass
            acc[m_idx]['mc'] += L.mass * Xi.p
    arm = dict(nd=nd,
               parent=np.zeros(nd, np.int32), jtype=np.zeros(nd, np.int32),
               jpos=np.zeros((nd, 3)), jrot=np.zeros((nd, 3, 3)), axis=np.zeros((nd, 3)),
               com=np.zeros((nd, 3)), mass=np.zeros(nd), inertia=np.zeros((nd, 3, 3)),
               lo=np.zeros(nd), hi=np.zeros(nd), jdamp=np.zeros(nd),
               urdf_index=np.zeros(nd, np.int32), names=[])
    for k, L in enumerate(movable):
        pm, pX = resolve(L.parent)
        Xj = pX * L.X
        arm['parent'][k] = pm
        arm['jtype'][k] = 1 if L.jtype == 'prismatic' else 0
        arm['jpos'][k] = Xj.p
        arm['jrot'][k] = Xj.R
        arm['axis'][k] = L.axis
        m = acc[k]['m']
        c = acc[k]['mc'] / m
        I = np.zeros((3, 3))
        for (mi, pi, Ii) in acc[k]['parts']:
            I += Ii + parallel_axis(mi, pi - c)
        arm['com'][k] = c
        arm['mass'][k] = m
        arm['inertia'][k] = I
        arm['lo'][k], arm['hi'][k] = L.lo, L.hi
        arm['jdamp'][k] = L.jdamping
        arm['urdf_index'][k] = L.index
        arm['names'].append(L.jname)
    return arm, colliders, sites
